"""SDXL `UNet2DConditionModel` as plain, fx-traceable PyTorch, with Diffusers-compatible parameter names.

This is the model definition `compile()` is applied to.  It plays the role of the reference's
`optimizers/unet_pt.py:416-541` (same architecture, same `state_dict` keys, same forward signature,
same eager op idioms at every fusion site so that one set of fx passes serves both), but it is
driven by a `UNetConfig` instead of hard-coded sizes, so small variants exist for tests and the
2048^2 / batch sweeps need no code change.  `UNetConfig.sdxl()` reproduces SDXL-base exactly
(2 567 463 684 parameters, 1680 tensors; cross-checked against implementations/sgm_/config.yaml:19-37).
"""
from __future__ import annotations

import math
from collections import namedtuple
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Sequence[int] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers_per_block: Sequence[int] = (0, 2, 10)  # 0 -> block without attention
    attention_head_dim: int = 64
    cross_attention_dim: int = 2048
    norm_num_groups: int = 32
    addition_time_embed_dim: int = 256
    text_embed_dim: int = 1280
    num_time_ids: int = 6
    sample_size: int = 128

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def add_embed_in_dim(self) -> int:  # Diffusers: projection_class_embeddings_input_dim (2816 for SDXL)
        return self.text_embed_dim + self.num_time_ids * self.addition_time_embed_dim

    @staticmethod
    def sdxl() -> "UNetConfig":
        return UNetConfig()

    @staticmethod
    def tiny() -> "UNetConfig":
        """A few-million-parameter variant with every structural feature of SDXL (all block kinds,
        shortcut convs, down/up-sampling, self+cross attention, GEGLU) and kernel-legal sizes."""
        return UNetConfig(block_out_channels=(64, 128, 256), layers_per_block=2, transformer_layers_per_block=(0, 1, 2),
                          attention_head_dim=64, cross_attention_dim=128, norm_num_groups=8,
                          addition_time_embed_dim=32, text_embed_dim=64, sample_size=32)


class Timesteps(nn.Module):
    """Sinusoidal embedding, cos half first (reference: unet_pt.py:17-36)."""

    def __init__(self, num_channels: int = 320):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps: torch.Tensor) -> torch.Tensor:
        half = self.num_channels // 2
        freqs = torch.exp(
            -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / (half - 0.0)
        )
        angles = timesteps[:, None].float() * freqs[None, :]
        return torch.cat([torch.cos(angles), torch.sin(angles)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, out_features)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(out_features, out_features)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class ResnetBlock2D(nn.Module):
    """GN-SiLU-conv3x3 (+time embedding) -> GN-SiLU-conv3x3, plus identity / 1x1 shortcut
    (reference: unet_pt.py:54-95)."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=1e-5, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=1e-5, affine=True)
        self.dropout = nn.Dropout(p=0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = (
            nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1) if in_channels != out_channels else None
        )

    def forward(self, input_tensor, temb):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        t = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = h + t
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return input_tensor + h


class Attention(nn.Module):
    """Multi-head attention with bias-free q/k/v projections (reference: unet_pt.py:98-147)."""

    def __init__(self, inner_dim: int, cross_attention_dim: Optional[int] = None, head_dim: int = 64):
        super().__init__()
        self.head_dim = head_dim
        self.num_heads = inner_dim // head_dim
        self.scale = head_dim ** -0.5
        kv_dim = inner_dim if cross_attention_dim is None else cross_attention_dim
        self.to_q = nn.Linear(inner_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(kv_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(kv_dim, inner_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner_dim, inner_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None):
        context = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self.to_q(hidden_states)
        k = self.to_k(context)
        v = self.to_v(context)
        b, t, c = q.size()
        q = q.view(q.size(0), q.size(1), self.num_heads, self.head_dim).transpose(1, 2)
        k = k.view(k.size(0), k.size(1), self.num_heads, self.head_dim).transpose(1, 2)
        v = v.view(v.size(0), v.size(1), self.num_heads, self.head_dim).transpose(1, 2)
        probs = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) * self.scale, dim=-1)
        out = torch.matmul(probs, v).transpose(1, 2).contiguous().view(b, t, c)
        for layer in self.to_out:
            out = layer(out)
        return out


class GEGLU(nn.Module):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.proj = nn.Linear(in_features, out_features * 2)

    def forward(self, x):
        state, gate = self.proj(x).chunk(2, dim=-1)
        return state * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for layer in self.net:
            x = layer(x)
        return x


class BasicTransformerBlock(nn.Module):
    """LN-self-attn, LN-cross-attn, LN-GEGLU-FF, each with a residual (reference: unet_pt.py:179-210)."""

    def __init__(self, dim: int, cross_attention_dim: int, head_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, head_dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, head_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    """GN -> tokens -> proj_in -> blocks -> proj_out -> image, + residual (reference: unet_pt.py:213-243)."""

    def __init__(self, channels: int, n_layers: int, cross_attention_dim: int, head_dim: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, cross_attention_dim, head_dim) for _ in range(n_layers)]
        )
        self.proj_out = nn.Linear(channels, channels)

    def forward(self, hidden_states, encoder_hidden_states=None):
        batch, _, height, width = hidden_states.shape
        res = hidden_states
        hidden_states = self.norm(hidden_states)
        inner_dim = hidden_states.shape[1]
        hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, encoder_hidden_states)
        hidden_states = self.proj_out(hidden_states)
        hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        return hidden_states + res


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """`DownBlock2D` / `CrossAttnDownBlock2D` (reference: unet_pt.py:269-327), selected by n_tf_layers."""

    def __init__(self, cfg: UNetConfig, in_channels: int, out_channels: int, n_tf_layers: int, downsample: bool):
        super().__init__()
        temb = cfg.time_embed_dim
        if n_tf_layers > 0:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(out_channels, n_tf_layers, cfg.cross_attention_dim, cfg.attention_head_dim,
                                    cfg.norm_num_groups) for _ in range(cfg.layers_per_block)]
            )
        else:
            self.attentions = None
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb, cfg.norm_num_groups)
             for i in range(cfg.layers_per_block)]
        )
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if downsample else None

    def forward(self, hidden_states, temb, encoder_hidden_states=None):
        outputs = []
        for i, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb)
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states)
            outputs.append(hidden_states)
        if self.downsamplers is not None:
            hidden_states = self.downsamplers[0](hidden_states)
            outputs.append(hidden_states)
        return hidden_states, outputs


class UpBlock(nn.Module):
    """`UpBlock2D` / `CrossAttnUpBlock2D` (reference: unet_pt.py:330-388)."""

    def __init__(self, cfg: UNetConfig, in_channels: int, out_channels: int, prev_output_channel: int,
                 n_tf_layers: int, upsample: bool):
        super().__init__()
        temb = cfg.time_embed_dim
        n = cfg.layers_per_block + 1
        if n_tf_layers > 0:
            self.attentions = nn.ModuleList(
                [Transformer2DModel(out_channels, n_tf_layers, cfg.cross_attention_dim, cfg.attention_head_dim,
                                    cfg.norm_num_groups) for _ in range(n)]
            )
        else:
            self.attentions = None
        resnets = []
        for i in range(n):
            skip_channels = in_channels if i == n - 1 else out_channels
            main_channels = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(main_channels + skip_channels, out_channels, temb, cfg.norm_num_groups))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, encoder_hidden_states=None):
        for i, resnet in enumerate(self.resnets):
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, skip], dim=1)
            hidden_states = resnet(hidden_states, temb)
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states)
        if self.upsamplers is not None:
            hidden_states = self.upsamplers[0](hidden_states)
        return hidden_states


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, cfg: UNetConfig, channels: int, n_tf_layers: int):
        super().__init__()
        self.attentions = nn.ModuleList(
            [Transformer2DModel(channels, n_tf_layers, cfg.cross_attention_dim, cfg.attention_head_dim,
                                cfg.norm_num_groups)]
        )
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(channels, channels, cfg.time_embed_dim, cfg.norm_num_groups) for _ in range(2)]
        )

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.attentions[0](hidden_states, encoder_hidden_states=encoder_hidden_states)
        return self.resnets[1](hidden_states, temb)


class UNet2DConditionModel(nn.Module):
    """forward(sample, timesteps, encoder_hidden_states, added_cond_kwargs, **kwargs) -> [eps]
    (list-valued, so `unet(...)[0]` works in a Diffusers pipeline; reference: unet_pt.py:469-542)."""

    def __init__(self, cfg: Optional[UNetConfig] = None, adm_input: bool = False):
        super().__init__()
        cfg = cfg or UNetConfig.sdxl()
        self.cfg = cfg
        # adm_input: the micro-conditioning arrives already embedded, as ComfyUI / sgm pass it (`y` = [pooled text |
        # Fourier features of the size / crop ids], add_embed_in_dim wide) under added_cond_kwargs["adm"]; the
        # `add_time_proj` embedding of `time_ids` is then not part of the graph (stabletriton_b200/comfy.py)
        self.adm_input = bool(adm_input)
        # what a Diffusers pipeline reads from `unet.config` (reference: unet_pt.py:420-428)
        self.config = make_config_shim(cfg)

        ch = list(cfg.block_out_channels)
        tf = list(cfg.transformer_layers_per_block)
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], kernel_size=3, stride=1, padding=1)
        self.time_proj = Timesteps(ch[0])
        self.time_embedding = TimestepEmbedding(ch[0], cfg.time_embed_dim)
        self.add_time_proj = Timesteps(cfg.addition_time_embed_dim)
        self.add_embedding = TimestepEmbedding(cfg.add_embed_in_dim, cfg.time_embed_dim)

        down = []
        out_c = ch[0]
        for i in range(len(ch)):
            in_c, out_c = out_c, ch[i]
            down.append(DownBlock(cfg, in_c, out_c, tf[i], downsample=i < len(ch) - 1))
        self.down_blocks = nn.ModuleList(down)

        self.mid_block = UNetMidBlock2DCrossAttn(cfg, ch[-1], tf[-1])

        rev, rev_tf = ch[::-1], tf[::-1]
        up = []
        out_c = rev[0]
        for i in range(len(rev)):
            prev, out_c = out_c, rev[i]
            in_c = rev[min(i + 1, len(rev) - 1)]
            up.append(UpBlock(cfg, in_c, out_c, prev, rev_tf[i], upsample=i < len(rev) - 1))
        self.up_blocks = nn.ModuleList(up)

        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=1e-5, affine=True)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, sample, timesteps, encoder_hidden_states, added_cond_kwargs, **kwargs):
        timesteps = timesteps.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=sample.dtype))

        if self.adm_input:
            add_embeds = added_cond_kwargs.get("adm").to(emb.dtype)
        else:
            text_embeds = added_cond_kwargs.get("text_embeds")
            time_ids = added_cond_kwargs.get("time_ids")
            time_embeds = self.add_time_proj(time_ids.flatten()).reshape((text_embeds.shape[0], -1))
            add_embeds = torch.concat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
        emb = emb + self.add_embedding(add_embeds)

        sample = self.conv_in(sample)
        skips: List[torch.Tensor] = [sample]
        for block in self.down_blocks:
            sample, outs = block(sample, emb, encoder_hidden_states)
            skips = skips + outs
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        for block in self.up_blocks:
            n = len(block.resnets)
            sample = block(sample, skips[-n:], emb, encoder_hidden_states)
            skips = skips[:-n]

        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        return [sample]


def make_config_shim(cfg: UNetConfig):
    """The three attributes StableDiffusionXLPipeline reads off `unet.config`; the reference re-attaches
    them by hand after optimize_model (implementations/Diffusers/load_sdxl_pipeline.py:29-34)."""
    shim = namedtuple("config", "in_channels addition_time_embed_dim sample_size")
    shim.in_channels = cfg.in_channels
    shim.addition_time_embed_dim = cfg.addition_time_embed_dim
    shim.sample_size = cfg.sample_size
    return shim
