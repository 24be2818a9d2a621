"""VAE decode on the B200 kernels: latents -> image ("VAE-decode path").

The reference lists the VAE among its planned integrations and ships nothing for it (README.md:5,
implementations/ComfyUI/example.py is empty); in its Diffusers integration the pipeline's own
`AutoencoderKL.decode` runs in eager PyTorch after the loop (implementations/Diffusers/load_sdxl_pipeline.py:39,46).
This module is that decoder -- Diffusers' `AutoencoderKL` decode half with the SDXL VAE configuration and its state-dict
key names (`post_quant_conv.*`, `decoder.conv_in.*`, `decoder.mid_block.{resnets,attentions}.*`,
`decoder.up_blocks.N.{resnets,upsamplers}.*`, `decoder.conv_norm_out.*`, `decoder.conv_out.*`) -- with two forwards:

  * `forward(z)`      eager PyTorch, the model definition (what the oracle restates and the tests compare against);
  * `compile_vae(m)`  the same network on the hand-written sm_100a kernels: 3x3 convolutions as implicit GEMMs with
                      bias / residual epilogues that also emit the next GroupNorm's statistics, GroupNorm(+SiLU) from
                      those partials, nearest-2x upsampling, and the mid-block attention -- ONE head of width 512 over
                      H*W tokens, which no flash kernel of this library covers -- as Q K^T on the tensor-core GEMM with
                      an fp32 result, a row softmax, and P V on the GEMM again.  One CUDA graph per input shape.

Third party (Diffusers), so parity is pinned against this repo's own fp32 restatement (oracle/vae_oracle.py): unpinned.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import kernels as K


@dataclasses.dataclass(frozen=True)
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.13025

    @staticmethod
    def sdxl() -> "VAEConfig":
        return VAEConfig()

    @staticmethod
    def tiny() -> "VAEConfig":
        """Same topology at toy widths (channel counts stay multiples of 64: the tensor-core conv tiles K by 64)."""
        return VAEConfig(block_out_channels=(64, 128), layers_per_block=1)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class AttentionBlock(nn.Module):
    """Diffusers `Attention(channels, heads=1, dim_head=channels, norm_num_groups, residual_connection=True)`."""

    def __init__(self, channels: int, groups: int, eps: float):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax(q @ k.transpose(1, 2) * (1.0 / math.sqrt(c)), dim=-1)
        o = self.to_out[0](p @ v)
        return x + o.transpose(1, 2).reshape(b, c, h, w)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels: int, groups: int, eps: float):
        super().__init__()
        self.attentions = nn.ModuleList([AttentionBlock(channels, groups, eps)])
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, groups, eps) for _ in range(2)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, layers: int, groups: int, eps: float, upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups, eps) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return x if self.upsamplers is None else self.upsamplers[0](x)


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        ch = tuple(reversed(cfg.block_out_channels))
        g, eps = cfg.norm_num_groups, cfg.norm_eps
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(ch[0], g, eps)
        blocks, prev = [], ch[0]
        for i, c in enumerate(ch):
            blocks.append(UpDecoderBlock2D(prev, c, cfg.layers_per_block + 1, g, eps, upsample=i < len(ch) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=eps)
        self.conv_out = nn.Conv2d(ch[-1], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            h = blk(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKLDecoder(nn.Module):
    """`forward(latents) -> image`: Diffusers `AutoencoderKL.decode(latents / scaling_factor).sample`."""

    def __init__(self, cfg: Optional[VAEConfig] = None):
        super().__init__()
        self.cfg = cfg or VAEConfig.sdxl()
        self.post_quant_conv = nn.Conv2d(self.cfg.latent_channels, self.cfg.latent_channels, 1)
        self.decoder = Decoder(self.cfg)

    def forward(self, latents):
        return self.decoder(self.post_quant_conv(latents / self.cfg.scaling_factor))


def build_vae_decoder(cfg: Optional[VAEConfig] = None, seed: int = 0, device="cuda", dtype=torch.bfloat16) -> AutoencoderKLDecoder:
    """An `AutoencoderKLDecoder` with synthetic (hash-seeded, CPU/CUDA bit-identical) weights."""
    from . import synth

    cfg = cfg or VAEConfig.sdxl()
    with torch.device("meta"):
        model = AutoencoderKLDecoder(cfg)
    model.load_state_dict(synth.synth_state_dict(model, seed=seed, device=device, dtype=dtype), strict=True, assign=True)
    for p in model.parameters():
        p.requires_grad_(False)
    return model.eval()


# ------------------------------------------------------------------------------------------------
# The kernel path
# ------------------------------------------------------------------------------------------------
class CompiledVAEDecoder:
    """`decode(latents (B, 4, h, w)) -> image (B, 3, 8h, 8w)` bf16, every op on the sm_100a kernels, one CUDA graph per
    latent shape.  Weights are read live from the module (conv weights re-laid-out once to KRSC, values untouched; the
    q / k / v projections of the attention block row-concatenated into one GEMM operand)."""

    ATTN_ROWS = 4096  # query rows per score block: bounds the fp32 score matrix to ATTN_ROWS x T

    def __init__(self, model: AutoencoderKLDecoder, cuda_graph: bool = True):
        p = next(model.parameters())
        if p.device.type != "cuda":
            raise AssertionError("compile_vae: the model must live on a CUDA device (there is no CPU path)")
        if p.dtype != torch.bfloat16:
            raise TypeError("compile_vae: parameters must be bfloat16")
        self.model = model
        self.cfg = model.cfg
        self.cuda_graph = cuda_graph
        self._graphs: Dict[tuple, tuple] = {}
        with torch.no_grad():
            for m in model.modules():
                if isinstance(m, nn.Conv2d) and m.kernel_size == (3, 3) and m.in_channels > 7 and m.out_channels > 8:
                    m.weight.data = K.pack_conv_weight(m.weight.data)
            att = model.decoder.mid_block.attentions[0]
            self._wqkv = torch.cat([att.to_q.weight, att.to_k.weight, att.to_v.weight], dim=0).contiguous()
            self._bqkv = torch.cat([att.to_q.bias, att.to_k.bias, att.to_v.bias], dim=0).contiguous()

    # -- building blocks (x: logical NCHW, NHWC storage; part: GroupNorm partials of x from its producer, or None) -----
    def _gn(self, x, norm: nn.GroupNorm, silu: bool, part):
        return K.groupnorm_wrapper(x, norm.num_groups, norm.weight, norm.bias, norm.eps, silu,
                                   partials=(part,) if part is not None else None)

    def _resnet(self, x, part, blk: ResnetBlock2D):
        h = self._gn(x, blk.norm1, True, part)
        h, hp = K.conv2d(h, blk.conv1.weight, blk.conv1.bias, w_static=True, gn_stats=True)
        h = self._gn(h, blk.norm2, True, hp)
        sc = x if blk.conv_shortcut is None else K.conv2d(x, blk.conv_shortcut.weight, blk.conv_shortcut.bias, padding=0,
                                                           w_static=True)
        return K.conv2d(h, blk.conv2.weight, blk.conv2.bias, residual=sc, w_static=True, gn_stats=True)

    def _attention(self, x, part, att: AttentionBlock):
        b, c, h, w = x.shape
        t = h * w
        hn = self._gn(x, att.group_norm, False, part)
        tok = hn.permute(0, 2, 3, 1).reshape(b, t, c)          # NHWC storage: a view
        res = x.permute(0, 2, 3, 1).reshape(b, t, c)
        qkv = K.linear(tok, self._wqkv, self._bqkv, w_static=True)  # (b, t, 3c)
        q, k, v = qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]
        vt = K.transpose_tokens(v)                              # (b, c, t): the B operand of P @ V
        o = torch.empty((b, t, c), dtype=x.dtype, device=x.device)
        scale = 1.0 / math.sqrt(c)
        for i in range(b):
            for r0 in range(0, t, self.ATTN_ROWS):
                r1 = min(t, r0 + self.ATTN_ROWS)
                s = K.matmul_nt_f32(q[i, r0:r1], k[i])          # fp32 scores (rows, t)
                pr = K.softmax_rows(s, scale)
                K.linear(pr, vt[i], out=o[i, r0:r1])
        y, yp = K.linear(o, att.to_out[0].weight, att.to_out[0].bias, residual=res, w_static=True, gn_stats=t)
        return y.reshape(b, h, w, c).permute(0, 3, 1, 2), yp

    def _forward(self, latents: torch.Tensor, pre_scaled: bool = False) -> torch.Tensor:
        m, d = self.model, self.model.decoder
        z = K.pointwise_conv_small(latents, m.post_quant_conv.weight, m.post_quant_conv.bias,
                                   1.0 if pre_scaled else 1.0 / self.cfg.scaling_factor)
        x, part = K.conv2d(z, d.conv_in.weight, d.conv_in.bias, w_static=True, gn_stats=True)
        x, part = self._resnet(x, part, d.mid_block.resnets[0])
        x, part = self._attention(x, part, d.mid_block.attentions[0])
        x, part = self._resnet(x, part, d.mid_block.resnets[1])
        for blk in d.up_blocks:
            for r in blk.resnets:
                x, part = self._resnet(x, part, r)
            if blk.upsamplers is not None:
                up = K.upsample_nearest2x(x)
                x, part = K.conv2d(up, blk.upsamplers[0].conv.weight, blk.upsamplers[0].conv.bias, w_static=True, gn_stats=True)
        x = self._gn(x, d.conv_norm_out, True, part)
        return K.conv2d(x, d.conv_out.weight, d.conv_out.bias, nchw_output=True, w_static=True)

    @torch.no_grad()
    def eager_decode(self, latents: torch.Tensor, pre_scaled: bool = False) -> torch.Tensor:
        """The launch sequence on the current stream, no graph."""
        return self._forward(self._check(latents), pre_scaled)

    def _check(self, latents):
        if latents.device.type != "cuda" or latents.dtype != torch.bfloat16 or latents.dim() != 4 \
                or latents.shape[1] != self.cfg.latent_channels:
            raise ValueError(f"decode: expected a bf16 CUDA tensor (B, {self.cfg.latent_channels}, h, w), "
                             f"got {latents.dtype} {tuple(latents.shape)} on {latents.device}")
        return latents.contiguous()

    @torch.no_grad()
    def decode(self, latents: torch.Tensor, pre_scaled: bool = False) -> torch.Tensor:
        """pre_scaled: `latents` have already been divided by the scaling factor, as a Diffusers pipeline does before it
        calls `vae.decode` (pipeline_stable_diffusion_xl.py: `self.vae.decode(latents / self.vae.config.scaling_factor)`)."""
        latents = self._check(latents)
        if not self.cuda_graph:
            return self._forward(latents, pre_scaled)
        key = (tuple(latents.shape), latents.device.index, bool(pre_scaled))
        hit = self._graphs.get(key)
        if hit is None:
            static_in = latents.clone()
            side = torch.cuda.Stream(latents.device)
            side.wait_stream(torch.cuda.current_stream(latents.device))
            with torch.cuda.stream(side):
                for _ in range(2):  # lazy initialisation (padded conv_in / conv_out operands, workspace) stays out of the capture
                    self._forward(static_in, pre_scaled)
                torch.cuda.synchronize(latents.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
                    static_out = self._forward(static_in, pre_scaled)
            torch.cuda.synchronize(latents.device)
            hit = (graph, static_in, static_out)
            self._graphs[key] = hit
        graph, static_in, static_out = hit
        static_in.copy_(latents)
        graph.replay()
        return static_out.clone()

    __call__ = decode


def compile_vae(model: AutoencoderKLDecoder, cuda_graph: bool = True) -> CompiledVAEDecoder:
    """The VAE counterpart of `compile(unet)`: same model object, every op on the B200 kernels."""
    return CompiledVAEDecoder(model, cuda_graph=cuda_graph)
