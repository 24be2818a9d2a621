"""ORACLE -- test infrastructure only.  A CPU fp32 restatement of the reference's SDXL UNet denoise
step, written as pure functions over a Diffusers-keyed state dict.

Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs of `bench.py` may import this
module; the product (`stabletriton_b200/`) never does.  Every function cites the reference lines it
restates (paths relative to /root/reference/src/stabletriton/).

Pinning: the restatement is checked against the reference's own `optimizers/unet_pt.py` -- block by
block and for the whole UNet on identical weights and inputs -- by `oracle/make_golden.py`, which
imports the reference in the build container and writes `tests/golden/*.pt`; `tests/test_oracle.py`
replays those fixtures anywhere (the GPU box has no /root/reference).  The reference stores no golden
vectors of its own.

The Euler-discrete scheduler / CFG loop restates Diffusers 0.21.2 (`requirements.txt:1`,
`EulerDiscreteScheduler`: scaled-linear betas 0.00085..0.012, 1000 train steps, "leading" spacing,
steps_offset 1, epsilon prediction), a third-party dependency absent from /root/reference with no
reference test pinning it: **parity unpinned** for the scheduler; the same loop is applied to the
oracle and to the engine, so the comparison is still like for like.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _sub(sd: SD, prefix: str) -> SD:
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def _has(sd: SD, prefix: str) -> bool:
    p = prefix + "."
    return any(k.startswith(p) for k in sd)


def _count(sd: SD, prefix: str) -> int:
    """Number of consecutive integer-indexed children under `prefix` (a ModuleList)."""
    n = 0
    while _has(sd, f"{prefix}.{n}"):
        n += 1
    return n


# ---- embeddings ------------------------------------------------------------------------------------
def timesteps_embedding(t: torch.Tensor, num_channels: int) -> torch.Tensor:
    """optimizers/unet_pt.py:22-36 -- cat([cos, sin]) of t * exp(-ln(1e4) * i / half)."""
    half = num_channels // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - 0.0)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def linear(sd: SD, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd["weight"], sd.get("bias"))


def timestep_mlp(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """optimizers/unet_pt.py:46-51 -- linear_1, SiLU, linear_2."""
    return linear(_sub(sd, "linear_2"), F.silu(linear(_sub(sd, "linear_1"), x)))


# ---- per-op oracles (the *pattern* side of the reference's fx rewrites) -----------------------------
def group_norm(x: torch.Tensor, groups: int, w, b, eps: float, silu: bool = False) -> torch.Tensor:
    """optimizers/replace_groupnorm.py:24-30 (GroupNorm) and :43-50 (SiLU(GroupNorm))."""
    y = F.group_norm(x, groups, w, b, eps)
    return F.silu(y) if silu else y


def layer_norm(x: torch.Tensor, w, b, eps: float = 1e-5) -> torch.Tensor:
    """optimizers/replace_layernorm.py:31-37; naive form kernels/layer_norm.py:28-38."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def geglu(state: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """optimizers/replace_geglu.py:34-35; exact-erf GELU as kernels/geglu.py:11-14."""
    return state * F.gelu(gate)


def attention_core(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, head_dim: int,
                   q_chunk: int = 1024) -> torch.Tensor:
    """optimizers/replace_attention.py:76-86 (pattern) == optimizers/unet_pt.py:133-142: per-head
    softmax(q k^T * head_dim^-0.5) v on (B, T, C) tensors.  Queries are processed in chunks so that the
    2048^2 configuration (T = 16384) does not materialise a 21 GB score tensor."""
    b, t, c = q.shape
    scale = head_dim ** -0.5
    qh = q.view(b, t, num_heads, head_dim).transpose(1, 2)
    kh = k.view(b, k.shape[1], num_heads, head_dim).transpose(1, 2)
    vh = v.view(b, v.shape[1], num_heads, head_dim).transpose(1, 2)
    outs = []
    for s in range(0, t, q_chunk):
        scores = torch.matmul(qh[:, :, s:s + q_chunk], kh.transpose(-2, -1)) * scale
        outs.append(torch.matmul(torch.softmax(scores, dim=-1), vh))
    o = torch.cat(outs, dim=2)
    return o.transpose(1, 2).contiguous().view(b, t, c)


# ---- blocks -----------------------------------------------------------------------------------------
def resnet_block(sd: SD, x: torch.Tensor, temb: torch.Tensor, groups: int) -> torch.Tensor:
    """optimizers/unet_pt.py:74-95."""
    h = group_norm(x, groups, sd["norm1.weight"], sd["norm1.bias"], 1e-5, silu=True)
    h = F.conv2d(h, sd["conv1.weight"], sd["conv1.bias"], padding=1)
    t = linear(_sub(sd, "time_emb_proj"), F.silu(temb))[:, :, None, None]
    h = h + t
    h = group_norm(h, groups, sd["norm2.weight"], sd["norm2.bias"], 1e-5, silu=True)
    h = F.conv2d(h, sd["conv2.weight"], sd["conv2.bias"], padding=1)
    if "conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd["conv_shortcut.weight"], sd["conv_shortcut.bias"])
    return x + h


def attention(sd: SD, x: torch.Tensor, ctx: Optional[torch.Tensor], head_dim: int) -> torch.Tensor:
    """optimizers/unet_pt.py:121-147."""
    src = x if ctx is None else ctx
    q = F.linear(x, sd["to_q.weight"])
    k = F.linear(src, sd["to_k.weight"])
    v = F.linear(src, sd["to_v.weight"])
    o = attention_core(q, k, v, q.shape[-1] // head_dim, head_dim)
    return F.linear(o, sd["to_out.0.weight"], sd["to_out.0.bias"])


def feed_forward(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """optimizers/unet_pt.py:150-176 -- GEGLU projection then output Linear."""
    state, gate = F.linear(x, sd["net.0.proj.weight"], sd["net.0.proj.bias"]).chunk(2, dim=-1)
    return F.linear(geglu(state, gate), sd["net.2.weight"], sd["net.2.bias"])


def transformer_block(sd: SD, x: torch.Tensor, ctx: torch.Tensor, head_dim: int) -> torch.Tensor:
    """optimizers/unet_pt.py:189-210."""
    x = attention(_sub(sd, "attn1"), layer_norm(x, sd["norm1.weight"], sd["norm1.bias"]), None, head_dim) + x
    x = attention(_sub(sd, "attn2"), layer_norm(x, sd["norm2.weight"], sd["norm2.bias"]), ctx, head_dim) + x
    x = feed_forward(_sub(sd, "ff"), layer_norm(x, sd["norm3.weight"], sd["norm3.bias"])) + x
    return x


def transformer_2d(sd: SD, x: torch.Tensor, ctx: torch.Tensor, groups: int, head_dim: int) -> torch.Tensor:
    """optimizers/unet_pt.py:223-243."""
    b, c, h, w = x.shape
    res = x
    y = group_norm(x, groups, sd["norm.weight"], sd["norm.bias"], 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
    y = linear(_sub(sd, "proj_in"), y)
    for i in range(_count(sd, "transformer_blocks")):
        y = transformer_block(_sub(sd, f"transformer_blocks.{i}"), y, ctx, head_dim)
    y = linear(_sub(sd, "proj_out"), y)
    y = y.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous()
    return y + res


def down_block(sd: SD, x, temb, ctx, groups: int, head_dim: int):
    """optimizers/unet_pt.py:280-289 (DownBlock2D) and :313-327 (CrossAttnDownBlock2D)."""
    outs = []
    has_attn = _has(sd, "attentions")
    for i in range(_count(sd, "resnets")):
        x = resnet_block(_sub(sd, f"resnets.{i}"), x, temb, groups)
        if has_attn:
            x = transformer_2d(_sub(sd, f"attentions.{i}"), x, ctx, groups, head_dim)
        outs.append(x)
    if _has(sd, "downsamplers"):
        x = F.conv2d(x, sd["downsamplers.0.conv.weight"], sd["downsamplers.0.conv.bias"], stride=2, padding=1)
        outs.append(x)  # unet_pt.py:249-254
    return x, outs


def up_block(sd: SD, x, skips: Sequence[torch.Tensor], temb, ctx, groups: int, head_dim: int):
    """optimizers/unet_pt.py:349-367 (CrossAttnUpBlock2D) and :381-388 (UpBlock2D)."""
    skips = list(skips)
    has_attn = _has(sd, "attentions")
    for i in range(_count(sd, "resnets")):
        x = torch.cat([x, skips.pop()], dim=1)
        x = resnet_block(_sub(sd, f"resnets.{i}"), x, temb, groups)
        if has_attn:
            x = transformer_2d(_sub(sd, f"attentions.{i}"), x, ctx, groups, head_dim)
    if _has(sd, "upsamplers"):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")  # unet_pt.py:264-266
        x = F.conv2d(x, sd["upsamplers.0.conv.weight"], sd["upsamplers.0.conv.bias"], padding=1)
    return x


def mid_block(sd: SD, x, temb, ctx, groups: int, head_dim: int):
    """optimizers/unet_pt.py:404-413."""
    x = resnet_block(_sub(sd, "resnets.0"), x, temb, groups)
    x = transformer_2d(_sub(sd, "attentions.0"), x, ctx, groups, head_dim)
    return resnet_block(_sub(sd, "resnets.1"), x, temb, groups)


# ---- whole UNet ---------------------------------------------------------------------------------------
@torch.no_grad()
def unet_forward(sd: SD, sample, timesteps, encoder_hidden_states, added_cond_kwargs, *, groups: int = 32,
                 head_dim: int = 64, addition_time_embed_dim: int = 256) -> List[torch.Tensor]:
    """optimizers/unet_pt.py:469-542, fp32.  Returns `[eps]` like the reference."""
    sd = {k: v.float() for k, v in sd.items()}
    sample = sample.float()
    ctx = encoder_hidden_states.float()
    base = sd["conv_in.weight"].shape[0]

    t = timesteps.to(sample.device).expand(sample.shape[0])  # device-agnostic: the GPU parity tests run this file in fp32 on cuda
    emb = timestep_mlp(_sub(sd, "time_embedding"), timesteps_embedding(t, base))
    text_embeds = added_cond_kwargs["text_embeds"].float()
    time_ids = added_cond_kwargs["time_ids"].float()
    time_embeds = timesteps_embedding(time_ids.flatten(), addition_time_embed_dim).reshape(text_embeds.shape[0], -1)
    emb = emb + timestep_mlp(_sub(sd, "add_embedding"), torch.cat([text_embeds, time_embeds], dim=-1))

    x = F.conv2d(sample, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    skips = [x]
    for i in range(_count(sd, "down_blocks")):
        x, outs = down_block(_sub(sd, f"down_blocks.{i}"), x, emb, ctx, groups, head_dim)
        skips += outs
    x = mid_block(_sub(sd, "mid_block"), x, emb, ctx, groups, head_dim)
    for i in range(_count(sd, "up_blocks")):
        bsd = _sub(sd, f"up_blocks.{i}")
        n = _count(bsd, "resnets")
        x = up_block(bsd, x, skips[-n:], emb, ctx, groups, head_dim)
        skips = skips[:-n]
    x = group_norm(x, groups, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], 1e-5, silu=True)
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    return [x]


# ---- Euler-discrete + classifier-free guidance (Diffusers 0.21.2 restated; parity unpinned) -----------
def euler_sigmas(num_inference_steps: int, num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                 beta_end: float = 0.012, steps_offset: int = 1):
    """Returns (timesteps fp32 [n], sigmas fp32 [n+1], init_noise_sigma)."""
    import numpy as np

    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float64) ** 2
    alphas_cumprod = np.cumprod(1.0 - betas)
    all_sigmas = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
    step_ratio = num_train_timesteps // num_inference_steps
    timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].astype(np.float64) + steps_offset
    sigmas = np.interp(timesteps, np.arange(num_train_timesteps), all_sigmas)
    sigmas = np.concatenate([sigmas, [0.0]])
    init_noise_sigma = float((sigmas.max() ** 2 + 1) ** 0.5)
    return (torch.tensor(timesteps, dtype=torch.float32), torch.tensor(sigmas, dtype=torch.float32), init_noise_sigma)


@torch.no_grad()
def denoise_loop(unet_fn, latents: torch.Tensor, cond: dict, uncond: dict, num_steps: int, guidance: float = 5.0):
    """The pipeline loop around the UNet: per step scale_model_input, UNet on the
    [uncond ; cond] pair, guidance mix, Euler step.  `unet_fn(sample, t, ctx, added)` -> [eps].
    latents: (P, C, H, W) unit-variance noise; cond/uncond: dicts with encoder_hidden_states,
    text_embeds, time_ids for P prompts.  Returns (final latents fp32, list of per-step eps)."""
    timesteps, sigmas, init_sigma = euler_sigmas(num_steps)
    x = latents.float() * init_sigma
    ctx = torch.cat([uncond["encoder_hidden_states"], cond["encoder_hidden_states"]], dim=0)
    added = {
        "text_embeds": torch.cat([uncond["text_embeds"], cond["text_embeds"]], dim=0),
        "time_ids": torch.cat([uncond["time_ids"], cond["time_ids"]], dim=0),
    }
    eps_trace = []
    for i in range(num_steps):
        sigma, sigma_next = float(sigmas[i]), float(sigmas[i + 1])
        model_in = x / math.sqrt(sigma * sigma + 1.0)
        eps = unet_fn(torch.cat([model_in, model_in], dim=0), timesteps[i], ctx, added)[0].float()
        eps_u, eps_c = eps.chunk(2, dim=0)
        eps = eps_u + guidance * (eps_c - eps_u)
        eps_trace.append(eps)
        x = x + (sigma_next - sigma) * eps
    return x, eps_trace
