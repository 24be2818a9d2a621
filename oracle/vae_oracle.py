"""ORACLE -- test infrastructure only.  fp32 restatement of the VAE decode path as pure
functions over a Diffusers-keyed `AutoencoderKL` state dict (`post_quant_conv.*`, `decoder.*`).

Only `tests/` and tools that CHECK the product may import this module; `stabletriton_b200/` never does.

What it restates: Diffusers' `AutoencoderKL.decode(latents / scaling_factor).sample` with the SDXL VAE configuration
(`Decoder`: conv_in -> UNetMidBlock2D [resnet, single-head attention of width C, resnet] -> UpDecoderBlock2D x 4
[layers_per_block + 1 resnets, nearest-2x + conv] -> GroupNorm + SiLU + conv_out; resnets without time embedding,
eps 1e-6, 32 groups).  Diffusers is a third-party dependency that is absent from /root/reference, and the reference
itself ships no VAE code at all (README.md:5 lists it as planned; implementations/Diffusers/load_sdxl_pipeline.py:39,46
runs the pipeline's own eager VAE): there is nothing in the reference to pin this against -- **parity unpinned**.  It is
written independently of `stabletriton_b200/vae.py` (functional, structure discovered from the key names) and
`tests/test_vae.py` checks the two against each other in fp32 on the CPU before either is used to judge the kernels.

`attention(..., query_chunk=n)` evaluates the softmax in blocks of query rows, so that the 16 384-token attention of a
1024^2 decode does not materialise more than n x T scores at a time (same arithmetic, row by row).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _sub(sd: SD, prefix: str) -> SD:
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def _count(sd: SD, prefix: str) -> int:
    p = prefix + "."
    return len({k[len(p):].split(".", 1)[0] for k in sd if k.startswith(p)})


def group_norm(x, sd: SD, groups: int, eps: float, silu: bool):
    y = F.group_norm(x, groups, sd["weight"], sd["bias"], eps)
    return F.silu(y) if silu else y


def conv(x, sd: SD, padding: int):
    return F.conv2d(x, sd["weight"], sd["bias"], padding=padding)


def resnet_block(sd: SD, x, groups: int, eps: float):
    """Diffusers ResnetBlock2D with temb_channels=None: norm1 -> SiLU -> conv1 -> norm2 -> SiLU -> conv2, + shortcut."""
    h = conv(group_norm(x, _sub(sd, "norm1"), groups, eps, True), _sub(sd, "conv1"), 1)
    h = conv(group_norm(h, _sub(sd, "norm2"), groups, eps, True), _sub(sd, "conv2"), 1)
    sc = conv(x, _sub(sd, "conv_shortcut"), 0) if "conv_shortcut.weight" in sd else x
    return sc + h


def attention(sd: SD, x, groups: int, eps: float, query_chunk: int = 0):
    """Diffusers Attention(heads=1, dim_head=C, residual_connection=True): softmax(q k^T / sqrt(C)) v over H*W tokens."""
    b, c, h, w = x.shape
    t = group_norm(x, _sub(sd, "group_norm"), groups, eps, False).reshape(b, c, h * w).transpose(1, 2)
    q = F.linear(t, sd["to_q.weight"], sd["to_q.bias"])
    k = F.linear(t, sd["to_k.weight"], sd["to_k.bias"])
    v = F.linear(t, sd["to_v.weight"], sd["to_v.bias"])
    scale = 1.0 / math.sqrt(c)
    n = h * w
    step = query_chunk if query_chunk > 0 else n
    rows = []
    for r0 in range(0, n, step):
        p = torch.softmax(torch.matmul(q[:, r0:r0 + step], k.transpose(1, 2)) * scale, dim=-1)
        rows.append(torch.matmul(p, v))
    o = F.linear(torch.cat(rows, dim=1), sd["to_out.0.weight"], sd["to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(b, c, h, w)


def vae_decode(sd: SD, latents, *, groups: int = 32, eps: float = 1e-6, scaling_factor: float = 0.13025,
               query_chunk: int = 0):
    """latents (B, 4, h, w) -> image (B, 3, 8h, 8w) for the 4-level SDXL decoder (2^(levels-1) upsampling in general)."""
    z = conv(latents / scaling_factor, _sub(sd, "post_quant_conv"), 0)
    d = _sub(sd, "decoder")
    x = conv(z, _sub(d, "conv_in"), 1)
    x = resnet_block(_sub(d, "mid_block.resnets.0"), x, groups, eps)
    x = attention(_sub(d, "mid_block.attentions.0"), x, groups, eps, query_chunk)
    x = resnet_block(_sub(d, "mid_block.resnets.1"), x, groups, eps)
    for i in range(_count(d, "up_blocks")):
        blk = _sub(d, f"up_blocks.{i}")
        for j in range(_count(blk, "resnets")):
            x = resnet_block(_sub(blk, f"resnets.{j}"), x, groups, eps)
        if "upsamplers.0.conv.weight" in blk:
            x = conv(F.interpolate(x, scale_factor=2.0, mode="nearest"), _sub(blk, "upsamplers.0.conv"), 1)
    x = group_norm(x, _sub(d, "conv_norm_out"), groups, eps, True)
    return conv(x, _sub(d, "conv_out"), 1)
