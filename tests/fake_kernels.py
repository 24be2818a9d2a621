"""Test double for `stabletriton_b200.kernels`: the same tensor-level API implemented with eager
PyTorch, so the fx graph surgery can be validated on the CPU (no GPU in the build container).
Test infrastructure only -- the product never imports this; on a GPU box the real kernels run.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

import stabletriton_b200.kernels as K

CALLS: dict = {}


def _count(name):
    CALLS[name] = CALLS.get(name, 0) + 1


def _partials(rows: torch.Tensor, rows_per_image: int):
    """What the GEMM epilogue emits with gn_partial: (mean, M2) per 128-row tile and column, or None if not eligible."""
    m, c = rows.shape
    if K.gn_partial_rows(m, rows_per_image) == 0:
        return None
    t = rows.double().reshape(m // 128, 128, c)
    mean = t.mean(dim=1)
    m2 = ((t - mean[:, None, :]) ** 2).sum(dim=1)
    return torch.stack([mean, m2], dim=-1).float().contiguous()


def groupnorm_wrapper(input, num_groups, weight, bias, eps, activation=False, partials=None):
    _count("groupnorm")
    if partials is not None and all(p is not None for p in partials):
        # the statistics handed over by the producers must be THE statistics of `input` (right producer, right channel
        # order of a concatenation): rebuild the group moments from them as gn_finalize_kernel does and compare
        _count("groupnorm_from_partials")
        n, c, h, w = input.shape
        part = torch.cat(list(partials), dim=1).double()  # [tiles, C, 2]
        assert part.shape == (n * h * w // 128, c, 2), (part.shape, input.shape)
        tiles = h * w // 128
        part = part.reshape(n, tiles, num_groups, c // num_groups, 2)
        mean = part[..., 0].mean(dim=(1, 3))
        m2 = (part[..., 1] + 128.0 * (part[..., 0] - mean[:, None, :, None]) ** 2).sum(dim=(1, 3))
        var = m2 / (128.0 * tiles * (c // num_groups))
        xg = input.double().reshape(n, num_groups, -1)
        assert torch.allclose(mean, xg.mean(dim=2), rtol=1e-4, atol=1e-5), "partials: wrong mean"
        assert torch.allclose(var, xg.var(dim=2, unbiased=False), rtol=1e-3, atol=1e-6), "partials: wrong variance"
    y = F.group_norm(input, num_groups, weight, bias, eps)
    return F.silu(y) if activation else y


def layer_norm(x, weight, bias, eps):
    _count("layer_norm")
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


def linear(x, weight, bias=None, activation=False, residual=None, geglu=False, silu_input=False, block_n=0,
           w_static=False, gn_stats=0, out=None):
    _count("linear_geglu" if geglu else "linear")
    if out is not None:
        out.copy_(linear(x, weight, bias, activation, residual, geglu, silu_input))
        return out
    if silu_input:
        x = F.silu(x)
    y = F.linear(x, weight, bias)
    if geglu:
        s, g = y.chunk(2, dim=-1)
        y = s * F.gelu(g)
    if activation:
        y = F.silu(y)
    if residual is not None:
        assert residual.shape == y.shape, (residual.shape, y.shape)
        y = y + residual
    if gn_stats:
        return y, _partials(y.reshape(-1, y.shape[-1]), gn_stats)
    return y


def geglu_wrapper(state, gate):
    _count("geglu")
    return state * F.gelu(gate)


def attention_btc(q, k, v, num_heads, sm_scale):
    _count("attention")
    b, t, c = q.shape
    d = c // num_heads
    qh = q.reshape(b, t, num_heads, d).transpose(1, 2)
    kh = k.reshape(b, k.shape[1], num_heads, d).transpose(1, 2)
    vh = v.reshape(b, v.shape[1], num_heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-2, -1) * sm_scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(b, t, c)


def upsample_nearest2x(x):
    _count("upsample")
    return F.interpolate(x, scale_factor=2.0, mode="nearest")


def conv2d(x, weight, bias, stride=1, padding=1, temb=None, residual=None, nchw_output=False, block_n=0,
           w_static=False, gn_stats=False):
    _count("conv2d")
    y = F.conv2d(x, weight, bias, stride=stride, padding=padding)
    if temb is not None:
        assert temb.shape == y.shape[:2]
        y = y + temb[:, :, None, None]
    if residual is not None:
        assert residual.shape == y.shape
        y = y + residual
    if gn_stats:
        return y, _partials(y.permute(0, 2, 3, 1).reshape(-1, y.shape[1]), y.shape[2] * y.shape[3])
    return y


def matmul_nt_f32(a, b):
    _count("matmul_nt_f32")
    return a.float() @ b.float().t()


def softmax_rows(scores, scale, out=None):
    _count("softmax_rows")
    assert scores.dtype == torch.float32
    p = torch.softmax(scores * scale, dim=-1)
    return p if out is None else out.copy_(p)


def transpose_tokens(x):
    _count("transpose_tokens")
    return x.transpose(1, 2).contiguous()


def pointwise_conv_small(x, weight, bias, in_scale=1.0):
    _count("pointwise_conv_small")
    return F.conv2d(x * in_scale, weight, bias)


def concat_channels(a, b):
    _count("concat")
    return torch.cat([a, b], dim=1)


def timestep_embedding(t, num_channels):
    _count("timestep")
    import math
    half = num_channels // 2
    f = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    e = t.reshape(-1)[:, None].float() * f[None, :]
    return torch.cat([torch.cos(e), torch.sin(e)], dim=-1)


_NAMES = ["groupnorm_wrapper", "layer_norm", "linear", "geglu_wrapper", "attention_btc", "upsample_nearest2x",
          "conv2d", "concat_channels", "timestep_embedding", "matmul_nt_f32", "softmax_rows", "transpose_tokens",
          "pointwise_conv_small"]


@contextlib.contextmanager
def installed():
    """Temporarily route `stabletriton_b200.kernels.*` to the eager implementations above."""
    saved = {n: getattr(K, n) for n in _NAMES}
    CALLS.clear()
    try:
        for n in _NAMES:
            setattr(K, n, globals()[n])
        yield CALLS
    finally:
        for n, f in saved.items():
            setattr(K, n, f)
