"""The op seam: module-level functions that the rewritten fx graph calls.

Same mechanism as the reference: each function is registered with `torch.fx.wrap`, so it survives
re-tracing as a leaf `call_function` node, and it receives the *original* matched submodule so
weights are read live from the model.  Reference counterparts:

    group_norm_wrapper(v, groupnorm, activation)              optimizers/replace_groupnorm.py:18-21
    layer_norm_wrapper(v, layernorm)                          optimizers/replace_layernorm.py:17-27
    linear_wrapper(v, linear, activation)                     optimizers/replace_linear.py:20-24
    linear_wrapper_functional(v, weight, bias, activation)    optimizers/replace_linear.py:26-37
    geglu_wrapper(state, gate)          [geglu_triton]        optimizers/replace_geglu.py:23-30
    attention_wrapper(q, k, v, output, sm_scale, num_heads, head_dim)   optimizers/replace_attention.py:60-71

New seams the reference lacks (it left Linear / Conv2d / glue to cuBLAS / cuDNN / eager):
`linear_geglu_wrapper`, `conv2d_wrapper`, `concat_wrapper`, `timestep_wrapper`.  No wrapper casts or
mutates parameters (the reference's fp16 "hacks", replace_layernorm.py:19-22, are not reproduced).

Every wrapper here hands module *parameters* to the kernels, so it declares them static (`w_static=True`: the kernel may
fetch the weight ahead of its programmatic dependency on the preceding launch).  The tensor-level API in kernels.py
defaults to `w_static=False`, where a weight may be the output of the kernel launched just before.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.fx

from . import kernels as K


def group_norm_wrapper(v: torch.Tensor, groupnorm: torch.nn.GroupNorm, activation: bool,
                       partials: Optional[tuple] = None) -> torch.Tensor:
    """partials (new): the statistics the producer(s) of `v` emitted from their GEMM epilogues -- `(part,)`, or
    `(part_a, part_b)` when v = cat([a, b], 1) -- see linear_stats_wrapper / conv2d_stats_wrapper.  Any entry may be
    None (producer shape not eligible): the kernel then runs its own statistics pass, as without the argument."""
    return K.groupnorm_wrapper(v, groupnorm.num_groups, groupnorm.weight, groupnorm.bias, groupnorm.eps, activation,
                               partials=partials)


def layer_norm_wrapper(v: torch.Tensor, layernorm: torch.nn.LayerNorm) -> torch.Tensor:
    return K.layer_norm(v, layernorm.weight, layernorm.bias, layernorm.eps)


def linear_wrapper(v: torch.Tensor, linear: torch.nn.Linear, activation: bool,
                   residual: Optional[torch.Tensor] = None, silu_input: bool = False) -> torch.Tensor:
    return K.linear(v, linear.weight, linear.bias, activation=activation, residual=residual, silu_input=silu_input,
                    w_static=True)


def linear_stats_wrapper(v: torch.Tensor, linear: torch.nn.Linear, activation: bool,
                         residual: Optional[torch.Tensor] = None, rows_per_image: int = 0):
    """linear_wrapper for a result that feeds a GroupNorm (the transformer's proj_out + image residual,
    unet_pt.py:236-243 -> the next resnet's norm1): returns (y, partials).  rows_per_image = H*W of the feature map the
    token rows belong to; 0 = take it from v (B, H*W, C)."""
    return K.linear(v, linear.weight, linear.bias, activation=activation, residual=residual, w_static=True,
                    gn_stats=rows_per_image or (v.shape[-2] if v.dim() >= 3 else v.shape[0]))


def linear_wrapper_functional(v: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
                              activation: bool, silu_input: bool = False) -> torch.Tensor:
    # only compile()'s own fusion passes emit this call, always on (fused) parameter buffers
    return K.linear(v, weight, bias, activation=activation, silu_input=silu_input, w_static=True)


def linear_geglu_wrapper(v: torch.Tensor, linear: torch.nn.Linear) -> torch.Tensor:
    """GEGLU projection with the gate fused into the GEMM epilogue (unet_pt.py:155-158)."""
    return K.linear(v, linear.weight, linear.bias, geglu=True, w_static=True)


def geglu_wrapper(state: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    return K.geglu_wrapper(state, gate)


def attention_wrapper(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, output: Optional[torch.Tensor],
                      sm_scale: float, num_heads: int, head_dim: int) -> torch.Tensor:
    """q/k/v arrive un-split as (B, T, H*D) exactly as in the reference's rewritten graph; unlike the
    reference the heads ARE separated.  `output` is accepted for signature parity."""
    if head_dim != 64:
        raise ValueError(f"attention_wrapper: head_dim must be 64 (got {head_dim})")
    return K.attention_btc(q, k, v, num_heads, sm_scale)


def conv2d_wrapper(v: torch.Tensor, conv: torch.nn.Conv2d, temb: Optional[torch.Tensor] = None,
                   residual: Optional[torch.Tensor] = None, upsample: bool = False) -> torch.Tensor:
    """Conv2d with optional fused epilogues: + temb[:, :, None, None] (unet_pt.py:82-83), + residual
    (unet_pt.py:93); upsample=True applies nearest-2x first (unet_pt.py:265-266)."""
    if conv.groups != 1 or conv.dilation != (1, 1) or conv.stride[0] != conv.stride[1] \
            or conv.padding[0] != conv.padding[1] or isinstance(conv.padding, str):
        raise ValueError(f"conv2d_wrapper: unsupported convolution {conv}")
    if upsample:
        v = K.upsample_nearest2x(v)
    return K.conv2d(v, conv.weight, conv.bias, stride=conv.stride[0], padding=conv.padding[0], temb=temb,
                    residual=residual, nchw_output=conv.out_channels <= 8, w_static=True)


def conv2d_stats_wrapper(v: torch.Tensor, conv: torch.nn.Conv2d, temb: Optional[torch.Tensor] = None,
                         residual: Optional[torch.Tensor] = None, upsample: bool = False):
    """conv2d_wrapper for a result that feeds a GroupNorm: returns (y, partials) -- the implicit-GEMM epilogue also
    writes per-tile column statistics of what it stores, so the GroupNorm needs no statistics pass (reference: the
    GroupNorm kernel re-reads the whole activation, kernels/groupnorm.py:24-119)."""
    if conv.groups != 1 or conv.dilation != (1, 1) or conv.stride[0] != conv.stride[1] \
            or conv.padding[0] != conv.padding[1] or isinstance(conv.padding, str):
        raise ValueError(f"conv2d_wrapper: unsupported convolution {conv}")
    if upsample:
        v = K.upsample_nearest2x(v)
    return K.conv2d(v, conv.weight, conv.bias, stride=conv.stride[0], padding=conv.padding[0], temb=temb,
                    residual=residual, nchw_output=False, w_static=True, gn_stats=True)


def concat_wrapper(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return K.concat_channels(a, b)


def timestep_wrapper(t: torch.Tensor, num_channels: int) -> torch.Tensor:
    return K.timestep_embedding(t, num_channels)


for _name in ("group_norm_wrapper", "layer_norm_wrapper", "linear_wrapper", "linear_wrapper_functional",
              "linear_geglu_wrapper", "geglu_wrapper", "attention_wrapper", "conv2d_wrapper", "concat_wrapper",
              "timestep_wrapper", "linear_stats_wrapper", "conv2d_stats_wrapper"):
    torch.fx.wrap(_name)
