"""Kernel-level Python API: tensors in, tensors out, one C-ABI call each.

Mirrors the launcher functions of the reference's `stabletriton.kernels` package (same names and
argument meaning where the reference has one):

    groupnorm_wrapper   kernels/groupnorm.py:128-161
    layer_norm          kernels/layer_norm.py:338-346
    sdxl_forward        kernels/linear.py:173-222
    geglu_wrapper       kernels/geglu.py:28-35
    attention           kernels/attention_fa2.py:113-140
    implicit_gemm_fprop kernels/Conv_Kernels/conv_implicit_gemm.py:143-182

Like the reference launchers they allocate their outputs with torch, never synchronise and never read
device memory on the host, so they can be captured into a CUDA graph.  Unlike the reference they do
not cast or mutate their inputs: everything is bf16 on a CUDA device, validated up front
(`ValueError`/`TypeError` before any launch).  Activations are channels-last: a 4-D tensor has
logical shape (N, C, H, W) and NHWC strides; a token tensor (B, T, C) shares the same memory.
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch

from . import _cabi
from ._cabi import ST_EPI_F32OUT, ST_EPI_GEGLU, ST_EPI_SILU, ST_W_STATIC, check, lib

BF16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_bf16_cuda(name: str, *tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
        if t.device.type != "cuda":
            raise ValueError(f"{name}: tensors must live on a CUDA device (got {t.device}); there is no CPU path")
        if t.dtype != BF16:
            raise TypeError(f"{name}: tensors must be bfloat16 (got {t.dtype})")


_WORKSPACE: dict = {}  # device index -> fp32 scratch registered with the library (stream-K partial tiles)


def _ensure_workspace(device: torch.device) -> None:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _WORKSPACE:
        L = _cabi._load()
        with torch.cuda.device(idx):
            buf = torch.empty(L.st_workspace_bytes(), dtype=torch.uint8, device=device)
            check(L.st_set_workspace(buf.data_ptr(), buf.numel()), "set_workspace")
        _WORKSPACE[idx] = buf


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """Return x (logical NCHW) with dense NHWC strides, converting only if it has to."""
    if x.dim() != 4:
        raise ValueError(f"expected a 4-D (N, C, H, W) tensor, got shape {tuple(x.shape)}")
    n, c, h, w = x.shape
    if x.stride() == (h * w * c, 1, w * c, c):
        return x
    return x.contiguous(memory_format=torch.channels_last) if (c > 1 and h * w > 1) else \
        x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def _empty_nhwc(n: int, c: int, h: int, w: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty((n, h, w, c), dtype=BF16, device=like.device).permute(0, 3, 1, 2)


def _rows(x: torch.Tensor) -> tuple[torch.Tensor, int, int]:
    """View (..., C) as M rows of C contiguous elements with a uniform pitch; returns (tensor, M, ld)."""
    if x.dim() == 0:
        raise ValueError("expected a tensor with at least one dimension")
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if x.stride(-1) != 1:
        x = x.contiguous()
    c = x.shape[-1]
    m = x.numel() // max(c, 1)
    ld, expect, ok = None, None, True
    for d in range(x.dim() - 2, -1, -1):  # innermost leading dim first
        if x.shape[d] == 1:
            continue
        if ld is None:
            ld = x.stride(d)
        elif x.stride(d) != expect:
            ok = False
            break
        expect = x.stride(d) * x.shape[d]
    if ld is None:
        ld = c
    if not ok or ld < c:
        x = x.contiguous()
        ld = c
    return x, m, ld


# ------------------------------------------------------------------------------------------------
# GroupNorm (+SiLU)
# ------------------------------------------------------------------------------------------------
def gn_partial_rows(rows: int, rows_per_image: int) -> int:
    """Number of 128-row tiles of a producer whose output has `rows` rows (`rows_per_image` per image), or 0 if the
    producer cannot emit GroupNorm partials for it (a tile must lie inside one image)."""
    return rows // 128 if (rows % 128 == 0 and rows_per_image % 128 == 0) else 0


def groupnorm_wrapper(input: torch.Tensor, num_groups: int, weight: Optional[torch.Tensor],
                      bias: Optional[torch.Tensor], eps: float, activation: bool = False,
                      partials: Optional[tuple] = None) -> torch.Tensor:
    """torch.nn.GroupNorm(num_groups, C, eps) [+ SiLU] on a 4-D tensor (reference: groupnorm.py:128).

    partials: (part_a,) or (part_a, part_b) -- the per-tile column statistics emitted by the producer(s) of `input`
    (`linear(..., gn_stats=True)` / `conv2d(..., gn_stats=True)`; two of them when `input` is a channel concatenation).
    With them the activation is read once; without them a statistics pass runs first."""
    _require_bf16_cuda("groupnorm_wrapper", input, weight, bias)
    x = _nhwc(input)
    n, c, h, w = x.shape
    if c % num_groups != 0:
        raise ValueError(f"groupnorm_wrapper: C={c} is not divisible by num_groups={num_groups}")
    out = _empty_nhwc(n, c, h, w, x)
    L = lib()
    ws_bytes = L.st_groupnorm_workspace_bytes(n, h * w, c, num_groups)
    if ws_bytes == 0:
        raise ValueError(f"groupnorm_wrapper: unsupported shape N={n} C={c} HW={h * w} groups={num_groups}")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    if partials is not None and all(p is not None for p in partials):
        tiles = gn_partial_rows(n * h * w, h * w)
        pa = partials[0]
        pb = partials[1] if len(partials) > 1 else None
        ca = pa.shape[1]
        cb = pb.shape[1] if pb is not None else 0
        for t in (pa, pb):
            if t is not None and (t.dtype != torch.float32 or t.dim() != 3 or t.shape[0] != tiles or t.shape[2] != 2
                                  or not t.is_contiguous() or t.device != x.device):
                raise ValueError(f"groupnorm_wrapper: partial statistics must be contiguous fp32 [{tiles}, C, 2]")
        if tiles == 0 or ca + cb != c:
            raise ValueError(f"groupnorm_wrapper: partials cover {ca}+{cb} channels / {tiles} tiles, input has C={c}")
        check(L.st_groupnorm_from_partials_nhwc_bf16(
            x.data_ptr(), out.data_ptr(), _ptr(weight), _ptr(bias), ws.data_ptr(), n, h * w, c, num_groups, float(eps),
            int(bool(activation)), pa.data_ptr(), ca, _ptr(pb), cb, _stream(x)), "groupnorm_from_partials")
        return out
    check(L.st_groupnorm_nhwc_bf16(x.data_ptr(), out.data_ptr(), _ptr(weight), _ptr(bias), ws.data_ptr(), n, h * w, c,
                                   num_groups, float(eps), int(bool(activation)), _stream(x)), "groupnorm")
    return out


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """LayerNorm over the last dimension (reference: layer_norm.py:338-346)."""
    _require_bf16_cuda("layer_norm", x, weight, bias)
    xr, m, ld = _rows(x)
    n = x.shape[-1]
    if weight.numel() != n:
        raise ValueError(f"layer_norm: weight has {weight.numel()} elements, expected {n}")
    out = torch.empty(x.shape, dtype=BF16, device=x.device)
    check(lib().st_layernorm_bf16(xr.data_ptr(), ld, out.data_ptr(), n, weight.data_ptr(), _ptr(bias), m, n,
                                  float(eps), _stream(x)), "layernorm")
    return out


# ------------------------------------------------------------------------------------------------
# Linear family
# ------------------------------------------------------------------------------------------------
SMALL_M = 32


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, activation: bool = False,
           residual: Optional[torch.Tensor] = None, geglu: bool = False, silu_input: bool = False,
           block_n: int = 0, w_static: bool = False, gn_stats: int = 0, out: Optional[torch.Tensor] = None):
    """y = epi(x @ weight.T + bias) [+ residual]; weight is (N, K) as in nn.Linear.

    out: optional preallocated bf16 result (rows of N contiguous elements with a uniform pitch, e.g. a row slice of a
    larger buffer); not on the tiny-M path.

    gn_stats = rows per image (> 0): y feeds a GroupNorm -- also return the per-tile column statistics the GEMM epilogue
    can emit for free, as `(y, partials)`; partials is None when the shape does not allow it (see gn_partial_rows).

    activation: SiLU epilogue.  geglu: weight rows are [state ; gate], y = state * gelu(gate) with N/2
    columns.  residual (same shape as y) is added in fp32 before the single bf16 rounding.
    silu_input applies SiLU to x on load (tiny-M path only: the resnet time-embedding projection).
    w_static: the caller vouches that `weight` is a parameter -- not produced by the kernel launched just before this
    one -- so the kernel may fetch it ahead of its programmatic (PDL) dependency (ST_W_STATIC).  The op seam
    (wrappers.py) sets it for module parameters; it is dropped again if this call has to copy the weight first.
    """
    _require_bf16_cuda("linear", x, weight, bias, residual)
    _ensure_workspace(x.device)
    if weight.dim() != 2:
        raise ValueError("linear: weight must be 2-D (N, K)")
    n_rows, k = weight.shape
    if x.shape[-1] != k:
        raise ValueError(f"linear: x has {x.shape[-1]} features, weight expects {k}")
    if weight.stride(1) != 1:
        weight = weight.contiguous()
        w_static = False  # the copy is produced on this stream, right before the launch
    wflag = ST_W_STATIC if w_static else 0
    xr, m, lda = _rows(x)
    n_out = n_rows // 2 if geglu else n_rows
    ldd = n_out
    if out is None:
        out = torch.empty(x.shape[:-1] + (n_out,), dtype=BF16, device=x.device)
    else:
        _require_bf16_cuda("linear", out)
        if out.shape[-1] != n_out or out.numel() != m * n_out or out.stride(-1) != 1:
            raise ValueError(f"linear: out must hold {m} rows of {n_out} contiguous elements, got {tuple(out.shape)}")
        orows, _, ldd = _rows(out)
        if orows.data_ptr() != out.data_ptr():
            raise ValueError("linear: out rows must have a uniform pitch")
    L = lib()
    if m <= SMALL_M and not geglu and residual is None and ldd == n_out:
        check(L.st_linear_small_m_bf16(xr.data_ptr(), lda, weight.data_ptr(), weight.stride(0), _ptr(bias),
                                       out.data_ptr(), n_out, m, n_rows, k, int(silu_input), int(activation), wflag,
                                       _stream(x)), "linear_small_m")
        return (out, None) if gn_stats else out
    if silu_input:
        raise ValueError("linear: silu_input is only supported on the tiny-M path (M <= 16)")
    res_ptr, ldr = 0, 0
    if residual is not None:
        if residual.shape != out.shape:
            raise ValueError(f"linear: residual shape {tuple(residual.shape)} != output shape {tuple(out.shape)}")
        rr, _, ldr = _rows(residual)
        res_ptr = rr.data_ptr()
        residual = rr  # keep alive
    flags = (ST_EPI_SILU if activation else 0) | (ST_EPI_GEGLU if geglu else 0) | wflag
    part = None
    if gn_stats and not geglu and gn_partial_rows(m, gn_stats):
        part = torch.empty((m // 128, n_out, 2), dtype=torch.float32, device=x.device)
    check(L.st_gemm_bf16(xr.data_ptr(), lda, weight.data_ptr(), weight.stride(0), out.data_ptr(), ldd, m, n_rows, k,
                         _ptr(bias), res_ptr, ldr, flags, block_n, _ptr(part), _stream(x)), "gemm")
    return (out, part) if gn_stats else out


def matmul_nt_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """fp32 (M, N) = a (M, K) @ b (N, K)^T on the tensor-core GEMM, fp32 result (ST_EPI_F32OUT): attention scores of a
    head wider than the flash kernels cover (the VAE mid block: one head of 512), ahead of `softmax_rows`."""
    _require_bf16_cuda("matmul_nt_f32", a, b)
    _ensure_workspace(a.device)
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"matmul_nt_f32: expected (M, K) and (N, K), got {tuple(a.shape)} and {tuple(b.shape)}")
    ar, m, lda = _rows(a)
    br, n, ldb = _rows(b)
    out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    check(lib().st_gemm_bf16(ar.data_ptr(), lda, br.data_ptr(), ldb, out.data_ptr(), n, m, n, a.shape[1], 0, 0, 0,
                             ST_EPI_F32OUT, 0, 0, _stream(a)), "gemm_f32out")
    return out


def softmax_rows(scores: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 softmax(scale * scores) over the last dimension of an fp32 (M, N) matrix."""
    if scores.dtype != torch.float32 or scores.device.type != "cuda" or scores.dim() != 2 or scores.stride(1) != 1:
        raise ValueError("softmax_rows: expected a 2-D fp32 CUDA tensor with contiguous rows")
    m, n = scores.shape
    if out is None:
        out = torch.empty((m, n), dtype=BF16, device=scores.device)
    check(lib().st_softmax_rows_f32_bf16(scores.data_ptr(), scores.stride(0), out.data_ptr(), out.stride(0), m, n,
                                         float(scale), _stream(scores)), "softmax_rows")
    return out


def transpose_tokens(x: torch.Tensor) -> torch.Tensor:
    """(B, T, C) -> dense (B, C, T): the B operand V^T of P @ V when P @ V runs on the plain GEMM."""
    _require_bf16_cuda("transpose_tokens", x)
    if x.dim() != 3 or x.stride(2) != 1 or x.stride(0) != x.shape[1] * x.stride(1):
        x = x.contiguous()
    b, t, c = x.shape
    out = torch.empty((b, c, t), dtype=BF16, device=x.device)
    check(lib().st_nhwc_to_nchw_bf16(x.data_ptr(), x.stride(1), out.data_ptr(), b, t, c, _stream(x)), "transpose_tokens")
    return out


def pointwise_conv_small(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], in_scale: float = 1.0) -> torch.Tensor:
    """1x1 convolution between at most 8 channels on a dense NCHW tensor, input pre-scaled (the VAE's post_quant_conv)."""
    _require_bf16_cuda("pointwise_conv_small", x, weight, bias)
    x = x.contiguous()
    n, ci, h, w = x.shape
    co = weight.shape[0]
    wm = weight.reshape(co, ci).contiguous()
    out = torch.empty((n, co, h, w), dtype=BF16, device=x.device)
    check(lib().st_pointwise_conv_small_bf16(x.data_ptr(), wm.data_ptr(), _ptr(bias), out.data_ptr(), n, h * w, ci, co,
                                             float(in_scale), _stream(x)), "pointwise_conv_small")
    return out


def sdxl_forward(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], activation: bool) -> torch.Tensor:
    """Reference-compatible name and signature (kernels/linear.py:173): act(x . W^T + b), act = SiLU."""
    return linear(x, weight, bias, activation=bool(activation))


def geglu_wrapper(state: torch.Tensor, gate: torch.Tensor) -> torch.Tensor:
    """state * gelu_erf(gate), elementwise (reference: geglu.py:28-35).  Accepts the strided halves of
    `proj(x).chunk(2, -1)` directly -- no `.contiguous()` copies as in replace_geglu.py:38."""
    _require_bf16_cuda("geglu_wrapper", state, gate)
    if state.shape != gate.shape:
        raise ValueError("geglu_wrapper: state and gate must have the same shape")
    sr, m, lds = _rows(state)
    gr, _, ldg = _rows(gate)
    c = state.shape[-1]
    out = torch.empty(state.shape, dtype=BF16, device=state.device)
    check(lib().st_geglu_bf16(sr.data_ptr(), lds, gr.data_ptr(), ldg, out.data_ptr(), c, m, c, _stream(state)), "geglu")
    return out


# ------------------------------------------------------------------------------------------------
# Attention
# ------------------------------------------------------------------------------------------------
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, sm_scale: float) -> torch.Tensor:
    """softmax(q k^T * sm_scale) v for (B, H, T, D=64) tensors (reference: attention_fa2.py:113-140)."""
    _require_bf16_cuda("attention", q, k, v)
    if q.dim() != 4 or q.shape[-1] != 64:
        raise ValueError(f"attention: expected (B, H, T, 64) tensors, got {tuple(q.shape)}")
    q, k, v = (t if t.stride(-1) == 1 else t.contiguous() for t in (q, k, v))
    b, h, tq, _ = q.shape
    tk = k.shape[2]
    o = torch.empty((b, h, tq, 64), dtype=BF16, device=q.device)
    check(lib().st_attention_bf16(q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
                                  k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
                                  v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
                                  o.data_ptr(), o.stride(0), o.stride(1), o.stride(2),
                                  b, h, tq, tk, float(sm_scale), _stream(q)), "attention")
    return o


def attention_btc(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, num_heads: int, sm_scale: float) -> torch.Tensor:
    """Multi-head attention on un-split (B, T, H*64) tensors -- the layout the fx pattern hands over
    (replace_attention.py:76-86 receives q/k/v before the head split).  q/k/v may be column slices of
    one fused projection buffer.  Returns (B, Tq, H*64)."""
    _require_bf16_cuda("attention", q, k, v)
    if q.dim() != 3 or q.shape[-1] != num_heads * 64:
        raise ValueError(f"attention: expected (B, T, {num_heads}*64) tensors, got {tuple(q.shape)}")
    q, k, v = (t if t.stride(-1) == 1 else t.contiguous() for t in (q, k, v))
    b, tq, c = q.shape
    tk = k.shape[1]
    if k.shape != (b, tk, c) or v.shape != (b, tk, c):
        raise ValueError("attention: k/v shape mismatch")
    o = torch.empty((b, tq, c), dtype=BF16, device=q.device)
    check(lib().st_attention_bf16(q.data_ptr(), q.stride(0), 64, q.stride(1),
                                  k.data_ptr(), k.stride(0), 64, k.stride(1),
                                  v.data_ptr(), v.stride(0), 64, v.stride(1),
                                  o.data_ptr(), o.stride(0), 64, o.stride(1),
                                  b, num_heads, tq, tk, float(sm_scale), _stream(q)), "attention")
    return o


# ------------------------------------------------------------------------------------------------
# Convolution
# ------------------------------------------------------------------------------------------------
def pack_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """(K, C, R, S) Conv2d weight -> dense KRSC (channels-last) storage; a no-op if already packed."""
    if weight.dim() != 4:
        raise ValueError("pack_conv_weight: expected a 4-D weight")
    k, c, r, s = weight.shape
    if weight.stride() == (r * s * c, 1, s * c, c):
        return weight
    return weight.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


_PADDED: dict = {}  # id(tensor) -> (weakref to the tensor, {tag: (version, padded copy)})


def _padded(t: torch.Tensor, tag: str, make):
    """Zero-padded copy of a (tiny) weight / bias, built once per tensor object + version and reused -- the
    first call happens during warm-up, so CUDA-graph capture only ever sees the cached tensor.  Keyed by the
    identity of the live tensor object (a data_ptr key could alias a freed tensor of another model).
    Returns (copy, fresh): fresh = the copy was produced by THIS call (on the current stream)."""
    key = id(t)
    entry = _PADDED.get(key)
    if entry is None or entry[0]() is not t:
        entry = (weakref.ref(t, lambda _r, k=key: _PADDED.pop(k, None)), {})
        _PADDED[key] = entry
    hit = entry[1].get(tag)
    fresh = hit is None or hit[0] != t._version
    if fresh:
        with torch.no_grad():
            new = make()
            if hit is not None and hit[1].shape == new.shape:
                hit[1].copy_(new)  # in place: a captured graph keeps reading this address
                new = hit[1]
        hit = (t._version, new, make)
        entry[1][tag] = hit
    return hit[1], fresh


def refresh_padded_() -> int:
    """Rebuild, in place, every padded operand whose source tensor has been modified since it was made (the captured
    graphs of conv_in / conv_out read the padded copies, not the parameters).  Returns how many were rewritten."""
    n = 0
    for ref, tags in list(_PADDED.values()):
        t = ref()
        if t is None:
            continue
        for tag, (version, padded, make) in list(tags.items()):
            if version != t._version:
                with torch.no_grad():
                    padded.copy_(make())
                tags[tag] = (t._version, padded, make)
                n += 1
    return n


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    _require_bf16_cuda("upsample_nearest2x", x)
    x = _nhwc(x)
    n, c, h, w = x.shape
    out = _empty_nhwc(n, c, 2 * h, 2 * w, x)
    check(lib().st_upsample_nearest2x_nhwc_bf16(x.data_ptr(), out.data_ptr(), n, h, w, c, _stream(x)), "upsample")
    return out


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torch.cat([a, b], dim=1) for channels-last 4-D tensors."""
    _require_bf16_cuda("concat_channels", a, b)
    a, b = _nhwc(a), _nhwc(b)
    n, ca, h, w = a.shape
    cb = b.shape[1]
    if b.shape != (n, cb, h, w):
        raise ValueError("concat_channels: spatial/batch mismatch")
    out = _empty_nhwc(n, ca + cb, h, w, a)
    check(lib().st_concat_channels_bf16(a.data_ptr(), ca, b.data_ptr(), cb, out.data_ptr(), n * h * w, _stream(a)),
          "concat")
    return out


def conv2d(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1, padding: int = 1,
           temb: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
           nchw_output: bool = False, block_n: int = 0, w_static: bool = False, gn_stats: bool = False):
    out, part = _conv2d_impl(x, weight, bias, stride, padding, temb, residual, nchw_output, block_n, w_static, gn_stats)
    return (out, part) if gn_stats else out


def _conv2d_impl(x, weight, bias, stride, padding, temb, residual, nchw_output, block_n, w_static, gn_stats):
    """Conv2d for the SDXL UNet sites: 3x3/pad 1 (stride 1 or 2) and 1x1/pad 0, channels-last.

    temb: (N, K) added per (image, channel) after the bias (unet_pt.py:82-83).
    residual: (N, K, H, W) added after the bias (unet_pt.py:93).  Both are fused into the GEMM epilogue.
    nchw_output: write a dense NCHW result (conv_out -> the latent the scheduler consumes).
    w_static: as in `linear` (dropped if the weight has to be re-packed / padded by this call).
    gn_stats: the result feeds a GroupNorm -- return `(y, partials)` with the per-tile column statistics emitted by the
    GEMM epilogue (None where a path / shape cannot produce them; the GroupNorm then runs its own statistics pass).
    """
    _require_bf16_cuda("conv2d", x, weight, bias, temb, residual)
    _ensure_workspace(x.device)
    if x.dim() != 4:
        raise ValueError("conv2d: expected a 4-D input")
    k, c, r, s = weight.shape
    n, cx, h, w = x.shape
    if cx != c:
        raise ValueError(f"conv2d: input has {cx} channels, weight expects {c}")
    L = lib()
    stream = _stream(x)

    def mkpart(rows: int, rows_per_image: int):
        if gn_stats and gn_partial_rows(rows, rows_per_image):
            return torch.empty((rows // 128, k, 2), dtype=torch.float32, device=x.device)
        return None

    wp = pack_conv_weight(weight)
    if wp is not weight:
        w_static = False  # packed on this stream just now (compile() pre-packs, so the hot path never gets here)
    wflag = ST_W_STATIC if w_static else 0
    if temb is not None and temb.shape != (n, k):
        raise ValueError(f"conv2d: temb must be (N, K) = ({n}, {k}), got {tuple(temb.shape)}")

    if (r, s) == (1, 1):
        if padding != 0 or stride != 1:
            raise ValueError("conv2d: 1x1 convolution supports stride 1 / padding 0 only")
        xn = _nhwc(x)
        out = _empty_nhwc(n, k, h, w, x)
        res_ptr = 0
        if residual is not None:
            residual = _nhwc(residual)
            res_ptr = residual.data_ptr()
        if temb is not None:
            raise ValueError("conv2d: temb epilogue is only wired for 3x3 convolutions")
        part = mkpart(n * h * w, h * w)
        check(L.st_gemm_bf16(xn.data_ptr(), c, wp.data_ptr(), c, out.data_ptr(), k, n * h * w, k, c, _ptr(bias),
                             res_ptr, k, wflag, block_n, _ptr(part), stream), "conv1x1")
        return out, part

    if (r, s) != (3, 3) or padding != 1:
        raise ValueError(f"conv2d: unsupported kernel {r}x{s} / padding {padding}")

    if c <= 7 or k <= 8:  # conv_in / conv_out: reshaped so that they run on the tensor-core GEMM too
        if stride != 1 or temb is not None or residual is not None:
            raise ValueError("conv2d: small-channel path supports plain stride-1 convolution only")
        if c <= 7:
            # conv_in: im2col to [M, 64] (9*C real columns) x weight padded to (K, 64); NCHW input consumed in place
            wref = weakref.ref(weight)  # the maker re-reads the LIVE weight (refresh_padded_) without keeping it alive
            wpad, fresh = _padded(weight, "smallc", lambda: torch.nn.functional.pad(
                pack_conv_weight(wref()).permute(0, 2, 3, 1).reshape(k, 9 * c), (0, 64 - 9 * c)).contiguous())
            col = torch.empty((n * h * w, 64), dtype=BF16, device=x.device)
            xs = x.stride()
            check(L.st_im2col3x3_smallc_bf16(x.data_ptr(), xs[0], xs[2], xs[3], xs[1], col.data_ptr(), n, h, w, c,
                                             stream), "im2col_smallc")
            out = _empty_nhwc(n, k, h, w, x)
            part = mkpart(n * h * w, h * w)
            check(L.st_gemm_bf16(col.data_ptr(), 64, wpad.data_ptr(), 64, out.data_ptr(), k, n * h * w, k, 64,
                                 _ptr(bias), 0, 0, 0 if fresh else wflag, block_n, _ptr(part), stream), "conv_in")
            return out, part
        # conv_out: pad the K <= 8 output channels to 8, implicit GEMM, then gather the real channels
        if c % 64 != 0 or (h * w) % 128 != 0 and 128 % (h * w) != 0:
            xn = _nhwc(x)  # shapes the tensor-core conv cannot tile: CUDA-core direct kernel
            out = torch.empty((n, k, h, w), dtype=BF16, device=x.device) if nchw_output else _empty_nhwc(n, k, h, w, x)
            os_ = out.stride()
            check(L.st_conv3x3_direct_bf16(xn.data_ptr(), h * w * c, w * c, c, 1, wp.data_ptr(), _ptr(bias),
                                           out.data_ptr(), os_[0], os_[2], os_[3], os_[1], n, h, w, c, k, stream),
                  "conv_out")
            return out, None
        xn = _nhwc(x)
        wref = weakref.ref(weight)
        w8, fresh = _padded(weight, "k8", lambda: torch.nn.functional.pad(
            pack_conv_weight(wref()).permute(0, 2, 3, 1).reshape(k, 9 * c), (0, 0, 0, 8 - k)).contiguous())
        b8 = None
        if bias is not None:
            bref = weakref.ref(bias)
            b8 = _padded(bias, "k8", lambda: torch.nn.functional.pad(bref(), (0, 8 - k)).contiguous())[0]
        y8 = torch.empty((n * h * w, 8), dtype=BF16, device=x.device)
        check(L.st_conv3x3_nhwc_bf16(xn.data_ptr(), w8.data_ptr(), _ptr(b8), y8.data_ptr(), n, h, w, c, 8, 0, 0, 0,
                                     0 if fresh else wflag, 64, 0, stream), "conv_out")
        if not nchw_output:
            return y8.view(n, h, w, 8)[..., :k].permute(0, 3, 1, 2), None
        out = torch.empty((n, k, h, w), dtype=BF16, device=x.device)
        check(L.st_nhwc_to_nchw_bf16(y8.data_ptr(), 8, out.data_ptr(), n, h * w, k, stream), "nhwc_to_nchw")
        return out, None

    xn = _nhwc(x)
    if stride == 2:
        if temb is not None or residual is not None:
            raise ValueError("conv2d: stride-2 path has no fused temb/residual epilogue")
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        col = torch.empty((n * ho * wo, 9 * c), dtype=BF16, device=x.device)
        check(L.st_im2col3x3_nhwc_bf16(xn.data_ptr(), col.data_ptr(), n, h, w, c, 2, stream), "im2col")
        out = _empty_nhwc(n, k, ho, wo, x)
        part = mkpart(n * ho * wo, ho * wo)
        check(L.st_gemm_bf16(col.data_ptr(), 9 * c, wp.data_ptr(), 9 * c, out.data_ptr(), k, n * ho * wo, k, 9 * c,
                             _ptr(bias), 0, 0, wflag, block_n, _ptr(part), stream), "conv_s2")
        return out, part
    if stride != 1:
        raise ValueError("conv2d: stride must be 1 or 2")

    out = _empty_nhwc(n, k, h, w, x)
    res_ptr = 0
    if residual is not None:
        if residual.shape != out.shape:
            raise ValueError("conv2d: residual shape mismatch")
        residual = _nhwc(residual)
        res_ptr = residual.data_ptr()
    part = mkpart(n * h * w, h * w)
    check(L.st_conv3x3_nhwc_bf16(xn.data_ptr(), wp.data_ptr(), _ptr(bias), out.data_ptr(), n, h, w, c, k, _ptr(temb),
                                 temb.stride(0) if temb is not None else 0, res_ptr, wflag, block_n, _ptr(part), stream),
          "conv3x3")
    return out, part


def implicit_gemm_fprop(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Reference-compatible name (conv_implicit_gemm.py:143-182): a NHWC (N,H,W,C), b KRSC (K,3,3,C) ->
    NPQK (N,H,W,K); 3x3, pad 1, stride 1, no bias."""
    _require_bf16_cuda("implicit_gemm_fprop", a, b)
    y = conv2d(a.permute(0, 3, 1, 2), b.permute(0, 3, 1, 2), None)
    return y.permute(0, 2, 3, 1)


# ------------------------------------------------------------------------------------------------
# Embedding / scheduler glue
# ------------------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, num_channels: int) -> torch.Tensor:
    """cat([cos, sin]) sinusoidal embedding (unet_pt.py:22-36) of a 1-D fp32 tensor; bf16 (B, num_channels)."""
    if t.device.type != "cuda":
        raise ValueError("timestep_embedding: tensor must live on a CUDA device")
    t = t.reshape(-1).to(torch.float32).contiguous()
    half = num_channels // 2
    out = torch.empty((t.numel(), 2 * half), dtype=BF16, device=t.device)
    check(lib().st_timestep_embedding_bf16(t.data_ptr(), out.data_ptr(), 2 * half, t.numel(), half, _stream(t)),
          "timestep_embedding")
    return out
