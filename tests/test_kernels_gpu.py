"""Per-op parity on the GPU: every C-ABI kernel (called through stabletriton_b200.kernels) against the
per-op oracles of oracle/unet_oracle.py -- i.e. the *pattern* side of the reference's fx rewrites -- on
seeded bf16 inputs at SDXL shapes, plus the edge cases (ragged rows/columns, Tk = 77, small feature maps).
The oracle runs in fp32 on the same bf16-rounded inputs; tolerance = bf16 output rounding (max|d|/max|ref|
<= 1e-2, cosine >= 0.9999) unless stated."""
import importlib.util
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle():
    spec = importlib.util.spec_from_file_location("unet_oracle", os.path.join(ROOT, "oracle", "unet_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


O = _oracle()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(hash((shape, seed)) % (2 ** 31))
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


def check(got, ref, rel_tol=1e-2, cos_tol=0.9999):
    rel, cos = parity(got.float(), ref.float())
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert rel <= rel_tol and cos >= cos_tol, (rel, cos)
    return rel, cos


@pytest.fixture(scope="module")
def K(built_lib):
    import stabletriton_b200.kernels as kernels
    return kernels


@pytest.mark.parametrize("n,c,hw,groups,silu,eps", [
    (2, 320, 128, 32, True, 1e-5), (2, 640, 64, 32, True, 1e-5), (2, 1280, 32, 32, False, 1e-6),
    (2, 960, 64, 32, True, 1e-5), (2, 2560, 32, 32, True, 1e-5), (1, 1920, 32, 32, True, 1e-5),
    (3, 64, 10, 8, False, 1e-5),  # odd spatial size, non-SDXL group width
    (2, 640, 128, 32, True, 1e-5), (1, 960, 128, 32, True, 1e-5),  # group slab split over 4- / 8-CTA clusters
    (2, 1280, 64, 32, False, 1e-6), (2, 320, 64, 32, True, 1e-5),  # 2-CTA cluster; 4-byte vectors (10 channels / group)
    (2, 56, 12, 8, True, 1e-5),   # 7 channels per group: not slab-eligible (odd width) -> stats + apply kernels
])
def test_groupnorm(K, n, c, hw, groups, silu, eps):
    x = rnd(n, c, hw, hw, seed=1) + 0.5  # non-zero mean
    w, b = rnd(c, seed=2) * 0.1 + 1.0, rnd(c, seed=3) * 0.1
    ref = O.group_norm(x.float(), groups, w.float(), b.float(), eps, silu)
    xg = x.cuda().contiguous(memory_format=torch.channels_last)
    got = K.groupnorm_wrapper(xg, groups, w.cuda(), b.cuda(), eps, silu)
    assert got.is_contiguous(memory_format=torch.channels_last)
    check(got.cpu(), ref, rel_tol=1.5e-2)
    # NCHW-contiguous input is accepted too (converted once)
    got2 = K.groupnorm_wrapper(x.cuda(), groups, w.cuda(), b.cuda(), eps, silu)
    assert torch.equal(got2, got)


def test_groupnorm_launches_on_two_streams_may_overlap(K):
    """Re-entrancy (no global mutable state shared between calls): every GroupNorm launch has its own
    arrival tickets, so launches on different streams -- or in parallel branches of one CUDA graph -- may overlap."""
    xs = [rnd(2, 320, 64, 64, seed=70 + i).cuda().contiguous(memory_format=torch.channels_last) for i in range(2)]
    w, b = (rnd(320, seed=72) * 0.1 + 1.0).cuda(), (rnd(320, seed=73) * 0.1).cuda()
    ref = [K.groupnorm_wrapper(x, 32, w, b, 1e-5, True) for x in xs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [[], []]
    for _ in range(100):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                outs[i].append(K.groupnorm_wrapper(xs[i], 32, w, b, 1e-5, True))
    torch.cuda.synchronize()
    for i in range(2):
        assert all(torch.equal(o, ref[i]) for o in outs[i])


def test_groupnorm_large_mean_is_stable(K):
    x = (rnd(2, 320, 32, 32, seed=4) * 0.5 + 40.0)
    w, b = torch.ones(320, dtype=torch.bfloat16), torch.zeros(320, dtype=torch.bfloat16)
    ref = O.group_norm(x.float(), 32, w.float(), b.float(), 1e-5)
    got = K.groupnorm_wrapper(x.cuda(), 32, w.cuda(), b.cuda(), 1e-5, False)
    check(got.cpu(), ref, rel_tol=2e-2)


@pytest.mark.parametrize("m,n", [(8192, 640), (2048, 1280), (77, 2048), (5, 64)])
def test_layernorm(K, m, n):
    x = rnd(2, m // 2 if m % 2 == 0 else m, n, seed=5) * 2 + 0.3
    w, b = rnd(n, seed=6) * 0.1 + 1.0, rnd(n, seed=7) * 0.1
    ref = O.layer_norm(x.float(), w.float(), b.float(), 1e-5)
    got = K.layer_norm(x.cuda(), w.cuda(), b.cuda(), 1e-5)
    check(got.cpu(), ref)


@pytest.mark.parametrize("m,k,n,act,res", [
    (2048, 1280, 1280, False, True), (8192, 640, 640, False, False), (154, 2048, 1280, False, False),
    (2048, 5120, 1280, False, True), (300, 192, 200, True, True), (2, 1280, 320, True, False),
    (16, 2816, 1280, False, False),
])
def test_linear(K, m, k, n, act, res):
    x, w, b = rnd(m, k, seed=8), rnd(n, k, scale=k ** -0.5, seed=9), rnd(n, seed=10) * 0.1
    r = rnd(m, n, seed=11) if res else None
    ref = F.linear(x.float(), w.float(), b.float())
    ref = F.silu(ref) if act else ref
    ref = ref + r.float() if res else ref
    got = K.linear(x.cuda(), w.cuda(), b.cuda(), activation=act, residual=None if r is None else r.cuda())
    check(got.cpu(), ref)
    if m > 32 and not res:
        got2 = K.sdxl_forward(x.cuda(), w.cuda(), b.cuda(), act)  # reference-named entry point
        assert torch.equal(got2, got)


def test_linear_3d_strided_input_and_qkv_slices(K):
    """(B, T, 3C) fused projection output consumed through column slices, as the fx passes arrange it."""
    x = rnd(2, 256, 640, seed=12)
    wqkv = rnd(1920, 640, scale=640 ** -0.5, seed=13)
    fused = K.linear(x.cuda(), wqkv.cuda())
    q, k, v = fused[..., :640], fused[..., 640:1280], fused[..., 1280:]
    ref_q, ref_k, ref_v = F.linear(x.float(), wqkv.float()).chunk(3, dim=-1)
    check(q.cpu(), ref_q)
    got = K.attention_btc(q, k, v, 10, 0.125)
    ref = O.attention_core(ref_q.to(torch.bfloat16).float(), ref_k.to(torch.bfloat16).float(),
                           ref_v.to(torch.bfloat16).float(), 10, 64)
    check(got.cpu(), ref, rel_tol=2e-2, cos_tol=0.9995)
    # a slice is also a legal GEMM input (row pitch 1920)
    wo = rnd(640, 640, scale=640 ** -0.5, seed=14)
    check(K.linear(q, wo.cuda()).cpu(), F.linear(q.float().cpu(), wo.float()))


@pytest.mark.parametrize("m,k,n", [(2048, 1280, 10240), (8192, 640, 5120), (130, 128, 512)])
def test_linear_geglu_fused_epilogue(K, m, k, n):
    x, w, b = rnd(m, k, seed=15), rnd(n, k, scale=k ** -0.5, seed=16), rnd(n, seed=17) * 0.1
    s, g = F.linear(x.float(), w.float(), b.float()).chunk(2, dim=-1)
    ref = O.geglu(s, g)
    got = K.linear(x.cuda(), w.cuda(), b.cuda(), geglu=True)
    check(got.cpu(), ref)
    # standalone elementwise GEGLU on the strided halves of an un-fused projection (reference's seam)
    proj = K.linear(x.cuda(), w.cuda(), b.cuda())
    st, gt = proj.chunk(2, dim=-1)
    got2 = K.geglu_wrapper(st, gt)
    check(got2.cpu(), ref, rel_tol=2e-2)


@pytest.mark.parametrize("b,h,tq,tk", [(2, 10, 4096, 4096), (2, 20, 1024, 1024), (2, 10, 4096, 77), (2, 20, 1024, 77),
                                        (1, 3, 200, 333), (1, 1, 1, 77),
                                        # short-context kernel (Tk <= 80): full box, exactly 64, ragged, a single key
                                        (1, 2, 300, 80), (1, 2, 130, 64), (2, 3, 128, 17), (1, 1, 77, 1),
                                        # one-block two-CTA kernel (80 < Tk <= 128)
                                        (1, 2, 200, 100), (1, 2, 256, 128), (1, 2, 100, 81)])
def test_attention(K, b, h, tq, tk):
    q, k, v = rnd(b, tq, h * 64, seed=18), rnd(b, tk, h * 64, seed=19), rnd(b, tk, h * 64, seed=20)
    ref = O.attention_core(q.float(), k.float(), v.float(), h, 64)
    got = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
    check(got.cpu(), ref, rel_tol=2e-2, cos_tol=0.9995)
    if tq <= 1024:  # the reference kernel's own (B, H, T, D) calling convention (attention_fa2.py:113)
        qh, kh, vh = (t.cuda().view(b, -1, h, 64).transpose(1, 2).contiguous() for t in (q, k, v))
        got2 = K.attention(qh, kh, vh, 0.125).transpose(1, 2).reshape(b, tq, h * 64)
        assert torch.equal(got2, got)


@pytest.mark.parametrize("tk,growth", [(1024, 24.0), (1000, -0.9), (300, 6.0), (129, 6.0)])
def test_attention_reference_moves(K, tk, growth):
    """Self-attention sweep kernel (Tk > 128): keys scaled by 1 + growth * t / Tk make later K/V blocks carry much larger
    (or much smaller) scores, which exercises the lazy softmax-reference update -- O rescaled in tensor memory when the
    running row maximum grows by more than 2^8 -- the ragged last block and the single-key last block."""
    b, h, tq = 1, 2, 256
    q, k, v = rnd(b, tq, h * 64, seed=41) * 1.5, rnd(b, tk, h * 64, seed=42) * 1.5, rnd(b, tk, h * 64, seed=43)
    ramp = (1.0 + growth * torch.arange(tk, dtype=torch.float32) / tk).view(1, tk, 1)
    k = (k.float() * ramp).to(torch.bfloat16)
    ref = O.attention_core(q.float(), k.float(), v.float(), h, 64)
    got = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
    check(got.cpu(), ref, rel_tol=2e-2, cos_tol=0.9995)


@pytest.mark.parametrize("impl", [2, 4], ids=["pipelined", "resident"])
@pytest.mark.parametrize("b,h,tq,tk,growth", [(2, 20, 1024, 1024, 0.0), (1, 3, 200, 333, 0.0), (1, 2, 256, 1024, 24.0),
                                               (1, 2, 256, 1000, -0.9), (2, 3, 130, 129, 6.0), (1, 2, 100, 65, 0.0),
                                               (1, 38, 1024, 512, 0.0), (2, 2, 4096, 2048, 3.0)])
def test_attention_sweep_kernels_forced(K, impl, b, h, tq, tk, growth):
    """st_attention_bf16 sends a K/V sweep (Tk > 128) to the resident kernel while all query tiles fit on the machine at
    once (3 CTAs per SM) and to the pipelined kernel beyond that; here both are forced onto the same inputs, small and
    large, with moving / never-moving softmax references and ragged tails.  Same arithmetic per element but different
    block sizes (64 / 128 keys), so they agree within rounding, not bit for bit."""
    from stabletriton_b200 import _cabi
    L = _cabi._load()
    q, k, v = rnd(b, tq, h * 64, seed=61) * 1.5, rnd(b, tk, h * 64, seed=62) * 1.5, rnd(b, tk, h * 64, seed=63)
    if growth:
        ramp = (1.0 + growth * torch.arange(tk, dtype=torch.float32) / tk).view(1, tk, 1)
        k = (k.float() * ramp).to(torch.bfloat16)
    ref = O.attention_core(q.float(), k.float(), v.float(), h, 64)
    try:
        L.st_debug_set_attention_impl(impl)
        got = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
        again = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
        torch.cuda.synchronize()
    finally:
        L.st_debug_set_attention_impl(0)
    check(got.cpu(), ref, rel_tol=2e-2, cos_tol=0.9995)
    assert torch.equal(got, again)


def test_attention_resident_four_cta_form(K):
    """Saturated sweeps (>= 8 query tiles per SM) take the four-CTA form of the resident kernel (upper half of the score row
    re-read from tensor memory): same per-element arithmetic as the three-CTA form, so the outputs agree bit for bit -- checked
    on a launch big enough to be sent there by shape (B = 8, 20 heads, T = 1024: 1280 tiles) against the forced pipelined kernel
    and the oracle, ragged tail included."""
    from stabletriton_b200 import _cabi
    L = _cabi._load()
    b, h, tq, tk = 8, 20, 1024, 1000
    q, k, v = rnd(b, tq, h * 64, seed=81) * 1.5, rnd(b, tk, h * 64, seed=82) * 1.5, rnd(b, tk, h * 64, seed=83)
    ramp = (1.0 + 5.0 * torch.arange(tk, dtype=torch.float32) / tk).view(1, tk, 1)
    k = (k.float() * ramp).to(torch.bfloat16)
    got = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)     # by shape: resident, four CTAs per SM
    try:
        L.st_debug_set_attention_impl(2)
        pip = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
        torch.cuda.synchronize()
    finally:
        L.st_debug_set_attention_impl(0)
    ref = O.attention_core(q[:2].float(), k[:2].float(), v[:2].float(), h, 64)
    check(got[:2].cpu(), ref, rel_tol=2e-2, cos_tol=0.9995)
    rel, cos = parity(got.float(), pip.float())
    assert rel <= 1e-2 and cos >= 0.99999, (rel, cos)


def test_attention_long_sweep(K):
    """T = 16384 (the 2048^2 self-attention of SURVEY 8d config 5) against fp32 softmax(QK^T)V on the GPU, one head."""
    b, h, t = 1, 1, 16384
    q, k, v = (rnd(b, t, h * 64, seed=s).cuda() for s in (44, 45, 46))
    got = K.attention_btc(q, k, v, h, 0.125)
    qf, kf, vf = (x.float().view(b, t, h, 64).transpose(1, 2) for x in (q, k, v))
    ref = torch.softmax(qf[:, :, :2048] @ kf.transpose(-2, -1) * 0.125, dim=-1) @ vf  # first 2048 query rows, fp32
    check(got.view(b, t, h, 64).transpose(1, 2)[:, :, :2048].float().cpu(), ref.cpu(), rel_tol=2e-2, cos_tol=0.9995)


@pytest.mark.parametrize("n,c,k,hw,mode", [
    (2, 320, 320, 128, "temb"), (2, 640, 640, 64, "res"), (2, 1280, 1280, 32, "temb"), (2, 960, 640, 64, "plain"),
    (2, 2560, 1280, 32, "plain"), (1, 64, 128, 16, "res"), (3, 64, 64, 8, "temb"),  # small maps: multi-image tiles
])
def test_conv3x3_implicit_gemm(K, n, c, k, hw, mode):
    x, w, b = rnd(n, c, hw, hw, seed=21), rnd(k, c, 3, 3, scale=(9 * c) ** -0.5, seed=22), rnd(k, seed=23) * 0.1
    temb = rnd(n, k, seed=24) if mode == "temb" else None
    res = rnd(n, k, hw, hw, seed=25) if mode == "res" else None
    ref = F.conv2d(x.float(), w.float(), b.float(), padding=1)
    if temb is not None:
        ref = ref + temb.float()[:, :, None, None]
    if res is not None:
        ref = ref + res.float()
    got = K.conv2d(x.cuda().contiguous(memory_format=torch.channels_last), w.cuda(), b.cuda(),
                   temb=None if temb is None else temb.cuda(),
                   residual=None if res is None else res.cuda().contiguous(memory_format=torch.channels_last))
    assert got.is_contiguous(memory_format=torch.channels_last) or hw == 1
    check(got.cpu(), ref)


def test_conv_other_sites(K):
    """conv_in (C=4, NCHW input), conv_out (K=4, NCHW output), stride-2 downsampler, 1x1 shortcut, upsample."""
    x = rnd(2, 4, 32, 32, seed=26)
    w, b = rnd(320, 4, 3, 3, scale=1 / 6, seed=27), rnd(320, seed=28) * 0.1
    check(K.conv2d(x.cuda(), w.cuda(), b.cuda()).cpu(), F.conv2d(x.float(), w.float(), b.float(), padding=1))
    x = rnd(2, 320, 32, 32, seed=29)
    w, b = rnd(4, 320, 3, 3, scale=2880 ** -0.5, seed=30), rnd(4, seed=31) * 0.1
    got = K.conv2d(x.cuda(), w.cuda(), b.cuda(), nchw_output=True)
    assert got.is_contiguous()
    check(got.cpu(), F.conv2d(x.float(), w.float(), b.float(), padding=1))
    w, b = rnd(320, 320, 3, 3, scale=2880 ** -0.5, seed=32), rnd(320, seed=33) * 0.1
    check(K.conv2d(x.cuda(), w.cuda(), b.cuda(), stride=2).cpu(),
          F.conv2d(x.float(), w.float(), b.float(), stride=2, padding=1))
    w1, b1 = rnd(640, 320, 1, 1, scale=320 ** -0.5, seed=34), rnd(640, seed=35) * 0.1
    check(K.conv2d(x.cuda(), w1.cuda(), b1.cuda(), padding=0).cpu(), F.conv2d(x.float(), w1.float(), b1.float()))
    up = K.upsample_nearest2x(x.cuda())
    assert torch.equal(up.cpu(), F.interpolate(x.float(), scale_factor=2.0, mode="nearest").to(torch.bfloat16))
    a, bb = rnd(2, 64, 8, 8, seed=36), rnd(2, 128, 8, 8, seed=37)
    assert torch.equal(K.concat_channels(a.cuda(), bb.cuda()).cpu(), torch.cat([a, bb], dim=1))
    # reference-named NHWC x KRSC -> NPQK entry point
    xa, wk = rnd(1, 16, 16, 64, seed=38), rnd(64, 3, 3, 64, scale=1 / 24, seed=39)
    ref = F.conv2d(xa.float().permute(0, 3, 1, 2), wk.float().permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    check(K.implicit_gemm_fprop(xa.cuda(), wk.cuda()).cpu(), ref)


def test_timestep_embedding(K):
    t = torch.tensor([999.0, 1.0, 500.0, 958.0])
    for ch in (320, 256):
        got = K.timestep_embedding(t.cuda(), ch)
        check(got.cpu(), O.timesteps_embedding(t, ch), rel_tol=5e-3)


def test_shape_and_dtype_errors_are_raised_before_launch(K):
    with pytest.raises(TypeError):
        K.linear(torch.zeros(4, 64, device="cuda"), torch.zeros(64, 64, device="cuda"))  # fp32
    with pytest.raises(ValueError):
        K.linear(rnd(64, 100).cuda(), rnd(64, 100).cuda())  # K not a multiple of 64 on the tensor-core path
    with pytest.raises(ValueError):
        K.groupnorm_wrapper(rnd(1, 64, 8, 8).cuda(), 32, None, None, 1e-5)  # 2 channels per group
    with pytest.raises(ValueError):
        K.attention_btc(rnd(1, 8, 100).cuda(), rnd(1, 8, 100).cuda(), rnd(1, 8, 100).cuda(), 2, 0.1)
    with pytest.raises(ValueError):
        K.conv2d(rnd(1, 64, 12, 12).cuda(), rnd(64, 64, 3, 3).cuda(), None)  # 144 pixels: not tileable


def test_2048px_shapes(K):
    """SURVEY 8d config 5 (256 x 256 latents): the largest activations of the path, against fp32 torch on the GPU."""
    x = rnd(2, 320, 256, 256, seed=51).cuda().contiguous(memory_format=torch.channels_last)
    w, b = rnd(320, 320, 3, 3, scale=2880 ** -0.5, seed=52).cuda(), (rnd(320, seed=53) * 0.1).cuda()
    ref = F.conv2d(x.float(), w.float(), b.float(), padding=1)
    check(K.conv2d(x, w, b).float().cpu(), ref.cpu())
    gw, gb = (1 + 0.1 * rnd(320, seed=54)).cuda(), (0.1 * rnd(320, seed=55)).cuda()
    ref = F.silu(F.group_norm(x.float(), 32, gw.float(), gb.float(), 1e-5))
    check(K.groupnorm_wrapper(x, 32, gw, gb, 1e-5, activation=True).float().cpu(), ref.cpu())
    t = rnd(32768, 640, seed=56).cuda()
    lw, lb = (1 + 0.1 * rnd(640, seed=57)).cuda(), (0.1 * rnd(640, seed=58)).cuda()
    check(K.layer_norm(t, lw, lb, 1e-5).float().cpu(), F.layer_norm(t.float(), (640,), lw.float(), lb.float(), 1e-5).cpu())
    wl, bl = rnd(1920, 640, scale=640 ** -0.5, seed=59).cuda(), (rnd(1920, seed=60) * 0.1).cuda()
    check(K.linear(t, wl, bl).float().cpu(), F.linear(t.float(), wl.float(), bl.float()).cpu())


# ---- GroupNorm statistics emitted by the producer's epilogue (round 2) -----------------------------------------------
def _tile_stats(y_rows: torch.Tensor):
    """(mean, M2) per 128-row tile and column of a [M, C] tensor, in float64 -- what gn_partial must hold."""
    m, c = y_rows.shape
    t = y_rows.double().reshape(m // 128, 128, c)
    mean = t.mean(dim=1)
    return mean, ((t - mean[:, None, :]) ** 2).sum(dim=1)


def _check_partials(part, y_rows):
    mean, m2 = _tile_stats(y_rows.cpu())
    got = part.double().cpu()
    assert got.shape == (mean.shape[0], mean.shape[1], 2)
    assert torch.allclose(got[..., 0], mean, rtol=1e-5, atol=1e-5), (got[..., 0] - mean).abs().max()
    assert torch.allclose(got[..., 1], m2, rtol=2e-3, atol=1e-3), ((got[..., 1] - m2).abs() / (m2.abs() + 1e-3)).max()


@pytest.mark.parametrize("m,k,n,res,hw", [(2048, 1280, 1280, True, 1024), (8192, 640, 640, True, 4096),
                                          (2048, 320, 200, False, 1024), (256, 64, 72, False, 128)])
def test_gemm_epilogue_emits_groupnorm_partials(K, m, k, n, res, hw):
    """st_gemm_bf16(gn_partial=...): per-tile column statistics of exactly the bf16 values it stores; the stored
    output is bit-identical to the launch without them (the proj_out + residual -> norm1 site)."""
    x, w, b = rnd(m, k, seed=50), rnd(n, k, scale=k ** -0.5, seed=51), rnd(n, seed=52) * 0.1
    r = (rnd(m, n, seed=53) + 2.0) if res else None  # a residual with a non-zero mean: |mean| > sigma per column
    args = (x.cuda(), w.cuda(), b.cuda())
    kw = dict(residual=None if r is None else r.cuda())
    plain = K.linear(*args, **kw)
    y, part = K.linear(*args, **kw, gn_stats=hw)
    assert torch.equal(y, plain)
    _check_partials(part, y.float())
    # shapes whose tiles would straddle images get no partials (the GroupNorm then makes its own pass)
    y2, none = K.linear(*args, **kw, gn_stats=96)
    assert none is None and torch.equal(y2, plain)


@pytest.mark.parametrize("n,c,k,hw,mode", [(2, 320, 320, 128, "temb"), (2, 640, 640, 64, "res"),
                                           (2, 1280, 1280, 32, "res"), (1, 64, 128, 16, "plain")])
def test_conv_epilogue_emits_groupnorm_partials(K, n, c, k, hw, mode):
    x, w, b = rnd(n, c, hw, hw, seed=21), rnd(k, c, 3, 3, scale=(9 * c) ** -0.5, seed=22), rnd(k, seed=23) * 0.1
    temb = rnd(n, k, seed=24).cuda() if mode == "temb" else None
    res = (rnd(n, k, hw, hw, seed=25) + 1.5).cuda().contiguous(memory_format=torch.channels_last) if mode == "res" else None
    xg = x.cuda().contiguous(memory_format=torch.channels_last)
    plain = K.conv2d(xg, w.cuda(), b.cuda(), temb=temb, residual=res)
    y, part = K.conv2d(xg, w.cuda(), b.cuda(), temb=temb, residual=res, gn_stats=True)
    assert torch.equal(y, plain)
    _check_partials(part, y.permute(0, 2, 3, 1).reshape(-1, k).float())


@pytest.mark.parametrize("n,c,hw,groups,silu,eps,offset", [
    (2, 320, 128, 32, True, 1e-5, 0.5), (2, 640, 64, 32, True, 1e-5, 0.5), (2, 1280, 32, 32, False, 1e-6, 0.5),
    (2, 320, 32, 32, False, 1e-5, 40.0),  # |mean| = 80 sigma: the tile-wise (mean, M2) form must not cancel
    (16, 320, 32, 32, True, 1e-5, 0.5)])
def test_groupnorm_from_producer_partials(K, n, c, hw, groups, silu, eps, offset):
    """GroupNorm fed by the statistics of its producer: a 1x1 convolution writes x (and its partials), the GroupNorm
    reads x once.  Must agree with the fp32 oracle on the SAME x and with the stand-alone two-pass kernel."""
    src = rnd(n, 64, hw, hw, seed=60).cuda().contiguous(memory_format=torch.channels_last)
    w1 = (rnd(c, 64, 1, 1, scale=0.125 if offset < 10 else 0.0625, seed=61)).cuda()
    b1 = (rnd(c, seed=62) * 0.1 + offset).cuda()
    x, part = K.conv2d(src, w1, b1, padding=0, gn_stats=True)
    assert part is not None
    gw, gb = (rnd(c, seed=2) * 0.1 + 1.0).cuda(), (rnd(c, seed=3) * 0.1).cuda()
    got = K.groupnorm_wrapper(x, groups, gw, gb, eps, silu, partials=(part,))
    alone = K.groupnorm_wrapper(x, groups, gw, gb, eps, silu)
    ref = O.group_norm(x.float().cpu(), groups, gw.float().cpu(), gb.float().cpu(), eps, silu)
    check(got.cpu(), ref, rel_tol=2e-2 if offset > 10 else 1.5e-2)
    rel, _ = parity(got.float(), alone.float())
    assert rel <= 1.6e-2  # the two kernels may differ by one bf16 ulp where the statistics differ in the last bits


def test_groupnorm_of_a_concatenation_uses_both_producers_partials(K):
    """Up-block site (unet_pt.py:356): norm1(cat([hidden, skip], 1)) -- the concatenation's statistics are its two
    producers' partials side by side, also where a group straddles the seam (1280 + 640 channels, 60 per group)."""
    n, hw = 2, 32
    src = rnd(n, 64, hw, hw, seed=60).cuda().contiguous(memory_format=torch.channels_last)
    a, pa = K.conv2d(src, rnd(1280, 64, 1, 1, scale=0.125, seed=61).cuda(), (rnd(1280, seed=62) + 0.3).cuda(), padding=0,
                     gn_stats=True)
    b, pb = K.conv2d(src, rnd(640, 64, 1, 1, scale=0.25, seed=63).cuda(), (rnd(640, seed=64) - 0.2).cuda(), padding=0,
                     gn_stats=True)
    x = K.concat_channels(a, b)
    gw, gb = (rnd(1920, seed=2) * 0.1 + 1.0).cuda(), (rnd(1920, seed=3) * 0.1).cuda()
    got = K.groupnorm_wrapper(x, 32, gw, gb, 1e-5, True, partials=(pa, pb))
    ref = O.group_norm(x.float().cpu(), 32, gw.float().cpu(), gb.float().cpu(), 1e-5, True)
    check(got.cpu(), ref, rel_tol=1.5e-2)
    with pytest.raises(ValueError):
        K.groupnorm_wrapper(x, 32, gw, gb, 1e-5, True, partials=(pa,))  # partials do not cover all channels


def test_groupnorm_large_batch(K):
    """N = 16 at (320, 128^2): 168 MB, more than the L2 holds."""
    x = (rnd(16, 320, 128, 128, seed=5) + 0.5).cuda().contiguous(memory_format=torch.channels_last)
    w, b = (rnd(320, seed=2) * 0.1 + 1.0).cuda(), (rnd(320, seed=3) * 0.1).cuda()
    got = K.groupnorm_wrapper(x, 32, w, b, 1e-5, True)
    ref = F.silu(F.group_norm(x.float(), 32, w.float(), b.float(), 1e-5))
    check(got.cpu(), ref.cpu(), rel_tol=1.5e-2)


def test_two_captured_graphs_with_groupnorm_replay_concurrently(K):
    """the last-CTA tickets of the stand-alone GroupNorm live in the call's own workspace, so two captured
    graphs (each with its own pool) replayed at the same time on two streams cannot disturb each other."""
    w, b = (rnd(320, seed=72) * 0.1 + 1.0).cuda(), (rnd(320, seed=73) * 0.1).cuda()
    xs = [rnd(2, 320, 64, 64, seed=80 + i).cuda().contiguous(memory_format=torch.channels_last) for i in range(2)]
    ref = [K.groupnorm_wrapper(x, 32, w, b, 1e-5, True) for x in xs]
    graphs, outs, streams = [], [], [torch.cuda.Stream(), torch.cuda.Stream()]
    for i in range(2):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(streams[i]):
            K.groupnorm_wrapper(xs[i], 32, w, b, 1e-5, True)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=streams[i]):
                y = xs[i]
                for _ in range(8):  # a chain of GroupNorms keeps each graph busy for a while
                    y = K.groupnorm_wrapper(xs[i], 32, w, b, 1e-5, True)
        graphs.append(g)
        outs.append(y)
    torch.cuda.synchronize()
    for _ in range(50):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                graphs[i].replay()
    torch.cuda.synchronize()
    for i in range(2):
        assert torch.equal(outs[i], ref[i])


def test_outputs_do_not_overrun_and_repeat_bit_for_bit(K):
    """compute-sanitizer is closed on the GPU pool, so the memcheck / racecheck role is played by canaries and
    repetition: every kernel family writes into a buffer embedded in sentinel-filled memory (TMA stores clipped at
    ragged edges, strided epilogues) and is run 20 times -- any out-of-bounds store trips a sentinel, any shared-memory
    race between the TMA / MMA / epilogue warps shows up as a run-to-run difference."""
    torch.manual_seed(0)

    def guarded(fn, reps=20):
        first = None
        for i in range(reps):
            torch.cuda.empty_cache()
            pad = torch.full((1 << 20,), -7.0, dtype=torch.bfloat16, device="cuda")  # neighbours in the caching allocator
            out = fn()
            pad2 = torch.full((1 << 20,), -7.0, dtype=torch.bfloat16, device="cuda")
            torch.cuda.synchronize()
            assert bool((pad == -7.0).all()) and bool((pad2 == -7.0).all()), "sentinel overwritten"
            assert torch.isfinite(out.float()).all()
            if first is None:
                first = out.clone()
            else:
                assert torch.equal(out, first), f"run {i} differs from run 0"
            del pad, pad2
        return first

    x, w, b = rnd(300, 192, seed=1).cuda(), rnd(328, 192, scale=0.07, seed=2).cuda(), rnd(328, seed=3).cuda()
    r = rnd(300, 328, seed=4).cuda()
    guarded(lambda: K.linear(x, w, b, residual=r))                       # ragged M and N, TMA-store clipping
    xg, wg, bg = rnd(384, 128, seed=5).cuda(), rnd(512, 128, scale=0.09, seed=6).cuda(), rnd(512, seed=7).cuda()
    guarded(lambda: K.linear(xg, wg, bg, geglu=True))
    xc = rnd(2, 64, 16, 16, seed=8).cuda().contiguous(memory_format=torch.channels_last)
    wc, bc = rnd(192, 64, 3, 3, scale=1 / 24, seed=9).cuda(), rnd(192, seed=10).cuda()
    guarded(lambda: K.conv2d(xc, wc, bc, temb=rnd(2, 192, seed=11).cuda()))
    guarded(lambda: K.conv2d(xc, wc, bc, gn_stats=True)[1])
    q, k, v = rnd(1, 300, 128, seed=12).cuda(), rnd(1, 300, 128, seed=13).cuda(), rnd(1, 300, 128, seed=14).cuda()
    guarded(lambda: K.attention_btc(q, k, v, 2, 0.125))                  # pipelined kernel, ragged Tq / Tk
    kc, vc = rnd(1, 77, 128, seed=15).cuda(), rnd(1, 77, 128, seed=16).cuda()
    guarded(lambda: K.attention_btc(q, kc, vc, 2, 0.125))                # one-block kernel
    xn = rnd(2, 320, 20, 20, seed=17).cuda().contiguous(memory_format=torch.channels_last)
    gw, gb = rnd(320, seed=18).cuda(), rnd(320, seed=19).cuda()
    guarded(lambda: K.groupnorm_wrapper(xn, 32, gw, gb, 1e-5, True))
    xl = rnd(77, 640, seed=20).cuda()
    guarded(lambda: K.layer_norm(xl, rnd(640, seed=21).cuda(), rnd(640, seed=22).cuda(), 1e-5))


@pytest.mark.parametrize("m,k,n,block_n", [
    (2048, 1280, 1280, 0), (2048, 320, 1280, 64), (2048, 320, 1280, 128), (2048, 320, 1280, 160), (2048, 320, 1280, 192),
    (2048, 320, 1280, 256), (2048, 320, 1280, -160), (2048, 320, 1280, -256),
    (8192, 640, 640, 0),      # two tiles per CTA: the first tile's residual by row loads, the last one's by TMA
    (40000, 64, 320, 160),    # many tiles per CTA, one k-block each: the ring is still in its first pass when the residual lands
    (300, 192, 328, 0), (130, 64, 72, 0), (257, 128, 200, 160),  # ragged rows / columns: boxes clipped on both edges
])
def test_linear_residual_tile_by_tma(K, m, k, n, block_n):
    """The residual of a CTA's last output tile is fetched by TMA into the idle operand ring (same boxes / swizzle as the
    store map) and added from there; earlier tiles of a multi-tile CTA use per-thread row loads.  Both against the oracle,
    with the residual given as a column slice of a wider tensor (row pitch != N), for every tile width and for CTA pairs."""
    x, w, b = rnd(m, k, seed=71), rnd(n, k, scale=k ** -0.5, seed=72), rnd(n, seed=73) * 0.1
    wide = rnd(m, n + 24, seed=74)
    r = wide[:, 16:16 + n]                      # pitch n + 24, 32-byte aligned start
    ref = F.linear(x.float(), w.float(), b.float()) + r.float()
    rc = wide.cuda()[:, 16:16 + n]
    got = K.linear(x.cuda(), w.cuda(), b.cuda(), residual=rc, block_n=block_n)
    check(got.cpu(), ref)
    again = K.linear(x.cuda(), w.cuda(), b.cuda(), residual=rc, block_n=block_n)
    assert torch.equal(got, again)
    inplace = rc.clone()                        # residual aliasing the output: a tile is read and written by one CTA only
    K.linear(x.cuda(), w.cuda(), b.cuda(), residual=inplace, block_n=block_n, out=inplace)
    assert torch.equal(inplace, got)


# ---- tile shapes of round 2: 160-wide tiles (32-column tail group) and CTA pairs (cta_group::2) -----------------------
@pytest.mark.parametrize("m,k,n,block_n", [
    (2048, 1280, 1280, 160), (2048, 1280, 1280, -160), (2048, 5120, 1280, -160), (2048, 1280, 1280, -192),
    (2048, 1280, 3840, -256), (400, 192, 328, 160), (400, 192, 328, -160), (512, 64, 200, -160), (8192, 640, 640, -160)])
def test_linear_tile_shapes_and_cta_pairs(K, m, k, n, block_n):
    """block_n > 0 forces the tile width, block_n < 0 a CTA-pair launch (two CTAs share one 256 x |block_n| tcgen05
    tile, each staging half of the weight tile).  Same result, bit for bit, as the launcher's own choice -- the K loop
    order per output element does not depend on the tile shape -- and the GroupNorm partials ride along."""
    x, w, b = rnd(m, k, seed=90), rnd(n, k, scale=k ** -0.5, seed=91), rnd(n, seed=92) * 0.1
    r = rnd(m, n, seed=93)
    ref = F.linear(x.float(), w.float(), b.float()) + r.float()
    auto = K.linear(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda(), block_n=128)
    got = K.linear(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda(), block_n=block_n)
    check(got.cpu(), ref)
    assert torch.equal(got, auto)
    if m % 128 == 0:
        y, part = K.linear(x.cuda(), w.cuda(), b.cuda(), residual=r.cuda(), block_n=block_n, gn_stats=128)
        assert torch.equal(y, got)
        _check_partials(part, y.float())


@pytest.mark.parametrize("m,k,n,block_n", [(2048, 1280, 10240, -256), (8192, 640, 5120, -256), (256, 128, 512, -256)])
def test_geglu_cta_pairs(K, m, k, n, block_n):
    x, w, b = rnd(m, k, seed=94), rnd(n, k, scale=k ** -0.5, seed=95), rnd(n, seed=96) * 0.1
    got = K.linear(x.cuda(), w.cuda(), b.cuda(), geglu=True, block_n=block_n)
    plain = K.linear(x.cuda(), w.cuda(), b.cuda(), geglu=True, block_n=128)
    assert torch.equal(got, plain)
    s, g = F.linear(x.float(), w.float(), b.float()).chunk(2, dim=-1)
    check(got.cpu(), s * F.gelu(g))


@pytest.mark.parametrize("n,c,k,hw,block_n", [(2, 320, 320, 128, -160), (2, 640, 640, 64, -160), (2, 1280, 1280, 32, -160),
                                              (2, 320, 320, 128, -256), (2, 640, 320, 128, 160), (4, 64, 64, 8, -160)])
def test_conv_tile_shapes_and_cta_pairs(K, n, c, k, hw, block_n):
    x, w, b = rnd(n, c, hw, hw, seed=21), rnd(k, c, 3, 3, scale=(9 * c) ** -0.5, seed=22), rnd(k, seed=23) * 0.1
    temb = rnd(n, k, seed=24).cuda()
    xg = x.cuda().contiguous(memory_format=torch.channels_last)
    plain = K.conv2d(xg, w.cuda(), b.cuda(), temb=temb, block_n=128)
    got, part = K.conv2d(xg, w.cuda(), b.cuda(), temb=temb, block_n=block_n, gn_stats=True)
    assert torch.equal(got, plain)
    check(got.cpu(), F.conv2d(x.float(), w.float(), b.float(), padding=1) + temb.float().cpu()[:, :, None, None])
    if part is not None:
        _check_partials(part, got.permute(0, 2, 3, 1).reshape(-1, k).float())


def test_launcher_tile_choice_is_reported(K):
    """st_debug_choose_tile: what the cost model picks (negative = CTA pair); the big multi-wave GEMMs must be paired."""
    from stabletriton_b200 import _cabi
    f = _cabi._load().st_debug_choose_tile
    f.restype = __import__("ctypes").c_int
    assert f(2048, 5120, 1, 1280) == -256 and f(8192, 8192, 0, 8192) == -256


def test_attention_polynomial_exp2_share(K):
    """ST_ATTN_POLY=2 instantiation of the pipelined kernel: a quarter of the exponentials evaluated by the FMA-pipe
    polynomial (Cody-Waite split + degree-3 minimax, relative error 9e-5) instead of MUFU.EX2 -- same parity bar."""
    from stabletriton_b200 import _cabi
    L = _cabi._load()
    b, h, tq, tk = 1, 3, 300, 1000
    q, k, v = rnd(b, tq, h * 64, seed=50) * 2.0, rnd(b, tk, h * 64, seed=51) * 2.0, rnd(b, tk, h * 64, seed=52)
    ref = O.attention_core(q.float(), k.float(), v.float(), h, 64)
    try:
        L.st_debug_set_attention_impl(2)
        L.st_debug_set_attention_parts(2)  # the polynomial share exists in the 8-exp-warp layout
        L.st_debug_set_attention_poly(0)
        base = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
        L.st_debug_set_attention_poly(2)
        poly = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
        torch.cuda.synchronize()
    finally:
        L.st_debug_set_attention_poly(-1)
        L.st_debug_set_attention_parts(0)
        L.st_debug_set_attention_impl(0)
    check(poly.cpu(), ref, rel_tol=2e-2, cos_tol=0.9995)
    rel, cos = parity(poly.float(), base.float())
    assert rel <= 8e-3 and cos >= 0.99999, (rel, cos)


@pytest.mark.parametrize("b,h,tq,tk", [(2, 10, 4096, 4096), (2, 20, 1024, 1024), (1, 2, 300, 1000), (1, 3, 129, 129)])
def test_attention_four_exp_warps_per_quadrant(K, b, h, tq, tk):
    """The 736-thread layout of the pipelined kernel (16 exp warps, one 32-column chunk each) against the oracle and
    against the 480-thread layout: both evaluate the same per-element arithmetic, so P and the row sums -- hence the
    output -- must agree to the last bit."""
    from stabletriton_b200 import _cabi
    L = _cabi._load()
    q, k, v = rnd(b, tq, h * 64, seed=40), rnd(b, tk, h * 64, seed=41), rnd(b, tk, h * 64, seed=42)
    ref = O.attention_core(q.float(), k.float(), v.float(), h, 64)
    try:
        L.st_debug_set_attention_impl(2)
        L.st_debug_set_attention_parts(2)
        two = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
        L.st_debug_set_attention_parts(4)
        four = K.attention_btc(q.cuda(), k.cuda(), v.cuda(), h, 0.125)
        torch.cuda.synchronize()
    finally:
        L.st_debug_set_attention_parts(0)
        L.st_debug_set_attention_impl(0)
    check(four.cpu(), ref, rel_tol=2e-2, cos_tol=0.9995)
    rel, cos = parity(four.float(), two.float())
    assert rel <= 4e-3, (rel, cos)  # the row sum is accumulated in a different order (4 partial sums instead of 2)
