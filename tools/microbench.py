"""BASELINE configs[2]: kernel micro-benchmarks at SDXL shapes, through the C ABI, written as data.

    python tools/microbench.py [--out profiles/r02_microbench.json]

Every case is captured into ONE CUDA graph that launches the kernel `rot` times, each launch on its own set of input /
output buffers, with `rot` chosen so that the buffers of one replay add up to more than twice the 126 MB L2: every
launch finds its operands L2-cold, as the timing rules ask ("use inputs larger than L2").  Reported per launch:
time (CUDA events around 10 replays / (10 * rot)), achieved GB/s (norms: algorithmic bytes = one read + one write,
SURVEY 8d) or TFLOP/s (attention, GEMM, conv), and the fraction of the measured peak (MEASURED_PEAKS.json).
A second figure, `warm`, replays one buffer set only (L2-resident operands: how the kernel behaves inside the step,
where its input was just written by the producer)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabletriton_b200 import kernels as K  # noqa: E402

BF16 = torch.bfloat16
L2_BYTES = 126e6


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), p["bf16_tflops"], "MEASURED_PEAKS.json"
    return 6650.0, 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def rand(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda", dtype=torch.float32) * scale).to(BF16)


def nhwc(n, c, h, w):
    return rand(n, h, w, c).permute(0, 3, 1, 2)


def time_graph(fns, iters=10):
    """fns: list of zero-argument callables, all launched inside one graph.  Returns ms per callable."""
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.no_grad():
        for f in fns:
            f()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for f in fns:
                f()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (iters * len(fns))


def bench_case(name, make, bytes_per_set, work, unit, peak):
    """make() -> a callable bound to a FRESH set of buffers.  work = algorithmic bytes or FLOPs per launch.
    Every case starts from an idle GPU (1 s pause): the first version of this table timed the convolutions right after
    the 0.8 ms 8192^3 GEMM replays and reported them at 0.42 of peak -- power-capped clocks, not the kernel
    (tools/conv_probe.py: the same launches run at 1000-1050 TFLOP/s from idle)."""
    import time
    torch.cuda.synchronize()
    time.sleep(1.0)
    rot = max(2, min(64, int(2.5 * L2_BYTES / max(bytes_per_set, 1)) + 1))
    cold = time_graph([make() for _ in range(rot)])
    one = make()
    warm = time_graph([one] * 8)
    scale = 1e9 if unit == "GB/s" else 1e12
    res = {"case": name, "unit": unit, "work_per_launch": work, "rotating_sets": rot,
           "us_cold": cold * 1e3, "achieved_cold": work / (cold * 1e-3) / scale, "frac_cold": work / (cold * 1e-3) / scale / peak,
           "us_warm": warm * 1e3, "achieved_warm": work / (warm * 1e-3) / scale, "frac_warm": work / (warm * 1e-3) / scale / peak}
    print(f"{name:58s} cold {res['us_cold']:8.1f} us {res['achieved_cold']:8.1f} {unit} ({res['frac_cold']:.2f})   "
          f"warm {res['us_warm']:8.1f} us {res['achieved_warm']:8.1f} ({res['frac_warm']:.2f})", flush=True)
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_microbench.json"))
    args = ap.parse_args()
    hbm, tf_sus, tf_burst, src = peaks()
    out = {"peaks": {"hbm_gbs": hbm, "bf16_tflops_sustained": tf_sus, "bf16_tflops_burst": tf_burst, "source": src},
           "method": __doc__.split("\n\n")[2].replace("\n", " "), "gpu": torch.cuda.get_device_name(0), "cases": []}
    add = out["cases"].append

    # ---- GroupNorm + SiLU, 32 groups (a2) --------------------------------------------------------------------------
    for n in (2, 16):
        for c, hw in ((320, 128), (640, 64), (1280, 32)):
            def make(n=n, c=c, hw=hw):
                x = nhwc(n, c, hw, hw)
                w, b = rand(c), rand(c)
                return lambda: K.groupnorm_wrapper(x, 32, w, b, 1e-5, True)
            nbytes = 2 * n * c * hw * hw * 2
            add(bench_case(f"groupnorm+silu N={n} C={c} {hw}x{hw}", make, nbytes, nbytes, "GB/s", hbm))
    # ---- LayerNorm (a3) -------------------------------------------------------------------------------------------------
    for m, n in ((8192, 640), (2048, 1280), (65536, 1280)):
        def make(m=m, n=n):
            x, w, b = rand(m, n), rand(n), rand(n)
            return lambda: K.layer_norm(x, w, b, 1e-5)
        nbytes = 2 * m * n * 2
        add(bench_case(f"layernorm M={m} N={n}", make, nbytes, nbytes, "GB/s", hbm))
    # ---- attention (a4): self T=4096 / 1024 (/16384 for 2048^2), cross Tk=77 ------------------------------------------
    for b, h, tq, tk in ((2, 10, 4096, 4096), (2, 20, 1024, 1024), (2, 10, 16384, 16384), (2, 10, 4096, 77), (2, 20, 1024, 77)):
        def make(b=b, h=h, tq=tq, tk=tk):
            q, k, v = rand(b, tq, h * 64), rand(b, tk, h * 64), rand(b, tk, h * 64)
            return lambda: K.attention_btc(q, k, v, h, 0.125)
        nbytes = (2 * b * tq * h * 64 + 2 * b * tk * h * 64) * 2
        flops = 4.0 * b * h * tq * tk * 64
        add(bench_case(f"attention B={b} H={h} Tq={tq} Tk={tk}", make, nbytes, flops, "TFLOP/s", tf_sus))
    # ---- Linear / GEGLU (a5, a6) ------------------------------------------------------------------------------------------
    for m, n, k, geglu, res in ((2048, 10240, 1280, True, False), (2048, 1280, 5120, False, True),
                                (2048, 1280, 1280, False, True), (2048, 3840, 1280, False, False),
                                (8192, 5120, 640, True, False), (8192, 640, 640, False, True), (8192, 8192, 8192, False, False)):
        def make(m=m, n=n, k=k, geglu=geglu, res=res):
            x, w, bias = rand(m, k), rand(n, k, scale=k ** -0.5), rand(n)
            r = rand(m, n) if res else None
            return lambda: K.linear(x, w, bias, residual=r, geglu=geglu, w_static=True)
        nbytes = (m * k + n * k + m * (n // 2 if geglu else n) * (2 if res else 1)) * 2
        tag = " GEGLU" if geglu else (" +bias+residual" if res else " +bias")
        add(bench_case(f"linear M={m} N={n} K={k}{tag}", make, nbytes, 2.0 * m * n * k, "TFLOP/s", tf_sus))
    # ---- conv3x3 implicit GEMM (a7) ---------------------------------------------------------------------------------------
    for n, c, hw, kk in ((2, 320, 128, 320), (2, 640, 64, 640), (2, 1280, 32, 1280), (2, 2560, 32, 1280)):
        def make(n=n, c=c, hw=hw, kk=kk):
            x = nhwc(n, c, hw, hw)
            w = K.pack_conv_weight(rand(kk, c, 3, 3, scale=(9 * c) ** -0.5))
            bias, r = rand(kk), nhwc(n, kk, hw, hw)
            return lambda: K.conv2d(x, w, bias, residual=r, w_static=True)
        nbytes = (n * c * hw * hw + kk * c * 9 + 2 * n * kk * hw * hw) * 2
        add(bench_case(f"conv3x3 N={n} C={c} {hw}x{hw} K={kk} +bias+residual", make, nbytes,
                       2.0 * n * hw * hw * kk * c * 9, "TFLOP/s", tf_sus))

        def make_temb(n=n, c=c, hw=hw, kk=kk):  # resnet conv1: + bias + time-embedding row, GroupNorm partials emitted
            x = nhwc(n, c, hw, hw)
            w = K.pack_conv_weight(rand(kk, c, 3, 3, scale=(9 * c) ** -0.5))
            bias, temb = rand(kk), rand(n, kk)
            return lambda: K.conv2d(x, w, bias, temb=temb, w_static=True, gn_stats=True)
        add(bench_case(f"conv3x3 N={n} C={c} {hw}x{hw} K={kk} +bias+temb+gn-partials", make_temb, nbytes,
                       2.0 * n * hw * hw * kk * c * 9, "TFLOP/s", tf_sus))

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
