"""Developer timing probe (not the contract bench): graphed SDXL UNet forward at a given batch/latent,
plus a per-kernel-family breakdown obtained by re-issuing the recorded launches of one forward,
family by family, inside their own CUDA graphs."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stabletriton_b200 as st  # noqa: E402
from stabletriton_b200 import UNetConfig, _cabi, synth  # noqa: E402


def time_graph(fn, iters=5):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    for _ in range(2):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--leave-one-out", action="store_true",
                    help="in-sequence marginal cost of each kernel family: time the whole launch sequence with that "
                         "family's launches left out (outputs are garbage, timing is not)")
    args = ap.parse_args()
    cfg = UNetConfig.tiny() if args.tiny else UNetConfig.sdxl()
    t0 = time.time()
    model = synth.build_unet(cfg, seed=7)
    print(f"model built in {time.time() - t0:.1f}s")
    t0 = time.time()
    compiled = st.compile(model, cuda_graph=False)
    print(f"compiled in {time.time() - t0:.1f}s; passes: {compiled.pass_report}")
    inp = synth.synth_inputs(args.batch, args.latent, cfg, device="cuda", dtype=torch.bfloat16)

    # Capture one forward into a graph that is kept alive: its private pool pins every activation
    # address, so the recorded C-ABI calls can be re-issued later, family by family.
    keep = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(side):
        compiled(**inp)
        torch.cuda.synchronize()
        with torch.cuda.graph(keep, stream=side):
            _cabi.start_recording()
            out = compiled(**inp)[0]  # noqa: F841
            calls = _cabi.stop_recording()
    torch.cuda.synchronize()
    for _ in range(3):
        keep.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        keep.replay()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"UNet forward B={args.batch} {args.latent}x{args.latent}: {ms:.3f} ms  ({1000.0 / ms:.2f} it/s), "
          f"{len(calls)} C-ABI calls per forward")

    fam = {}
    for name, a in calls:
        key = name
        if name == "st_gemm_bf16":
            key += f" M{a[6]} N{a[7]} K{a[8]} f{a[12]}" + (" +res" if a[10] else "")
        elif name == "st_conv3x3_nhwc_bf16":
            key += f" N{a[4]} {a[5]}x{a[6]} C{a[7]} K{a[8]}"
        elif name == "st_attention_bf16":
            key += f" B{a[16]} H{a[17]} Tq{a[18]} Tk{a[19]}"
        elif name == "st_groupnorm_nhwc_bf16":
            key += f" N{a[5]} HW{a[6]} C{a[7]}"
        elif name == "st_layernorm_bf16":
            key += f" M{a[6]} N{a[7]}"
        fam.setdefault(key, []).append((name, a))
    if args.leave_one_out:
        def group(name, a):
            if name == "st_gemm_bf16":
                m, n, k = a[6], a[7], a[8]
                if m == 2048 and n == 1280 and k == 1280:
                    return "gemm 2048x1280x1280"
                if m == 2048 and k == 5120:
                    return "gemm 2048x1280x5120 (ff out)"
                if m == 2048 and n == 3840:
                    return "gemm 2048x3840x1280 (qkv)"
                if m == 2048 and n == 10240:
                    return "gemm 2048x10240x1280 (geglu)"
                if m == 8192:
                    return "gemm M=8192 (64x64 level)"
                return "gemm other"
            if name == "st_conv3x3_nhwc_bf16":
                return "conv3x3"
            if name == "st_attention_bf16":
                return f"attention Tq{a[18]} Tk{a[19]}"
            if name == "st_groupnorm_nhwc_bf16":
                return "groupnorm"
            if name == "st_layernorm_bf16":
                return "layernorm"
            return "glue"
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            full = time_graph(lambda: _cabi.replay(calls, torch.cuda.current_stream().cuda_stream), iters=10)
        print(f"  all {len(calls)} launches replayed in sequence: {full:.3f} ms")
        groups = sorted({group(n, a) for n, a in calls})
        for gname in groups:
            rest = [(n, a) for n, a in calls if group(n, a) != gname]
            only = [(n, a) for n, a in calls if group(n, a) == gname]
            with torch.cuda.stream(stream):
                t_rest = time_graph(lambda: _cabi.replay(rest, torch.cuda.current_stream().cuda_stream), iters=10)
                t_only = time_graph(lambda: _cabi.replay(only, torch.cuda.current_stream().cuda_stream), iters=10)
            print(f"  {gname:40s} launches={len(only):4d}  in sequence {full - t_rest:7.3f} ms   alone {t_only:7.3f} ms")
        return

    stream = torch.cuda.Stream()
    total = 0.0
    for name, lst in sorted(fam.items()):
        with torch.cuda.stream(stream):
            t = time_graph(lambda: _cabi.replay(lst, torch.cuda.current_stream().cuda_stream))
        total += t
        print(f"  {name:62s} calls={len(lst):4d}  {t:8.3f} ms")
    print(f"  sum of families {total:.3f} ms")


if __name__ == "__main__":
    main()
