"""Probe: the CFG pair as ONE batch-2 forward vs TWO concurrent batch-1 forwards (one per CUDA-graph branch).

At batch 2 most kernels of the SDXL step are latency-bound (80-112 CTAs on 148 SMs, ~9 us of fixed cost per launch);
two independent batch-1 chains can fill each other's gaps.  Rows of a batch are independent in every op of the UNet
(GroupNorm is per image, LayerNorm per token, attention per (image, head)), so the result is the same."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stabletriton_b200 as st  # noqa: E402
from stabletriton_b200 import UNetConfig, synth  # noqa: E402


def timed(graph, iters=10):
    for _ in range(3):
        graph.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    latent = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    cfg = UNetConfig.sdxl()
    model = synth.build_unet(cfg, seed=7)
    compiled = st.compile(model, cuda_graph=False)
    inp = synth.synth_inputs(batch, latent, cfg, device="cuda", dtype=torch.bfloat16)

    def rows(lo, hi):
        return dict(sample=inp["sample"][lo:hi].contiguous(), timesteps=inp["timesteps"],
                    encoder_hidden_states=inp["encoder_hidden_states"][lo:hi].contiguous(),
                    added_cond_kwargs={k: v[lo:hi].contiguous() for k, v in inp["added_cond_kwargs"].items()})

    halves = [rows(0, batch // 2), rows(batch // 2, batch)]
    main_s, side_s = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.no_grad(), torch.cuda.stream(main_s):
        ref = compiled(**inp)[0]
        for h in halves:
            compiled(**h)
        torch.cuda.synchronize()
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, stream=main_s):
            out1 = compiled(**inp)[0]
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, stream=main_s):
            side_s.wait_stream(main_s)
            a = compiled(**halves[0])[0]
            with torch.cuda.stream(side_s):
                b = compiled(**halves[1])[0]
            main_s.wait_stream(side_s)
            out2 = torch.cat([a, b], dim=0)
        g3 = torch.cuda.CUDAGraph()  # the two halves one after the other on one stream
        with torch.cuda.graph(g3, stream=main_s):
            a3 = compiled(**halves[0])[0]
            b3 = compiled(**halves[1])[0]
    torch.cuda.synchronize()
    t1, t2, t3 = timed(g1), timed(g2), timed(g3)
    d = (out2.float() - out1.float()).abs().max().item()
    print(f"batch {batch} latent {latent}: one batch-{batch} forward {t1:.3f} ms | two concurrent batch-{batch // 2} "
          f"forwards {t2:.3f} ms | the same two back to back {t3:.3f} ms | max|diff| {d:.3e} "
          f"(max|ref| {ref.float().abs().max().item():.3f})")


if __name__ == "__main__":
    main()
