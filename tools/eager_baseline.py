"""Same-box honest bar (not part of the bench contract): the SAME UNet definition run by stock PyTorch on the
B200 -- bf16, cuBLAS / cuDNN / SDPA flash attention, channels-last -- eager and under a CUDA graph, next to the compiled
engine.  Prints one JSON line per arm.

    python tools/eager_baseline.py [--batch 2] [--latent 128]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stabletriton_b200 as st  # noqa: E402
from stabletriton_b200 import UNetConfig, synth  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def sdpa_forward(self, hidden_states, encoder_hidden_states=None):
    """Attention.forward with the explicit softmax(QK^T)V replaced by torch SDPA (flash attention on B200) -- what a
    Diffusers user gets from AttnProcessor2_0."""
    context = hidden_states if encoder_hidden_states is None else encoder_hidden_states
    q, k, v = self.to_q(hidden_states), self.to_k(context), self.to_v(context)
    b, t, c = q.size()
    q = q.view(b, t, self.num_heads, self.head_dim).transpose(1, 2)
    k = k.view(b, k.size(1), self.num_heads, self.head_dim).transpose(1, 2)
    v = v.view(b, v.size(1), self.num_heads, self.head_dim).transpose(1, 2)
    out = torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, t, c)
    for layer in self.to_out:
        out = layer(out)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    cfg = UNetConfig.sdxl()
    model = synth.build_unet(cfg, seed=7)
    inp = synth.synth_inputs(args.batch, args.latent, cfg, device="cuda", dtype=torch.bfloat16)
    results = {}
    with torch.no_grad():
        eager = model.to(memory_format=torch.channels_last)
        ref = eager(**inp)[0].float()
        results["torch eager bf16, attention as written (softmax(QK^T)V)"] = timeit(lambda: eager(**inp), args.iters)
        from stabletriton_b200.unet import Attention
        plain_forward = Attention.forward
        Attention.forward = sdpa_forward
        results["torch eager bf16 (cuBLAS/cuDNN/SDPA)"] = timeit(lambda: eager(**inp), args.iters)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            eager(**inp)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                eager(**inp)
        torch.cuda.synchronize()
        results["torch eager bf16 (cuBLAS/cuDNN/SDPA) + CUDA graph"] = timeit(g.replay, args.iters)
        del g
        Attention.forward = plain_forward  # the fx passes match the reference's explicit pattern
        compiled = st.compile(model, cuda_graph=True)
        out = compiled(**inp)[0].float()
        results["stabletriton_b200.compile (CUDA graph)"] = timeit(lambda: compiled(**inp), args.iters)
    rel = ((out - ref).abs().max() / ref.abs().max()).item()
    for k, v in results.items():
        print(json.dumps({"arm": k, "ms_per_forward": round(v, 3), "it_per_s": round(1e3 / v, 2), "batch": args.batch,
                          "latent": args.latent}))
    print(json.dumps({"compiled_vs_torch_bf16_max_rel": rel}))


if __name__ == "__main__":
    main()
