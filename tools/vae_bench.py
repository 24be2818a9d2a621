"""VAE decode timing: the SDXL VAE decoder on the sm_100a kernels, latents resident in HBM, one CUDA-graph
replay per image batch; beside it the same module through stock PyTorch (bf16, channels-last, cuDNN / cuBLAS / SDPA-math).

    python tools/vae_bench.py [--latent 128] [--batch 1] [--out profiles/r02_vae_decode.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabletriton_b200 import _cabi  # noqa: E402
from stabletriton_b200.vae import VAEConfig, build_vae_decoder, compile_vae  # noqa: E402


def decoder_flops(cfg: VAEConfig, latent: int, batch: int) -> float:
    ch = tuple(reversed(cfg.block_out_channels))
    conv = lambda hw, cin, cout, k=9: 2.0 * hw * hw * cin * cout * k  # noqa: E731
    res = lambda hw, cin, cout: conv(hw, cin, cout) + conv(hw, cout, cout) + (conv(hw, cin, cout, 1) if cin != cout else 0.0)  # noqa: E731
    hw, c0 = latent, ch[0]
    f = conv(hw, cfg.latent_channels, c0) + 2 * res(hw, c0, c0)
    t = hw * hw
    f += 2.0 * t * c0 * c0 * 4 + 4.0 * t * t * c0  # q, k, v, out projections + Q K^T + P V
    prev = c0
    for i, c in enumerate(ch):
        f += res(hw, prev, c) + cfg.layers_per_block * res(hw, c, c)
        prev = c
        if i < len(ch) - 1:
            hw *= 2
            f += conv(hw, c, c)
    f += conv(hw, ch[-1], cfg.out_channels)
    return f * batch


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_vae_decode.json"))
    args = ap.parse_args()
    cfg = VAEConfig.sdxl()
    model = build_vae_decoder(cfg, seed=11)
    z = (torch.randn(args.batch, 4, args.latent, args.latent, device="cuda") * 0.5).to(torch.bfloat16)
    vae = compile_vae(model)
    before = _cabi.launch_count()
    vae.eager_decode(z)
    torch.cuda.synchronize()
    launches = _cabi.launch_count() - before
    img = vae.decode(z)
    ms = time_ms(lambda: vae.decode(z))
    flops = decoder_flops(cfg, args.latent, args.batch)
    # stock PyTorch on the same module (weights already channels-last for the 3x3 convs)
    zt = z.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        ref = model(zt)
        ms_torch = time_ms(lambda: model(zt), iters=5, warm=2)
    d = (img.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    out = {
        "what": "SDXL VAE decode (Diffusers AutoencoderKL decoder, 49.5 M params, synthetic weights), bf16, latents resident in HBM",
        "latent": args.latent, "image": 8 * args.latent, "batch": args.batch,
        "ms_per_decode": ms, "images_per_s": args.batch / (ms * 1e-3), "tflop": flops / 1e12,
        "tflops_achieved": flops / (ms * 1e-3) / 1e12, "kernel_launches": int(launches),
        "torch_eager_bf16_ms": ms_torch, "speedup_vs_torch_eager": ms_torch / ms,
        "max_rel_diff_vs_torch_bf16": d, "gpu": torch.cuda.get_device_name(0),
    }
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
