"""Why does tools/microbench.py time the 128x128 / 64x64 residual convolutions at ~600 TFLOP/s when selftest conv1 and the
in-step replays give ~1000?  Times K.conv2d variants of one shape with the microbench harness."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stabletriton_b200 import kernels as K  # noqa: E402
from tools.microbench import nhwc, rand, time_graph  # noqa: E402


def main():
    for n, c, hw, kk in ((2, 320, 128, 320), (2, 640, 64, 640)):
        flops = 2.0 * n * hw * hw * kk * c * 9
        x = nhwc(n, c, hw, hw)
        w = K.pack_conv_weight(rand(kk, c, 3, 3, scale=(9 * c) ** -0.5))
        bias, r, temb = rand(kk), nhwc(n, kk, hw, hw), rand(n, kk)
        r_small = (nhwc(n, kk, hw, hw).float() * 0.01).to(torch.bfloat16)
        cases = {
            "plain": lambda: K.conv2d(x, w, bias, w_static=True),
            "temb": lambda: K.conv2d(x, w, bias, temb=temb, w_static=True),
            "residual": lambda: K.conv2d(x, w, bias, residual=r, w_static=True),
            "residual (values x 0.01)": lambda: K.conv2d(x, w, bias, residual=r_small, w_static=True),
            "residual + gn partials": lambda: K.conv2d(x, w, bias, residual=r, w_static=True, gn_stats=True),
            "temb + gn partials": lambda: K.conv2d(x, w, bias, temb=temb, w_static=True, gn_stats=True),
            "residual, x = zeros": lambda z=torch.zeros_like(x): K.conv2d(z, w, bias, residual=r, w_static=True),
        }
        for name, fn in cases.items():
            ms = time_graph([fn] * 8)
            print(f"conv {n}x{c}x{hw}x{hw}->{kk} {name:28s} {ms * 1e3:7.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
