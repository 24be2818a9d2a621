"""Developer probe: does splitting the N range of a multi-round GEMM into a part that fills whole rounds of CTA pairs and a
remainder with its own tile width beat one launch?  (QKV projection 2048 x 3840 x 1280: 120 pair tiles on 74 clusters.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stabletriton_b200 import kernels as K  # noqa: E402
from tools.quick_bench import time_graph  # noqa: E402


def main():
    torch.manual_seed(0)
    m, n, k = 2048, 3840, 1280
    x = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
    ws = [torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * k ** -0.5 for _ in range(8)]  # rotate: weights never L2-hot
    b = torch.randn(n, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)

    def whole(bn=0):
        for w in ws:
            K.linear(x, w, b, out=out, block_n=bn, w_static=True)

    def split(n1, bn1, bn2):
        for w in ws:
            K.linear(x, w[:n1], b[:n1], out=out[:, :n1], block_n=bn1, w_static=True)
            K.linear(x, w[n1:], b[n1:], out=out[:, n1:], block_n=bn2, w_static=True)

    ref = None
    for name, fn in [("one launch, model's tile", lambda: whole(0)), ("one launch, pair 256", lambda: whole(-256)),
                     ("one launch, 256", lambda: whole(256)), ("one launch, pair 192", lambda: whole(-192)),
                     ("2304 pair 256 + 1536 x 192", lambda: split(2304, -256, 192)),
                     ("2304 pair 256 + 1536 x 256", lambda: split(2304, -256, 256)),
                     ("2304 pair 256 + 1536 x 128", lambda: split(2304, -256, 128)),
                     ("2304 pair 256 + 1536 pair 192", lambda: split(2304, -256, -192)),
                     ("2304 pair 256 + 1536 x 160", lambda: split(2304, -256, 160)),
                     ("2560 pair 256 + 1280 x 160", lambda: split(2560, -256, 160)),
                     ("2048 pair 256 + 1792 pair 256", lambda: split(2048, -256, -256))]:
        ms = time_graph(fn, iters=5) / len(ws)
        fn()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        same = torch.equal(ref, out)
        print(f"{name:34s} {ms * 1e3:7.2f} us  {2.0 * m * n * k / ms * 1e-9:7.1f} TFLOP/s  bit-identical to the first: {same}")


if __name__ == "__main__":
    main()
