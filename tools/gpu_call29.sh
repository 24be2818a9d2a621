#!/bin/bash
# round-2 GPU call 29: VAE decode path -- new kernels, tiny + full SDXL decoder parity, timing
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_vae.py tests/test_cabi.py -m gpu -x -q -s ) > $O/pytest_gpu_r2ac_vae.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2ac_vae.log
timeout 600 python tools/vae_bench.py --out $O/r02_vae_decode.json > $O/vae_bench_r2ac.log 2>&1; echo "rc=$?" >> $O/vae_bench_r2ac.log
echo done
