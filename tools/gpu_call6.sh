#!/bin/bash
# round-2 GPU call 6: polynomial exp2 share sweep in the pipelined attention kernel
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for poly in 0 1 2 3 4; do
    for shape in "2 10 4096 4096" "2 20 1024 1024" "2 10 16384 16384"; do
      echo "== ST_ATTN_POLY=$poly attn1 $shape"; ST_ATTN_POLY=$poly timeout 120 $S attn1 $shape | grep -E "attention|FAIL|PASS" | tail -1
    done
  done
} > gpurun_out/attn_poly_r2f.log 2>&1
for poly in 0 2 3; do ST_ATTN_POLY=$poly timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2f_poly$poly.log 2>&1; done
ST_ATTN_POLY=2 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" > gpurun_out/pytest_gpu_r2f.log 2>&1
echo done
