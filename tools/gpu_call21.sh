#!/bin/bash
# round-2 GPU call 21: GEMM exit waits for the bulk stores' smem reads only (ST_GEMM_DRAIN=1 = old behaviour)
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for d in 1 0; do
    for shape in "2048 1280 1280 4 0 1 1" "2048 1280 5120 4 0 1 1" "8192 640 640 4 0 1 1" "2048 3840 1280 4 0 0 0" "2048 10240 1280 6 0 1 0"; do
      echo "== ST_GEMM_DRAIN=$d gemm1 $shape"; ST_GEMM_DRAIN=$d timeout 60 $S gemm1 $shape | grep -E "gemm M" | tail -1
    done
  done
} > $O/gemm_drain_r2u.log 2>&1
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q ) > $O/pytest_gpu_r2u.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2u.log
for pass in a b; do
  ST_GEMM_DRAIN=1 timeout 600 python tools/quick_bench.py > $O/qb_r2u_drain1_$pass.log 2>&1
  ST_GEMM_DRAIN=0 timeout 600 python tools/quick_bench.py > $O/qb_r2u_drain0_$pass.log 2>&1
done
echo done
