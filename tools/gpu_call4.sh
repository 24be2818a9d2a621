#!/bin/bash
# round-2 GPU call 4: pair policy A/B in sequence; 16-exp-warp attention
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or tile or pairs" ) > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2d.log
S=stabletriton_b200/csrc/selftest
{
  for parts in 2 4; do
    for shape in "2 10 4096 4096" "2 20 1024 1024" "2 10 16384 16384" "2 10 1000 1000"; do
      echo "== ST_ATTN_PARTS=$parts attn1 $shape"; ST_ATTN_PARTS=$parts timeout 120 $S attn1 $shape | grep -E "attention|FAIL|PASS" | tail -2
    done
  done
} > gpurun_out/attn_parts_r2d.log 2>&1
for pol in 0 1 2; do ST_GEMM_CLUSTER=$pol timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2d_cluster$pol.log 2>&1; done
ST_ATTN_PARTS=4 timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2d_parts4.log 2>&1
echo done
