#!/bin/bash
# round-2 GPU call 18: 16 exp warps + MUFU token ring (ST_ATTN_PARTS=4 ST_ATTN_PINGPONG=1)
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for cfg in "2 0" "4 0" "4 1"; do
    set -- $cfg
    for shape in "2 10 4096 4096" "2 20 1024 1024" "2 10 16384 16384" "2 10 1000 1000" "1 3 300 300"; do
      echo "== PARTS=$1 PINGPONG=$2 attn1 $shape"; ST_ATTN_PARTS=$1 ST_ATTN_PINGPONG=$2 timeout 60 $S attn1 $shape | grep -E "attention" | tail -1
    done
  done
} > $O/attn_ring_r2r.log 2>&1
( ST_ATTN_PARTS=4 ST_ATTN_PINGPONG=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or overrun" ) > $O/pytest_gpu_r2r.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2r.log
ST_ATTN_PARTS=4 ST_ATTN_PINGPONG=1 timeout 600 python tools/quick_bench.py > $O/qb_r2r_ring.log 2>&1
echo done
