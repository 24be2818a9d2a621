#!/bin/bash
# round-2 GPU call 11: per-warp mbarrier arrivals in the attention kernels
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or overrun" ) > $O/pytest_gpu_r2k.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2k.log
{
  for shape in "2 10 4096 4096" "2 20 1024 1024" "2 10 16384 16384" "2 10 1000 1000" "2 20 1024 77" "2 10 4096 77" "2 20 1024 100"; do
    echo "== attn1 $shape"; timeout 120 $S attn1 $shape | grep -E "attention" | tail -1
  done
  for parts in 4; do
    for shape in "2 10 4096 4096" "2 20 1024 1024"; do
      echo "== ST_ATTN_PARTS=$parts attn1 $shape"; ST_ATTN_PARTS=$parts timeout 120 $S attn1 $shape | grep -E "attention" | tail -1
    done
  done
  for poly in 2; do
    for shape in "2 10 4096 4096" "2 20 1024 1024"; do
      echo "== ST_ATTN_POLY=$poly attn1 $shape"; ST_ATTN_POLY=$poly timeout 120 $S attn1 $shape | grep -E "attention" | tail -1
    done
  done
  for ab in 1 4 8 13 63; do
    echo "== ST_ATTN_ABLATE=$ab attn1 2 10 4096 4096"; ST_ATTN_ABLATE=$ab timeout 120 $S attn1 2 10 4096 4096 | grep -E "attention" | tail -1 | sed 's/.*nan=[0-9]* //'
  done
} > $O/attn_warp_arrive_r2k.log 2>&1
timeout 600 python tools/quick_bench.py > $O/qb_r2k.log 2>&1
echo done
