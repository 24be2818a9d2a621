#!/bin/bash
# round-2 GPU call 45: same-box A/B of the two late changes at other batch sizes (B = 1: the CFG-split rank; B = 16: config 4)
mkdir -p gpurun_out
O=gpurun_out
for b in 1 16; do
  for pass in a b; do
    ST_GEMM_RES_TMA=0 ST_ATTN_IMPL=noresident timeout 600 python tools/quick_bench.py --batch $b > $O/qb_r2as_b${b}_old_$pass.log 2>&1
    ST_GEMM_RES_TMA=0 timeout 600 python tools/quick_bench.py --batch $b > $O/qb_r2as_b${b}_resattn_$pass.log 2>&1
    timeout 600 python tools/quick_bench.py --batch $b > $O/qb_r2as_b${b}_new_$pass.log 2>&1
  done
done
echo done
