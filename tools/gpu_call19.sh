#!/bin/bash
# round-2 GPU call 19: exp-warp layout in the whole step, same box: 8 exp warps / 16 / 16 + MUFU token ring (two passes)
mkdir -p gpurun_out
O=gpurun_out
for pass in a b; do
  ST_ATTN_PARTS=2 timeout 600 python tools/quick_bench.py > $O/qb_r2s_parts2_$pass.log 2>&1
  ST_ATTN_PARTS=4 timeout 600 python tools/quick_bench.py > $O/qb_r2s_parts4_$pass.log 2>&1
  ST_ATTN_PARTS=4 ST_ATTN_RING=1 timeout 600 python tools/quick_bench.py > $O/qb_r2s_ring_$pass.log 2>&1
done
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or overrun" ) > $O/pytest_gpu_r2s.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2s.log
echo done
