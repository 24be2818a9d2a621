#!/bin/bash
# round-2 GPU call 17: MUFU-token ping-pong between the two exp warps of a quadrant
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for pp in 0 1; do
    for shape in "2 10 4096 4096" "2 20 1024 1024" "2 10 16384 16384" "2 10 1000 1000" "1 3 300 333"; do
      echo "== ST_ATTN_PINGPONG=$pp attn1 $shape"; ST_ATTN_PINGPONG=$pp timeout 120 $S attn1 $shape | grep -E "attention" | tail -1
    done
  done
  echo "== trace, pingpong"; ST_ATTN_PINGPONG=1 timeout 120 $S attn1 2 10 4096 4096 | grep -v "^device"
} > $O/attn_pingpong_r2q.log 2>&1
( ST_ATTN_PINGPONG=1 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or overrun" ) > $O/pytest_gpu_r2q.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2q.log
ST_ATTN_PINGPONG=1 timeout 600 python tools/quick_bench.py > $O/qb_r2q_pp1.log 2>&1
ST_ATTN_PINGPONG=0 timeout 600 python tools/quick_bench.py > $O/qb_r2q_pp0.log 2>&1
echo done
