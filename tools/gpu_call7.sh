#!/bin/bash
# round-2 GPU call 7: short-context (Tk <= 80) attention kernel -- parity, micro timing, whole step
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or overrun" ) > $O/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2g.log
S=stabletriton_b200/csrc/selftest
{
  for impl in short 2cta; do
    for shape in "2 20 1024 77" "2 10 4096 77" "16 20 1024 77"; do
      echo "== ST_ATTN_IMPL=$impl attn1 $shape"; ST_ATTN_IMPL=$impl timeout 120 $S attn1 $shape | grep -E "attention|FAIL|PASS" | tail -1
    done
  done
} > $O/attn_short_r2g.log 2>&1
timeout 600 python tools/quick_bench.py > $O/qb_r2g_short.log 2>&1
ST_ATTN_IMPL=2cta timeout 600 python tools/quick_bench.py > $O/qb_r2g_2cta.log 2>&1
echo done
