#!/bin/bash
# round-2 GPU call 38: resident attention kernel by default -- attention tests, whole-step A/B on one box
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" ) > $O/pytest_gpu_r2al_attn.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2al_attn.log
for pass in a b; do
  ST_ATTN_IMPL=pipelined timeout 600 python tools/quick_bench.py > $O/qb_r2al_pipelined_$pass.log 2>&1
  timeout 600 python tools/quick_bench.py > $O/qb_r2al_resident_$pass.log 2>&1
done
echo done
