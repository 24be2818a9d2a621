#!/bin/bash
# round-2 GPU call 1: parity suite, smoke, bench, family table, pair-kernel diagnostics, microbench, sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpuinfo.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2a.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r2a.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_v0.json 2> gpurun_out/bench_r2_v0.err; echo "bench rc=$?" >> gpurun_out/bench_r2_v0.err
timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2_base.log 2>&1
S=stabletriton_b200/csrc/selftest
{
  echo "== GEGLU 2048x10240x1280 plain"; $S gemm1 2048 10240 1280 6 0 1 0; $S trace 2048 10240 1280 6 256
  echo "== GEGLU pair"; ST_GEMM_CLUSTER=1 $S gemm1 2048 10240 1280 6 0 1 0; ST_GEMM_CLUSTER=1 $S trace 2048 10240 1280 6 256
  echo "== 2048x1280x5120 plain"; $S gemm1 2048 1280 5120 4 0 1 1; $S trace 2048 1280 5120 4 192
  echo "== 2048x1280x5120 pair 256"; ST_GEMM_CLUSTER=1 $S gemm1 2048 1280 5120 4 256 1 1; ST_GEMM_CLUSTER=1 $S trace 2048 1280 5120 4 256
  echo "== 2048x3840x1280 plain"; $S gemm1 2048 3840 1280 4 0 0 0
  echo "== 2048x3840x1280 pair"; ST_GEMM_CLUSTER=1 $S gemm1 2048 3840 1280 4 256 0 0
  echo "== 8192^3 plain"; $S gemm1 8192 8192 8192 4 0 0 0
  echo "== 8192^3 pair"; ST_GEMM_CLUSTER=1 $S gemm1 8192 8192 8192 4 256 0 0
} > gpurun_out/pair_diag_r2a.log 2>&1
timeout 900 python tools/microbench.py --out gpurun_out/r02_microbench_v0.json > gpurun_out/microbench_r2a.log 2>&1
timeout 1500 tools/sanitize.sh gpurun_out/sanitizer_r2a > gpurun_out/sanitize_r2a.log 2>&1
echo done
