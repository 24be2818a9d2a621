#!/bin/bash
# round-2 GPU call 2: fused GroupNorm statistics -- parity, family table, whole step; pair kernel in sequence
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2b.log
timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2_gnfused.log 2>&1
ST_GEMM_CLUSTER=1 timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2_gnfused_pair.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_v1.json 2> gpurun_out/bench_r2_v1.err; echo "bench rc=$?" >> gpurun_out/bench_r2_v1.err
timeout 900 python tools/microbench.py --out gpurun_out/r02_microbench_v1.json > gpurun_out/microbench_r2b.log 2>&1
echo done
