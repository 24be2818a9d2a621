#!/bin/bash
# round-2 GPU call 52: power-capped regime -- does parking the GEMM epilogue warps (suspend-time hint on the accumulator wait)
# raise the sustained clocks?  bench.py --lean with a 300-step timed region (5 s), same box, alternating
mkdir -p gpurun_out
O=gpurun_out
for pass in a b; do
  for hint in 0 2000 20000; do
    ST_GEMM_EPI_HINT=$hint timeout 600 python bench.py --lean --steps 300 --warmup 10 > $O/bench_r2az_hint${hint}_$pass.json 2> $O/bench_r2az_hint${hint}_$pass.err
  done
done
echo done
