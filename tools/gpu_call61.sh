#!/bin/bash
# round-2 GPU call 61: full bench line of the final tree (strong-scaling arm now on the four-CTA resident attention form)
mkdir -p gpurun_out
( time timeout 1200 python bench.py ) > gpurun_out/bench_r2_v6.json 2> gpurun_out/bench_r2_v6.err; echo "bench rc=$?" >> gpurun_out/bench_r2_v6.err
echo done
