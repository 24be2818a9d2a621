#!/bin/bash
# round-2 GPU call 60: final check of the tree -- whole GPU suite, smoke, lean bench line
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > $O/pytest_gpu_r2bf.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2bf.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2bf.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2bf.log
( time timeout 1200 python bench.py --lean ) > $O/bench_r2_v6_lean.json 2> $O/bench_r2_v6_lean.err; echo "bench rc=$?" >> $O/bench_r2_v6_lean.err
echo done
