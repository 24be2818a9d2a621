#!/bin/bash
# round-2 GPU call 5: evidence run of the current tree -- parity suite, smoke, bench (both arms), ncu launch list,
# CUPTI timeline, microbench, 2048^2 batch sweep
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > $O/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2e.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2e.log
( time timeout 1200 python bench.py ) > $O/bench_r2_v2.json 2> $O/bench_r2_v2.err; echo "bench rc=$?" >> $O/bench_r2_v2.err
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_r2_ref_v2.json 2> $O/bench_r2_ref_v2.err
timeout 600 python tools/quick_bench.py > $O/qb_r2e.log 2>&1
timeout 600 python tools/timeline_probe.py $O/r02_timeline_v1.json > $O/r02_timeline_v1.txt 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $O/r02_ncu_launches_v1.csv python tools/one_forward.py > $O/ncu_of_r2e.log 2>&1
python tools/ncu_launch_summary.py $O/r02_ncu_launches_v1.csv $O/r02_ncu_launch_summary_v1.json > $O/r02_ncu_launch_summary_v1.txt 2>&1
timeout 900 python tools/microbench.py --out $O/r02_microbench_v2.json > $O/microbench_r2e.log 2>&1
for p in 1 2 4 8; do
  timeout 600 python bench.py --lean --latent 256 --prompts $p --steps 10 --warmup 3 > $O/r02_2048px_p$p.json 2> $O/r02_2048px_p$p.err
done
echo done
