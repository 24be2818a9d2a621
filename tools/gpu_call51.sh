#!/bin/bash
# round-2 GPU call 51: evidence run of the final tree (resident attention kernel, VAE path) -- parity suite, smoke, bench
# (both arms), family table, CUPTI timeline, ncu launch list, ncu --set full of the resident attention kernel, VAE bench
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > $O/pytest_gpu_r2ay.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2ay.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2ay.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2ay.log
( time timeout 1200 python bench.py ) > $O/bench_r2_v5.json 2> $O/bench_r2_v5.err; echo "bench rc=$?" >> $O/bench_r2_v5.err
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_r2_ref_v5.json 2> $O/bench_r2_ref_v5.err
timeout 600 python tools/quick_bench.py > $O/qb_r2ay.log 2>&1
timeout 600 python tools/timeline_probe.py $O/r02_timeline_v4.json > $O/r02_timeline_v4.txt 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $O/r02_ncu_launches_v4.csv python tools/one_forward.py > $O/ncu_of_r2ay.log 2>&1
python tools/ncu_launch_summary.py $O/r02_ncu_launches_v4.csv $O/r02_ncu_launch_summary_v4.json > $O/r02_ncu_launch_summary_v4.txt 2>&1
echo done
