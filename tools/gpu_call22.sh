#!/bin/bash
# round-2 GPU call 22: evidence run of the final tree -- parity suite, smoke, bench, family table, CUPTI timeline,
# microbench, ncu launch list, ncu --set full of the top kernels (selftest single launches)
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
( time timeout 1500 python -m pytest tests -m gpu -x -q -s ) > $O/pytest_gpu_r2v.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2v.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2v.log 2>&1; echo "smoke rc=$?" >> $O/smoke_r2v.log
( time timeout 1200 python bench.py ) > $O/bench_r2_v3.json 2> $O/bench_r2_v3.err; echo "bench rc=$?" >> $O/bench_r2_v3.err
timeout 600 python tools/quick_bench.py > $O/qb_r2v.log 2>&1
timeout 600 python tools/timeline_probe.py $O/r02_timeline_v2.json > $O/r02_timeline_v2.txt 2>&1
timeout 900 python tools/microbench.py --out $O/r02_microbench_v3.json > $O/microbench_r2v.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file $O/r02_ncu_launches_v2.csv python tools/one_forward.py > $O/ncu_of_r2v.log 2>&1
python tools/ncu_launch_summary.py $O/r02_ncu_launches_v2.csv $O/r02_ncu_launch_summary_v2.json > $O/r02_ncu_launch_summary_v2.txt 2>&1
timeout 300 $S attn1 2 10 4096 4096 > $O/plain_attn.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_pipelined -s 1 -c 1 -f -o $O/r02_ncu_attn_t4096 $S attn1 2 10 4096 4096 > $O/ncu_attn.log 2>&1
timeout 300 $S attn1 2 20 1024 77 > $O/plain_attn_short.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_short -s 1 -c 1 -f -o $O/r02_ncu_attn_short $S attn1 2 20 1024 77 > $O/ncu_attn_short.log 2>&1
timeout 300 $S gemm1 2048 10240 1280 6 0 1 0 > $O/plain_gemm.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 1 -c 1 -f -o $O/r02_ncu_gemm_geglu $S gemm1 2048 10240 1280 6 0 1 0 > $O/ncu_gemm.log 2>&1
timeout 300 $S gemm1 2048 1280 1280 4 0 1 1 > $O/plain_gemm2.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 1 -c 1 -f -o $O/r02_ncu_gemm_small $S gemm1 2048 1280 1280 4 0 1 1 > $O/ncu_gemm2.log 2>&1
ls -la $O/*.ncu-rep
echo done
