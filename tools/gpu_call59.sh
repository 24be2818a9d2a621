#!/bin/bash
# round-2 GPU call 59: four-CTA resident form chosen by shape -- attention tests, B = 16 / B = 2 forward A/B on one box
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" ) > $O/pytest_gpu_r2be.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2be.log
for pass in a b; do
  ST_ATTN_RES_CTAS=3 ST_ATTN_IMPL=noresident timeout 600 python tools/quick_bench.py --batch 16 > $O/qb_r2be_b16_nores_$pass.log 2>&1
  timeout 600 python tools/quick_bench.py --batch 16 > $O/qb_r2be_b16_new_$pass.log 2>&1
done
timeout 600 python tools/quick_bench.py > $O/qb_r2be_b2_new.log 2>&1
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
for c in 3 4; do echo -n "T16384 resident x$c: " >> $O/pytest_gpu_r2be.log; ST_ATTN_RES_CTAS=$c ST_ATTN_IMPL=resident timeout 100 $S attn1 1 10 16384 16384 2>&1 | grep TFLOP | tail -1 | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O/pytest_gpu_r2be.log; done
echo -n "T16384 pipelined: " >> $O/pytest_gpu_r2be.log; ST_ATTN_IMPL=pipelined timeout 100 $S attn1 1 10 16384 16384 2>&1 | grep TFLOP | tail -1 | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O/pytest_gpu_r2be.log
echo done
