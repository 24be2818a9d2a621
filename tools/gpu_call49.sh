#!/bin/bash
# round-2 GPU call 49: conv pair rule in the launcher -- conv parity + whole-forward check (same box: default vs all pairs off)
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
timeout 600 $S conv 2>&1 | tail -2 > $O/conv_rule_r2aw.log
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_sdxl_parity_gpu.py -m gpu -x -q -k "conv or sdxl or unet" ) >> $O/conv_rule_r2aw.log 2>&1; echo "pytest rc=$?" >> $O/conv_rule_r2aw.log
for pass in a b; do
  timeout 600 python tools/quick_bench.py > $O/qb_r2aw_rule_$pass.log 2>&1
  ST_CONV_CLUSTER=0 timeout 600 python tools/quick_bench.py > $O/qb_r2aw_nopairs_$pass.log 2>&1
done
echo done
