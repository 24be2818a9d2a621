#!/bin/bash
# round-2 GPU call 35: resident kernel chain trace + spin waits
mkdir -p gpurun_out
O=gpurun_out/attn_chain_r2ai.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
for spin in 0 1 2 3; do
  for H in 18 40; do
    echo "== variant 3 spin=$spin H=$H ==" >> $O
    ST_ATTN_RES_SPIN=$spin ST_ATTN_IMPL=resident ST_ATTN_RES_VARIANT=3 timeout 100 $S attn1 2 $H 1024 1024 2>&1 | grep "TFLOP\|resident slot" | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
  done
done
ST_ATTN_RES_SPIN=3 ST_ATTN_IMPL=resident ST_ATTN_RES_VARIANT=3 timeout 100 $S attn1 2 10 4096 4096 2>&1 | grep "TFLOP" | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
echo done
