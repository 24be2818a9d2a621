#!/bin/bash
# round-2 GPU call 34: how many resident-kernel CTAs does an SM really hold?  (time steps at 148 * k tiles)
mkdir -p gpurun_out
O=gpurun_out/attn_occupancy_r2ah.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
for v in 3 5 6; do
  for H in 18 19 37 38 55 56 74 75; do
    echo "== variant $v H=$H tiles=$((H*8)) ==" >> $O
    ST_ATTN_IMPL=resident ST_ATTN_RES_VARIANT=$v timeout 100 $S attn1 1 $H 1024 1024 2>&1 | grep "TFLOP" | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
  done
done
echo done
