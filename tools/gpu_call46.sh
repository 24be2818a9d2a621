#!/bin/bash
# round-2 GPU call 46: resident attention kernel with a software-pipelined exp phase
mkdir -p gpurun_out
O=gpurun_out/attn_swp_r2at.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
ST_ATTN_IMPL=resident timeout 100 $S attn1 2 18 1024 1024 2>&1 | grep "TFLOP\|resident slot" | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
ST_ATTN_IMPL=resident timeout 300 $S attn 2>&1 | sed 's/worst@[^ ]* ref [-0-9.]*)//' >> $O
for shape in "2 20 1024 1024" "4 20 1024 1024" "16 20 1024 1024" "2 10 4096 4096" "16 10 4096 4096"; do
  for impl in pipelined resident; do
    echo -n "$shape $impl: " >> $O
    ST_ATTN_IMPL=$impl timeout 100 $S attn1 $shape 2>&1 | grep "TFLOP" | tail -1 | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
  done
done
echo done
