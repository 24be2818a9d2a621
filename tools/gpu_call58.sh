#!/bin/bash
# round-2 GPU call 58: resident attention kernel sized for FOUR CTAs per SM (second S half re-read from TMEM, <= 96 registers)
mkdir -p gpurun_out
O=gpurun_out/attn_res4_r2bd.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
ST_ATTN_RES_CTAS=4 ST_ATTN_IMPL=resident timeout 300 $S attn 2>&1 | sed 's/worst@[^ ]* ref [-0-9.]*)//' >> $O
for shape in "2 20 1024 1024" "4 20 1024 1024" "8 20 1024 1024" "16 20 1024 1024" "2 10 4096 4096" "4 10 4096 4096" "16 10 4096 4096" "1 74 1024 1024" "1 75 1024 1024"; do
  for c in 3 4; do
    echo -n "$shape resident x$c: " >> $O
    ST_ATTN_RES_CTAS=$c ST_ATTN_IMPL=resident timeout 100 $S attn1 $shape 2>&1 | grep "TFLOP" | tail -1 | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
  done
  echo -n "$shape pipelined: " >> $O
  ST_ATTN_IMPL=pipelined timeout 100 $S attn1 $shape 2>&1 | grep "TFLOP" | tail -1 | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
done
echo done
