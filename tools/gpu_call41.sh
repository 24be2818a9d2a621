#!/bin/bash
# round-2 GPU call 41: pipelined vs resident attention kernel at the shapes the launcher's cycle model separates
mkdir -p gpurun_out
O=gpurun_out/attn_model_r2ao.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
while read B H TQ TK; do
  for impl in pipelined resident auto; do
    echo -n "B=$B H=$H Tq=$TQ Tk=$TK tiles=$((B*H*TQ/128)) $impl: " >> $O
    if [ $impl = auto ]; then unset ST_ATTN_IMPL; else export ST_ATTN_IMPL=$impl; fi
    timeout 100 $S attn1 $B $H $TQ $TK 2>&1 | grep "TFLOP" | tail -1 | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
  done
done <<LIST
2 20 1024 1024
4 20 1024 1024
8 20 1024 1024
16 20 1024 1024
2 10 4096 4096
4 10 4096 4096
16 10 4096 4096
1 10 16384 16384
3 20 1024 1024
2 20 1024 300
LIST
echo done
