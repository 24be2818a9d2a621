#!/bin/bash
# round-2 GPU call 28b: MUFU-pipe time division with parked (nanosleep) waits
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for cfg in "0 0" "1024 20" "1024 50" "1100 30" "1200 30" "1300 30" "1100 100"; do
    set -- $cfg
    for shape in "2 10 4096 4096" "2 10 16384 16384"; do
      echo "== ST_ATTN_PACE=$1 SLEEP=$2 attn1 $shape"; ST_ATTN_PACE=$1 ST_ATTN_PACE_SLEEP=$2 timeout 60 $S attn1 $shape | grep -E "attention" | tail -1 | sed 's/.*worst@[^ ]* *//'
    done
  done
} > gpurun_out/attn_pace_r2ab.log 2>&1
echo done
