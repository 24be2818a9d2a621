#!/bin/bash
# round-2 GPU call 25: do the waiting warps' barrier polls (SYNCS through the MIO queue) slow the MUFU stream?
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for cfg in "0 0" "200 0" "1000 0" "5000 0" "0 100" "0 300" "0 1000" "1000 300"; do
    set -- $cfg
    for parts in 4 2; do
      for shape in "2 10 4096 4096" "2 20 1024 1024"; do
        echo "== HINT=$1 SLEEP=$2 PARTS=$parts attn1 $shape"; ST_ATTN_WAIT_HINT=$1 ST_ATTN_WAIT_SLEEP=$2 ST_ATTN_PARTS=$parts timeout 60 $S attn1 $shape | grep -E "attention" | tail -1 | sed 's/.*worst@[^ ]* *//'
      done
    done
  done
} > gpurun_out/attn_wait_r2x.log 2>&1
echo done
