#!/bin/bash
# round-2 GPU call 13: attention ablations, second batch (TMA traffic, tcgen05.ld of S)
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for ab in 0 64 128 192 63 127 191 255 100 36 44; do
    for shape in "2 10 4096 4096" "2 20 1024 1024"; do
      echo "== ST_ATTN_ABLATE=$ab attn1 $shape"; ST_ATTN_ABLATE=$ab timeout 120 $S attn1 $shape | grep -E "attention" | tail -1 | sed 's/.*nan=[0-9]* //'
    done
  done
} > gpurun_out/attn_ablate_r2m.log 2>&1
echo done
