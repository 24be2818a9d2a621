#!/bin/bash
# round-2 GPU call 10: ablation timing of the pipelined attention kernel (ST_ATTN_ABLATE bits; results are wrong by
# construction, only the time matters): which piece of the per-block pipeline sets the ~1.45 k-cycle period?
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for ab in 0 1 2 4 8 16 32 3 5 9 12 13 17 19 36 63; do
    for shape in "2 10 4096 4096" "2 20 1024 1024"; do
      echo "== ST_ATTN_ABLATE=$ab attn1 $shape"; ST_ATTN_ABLATE=$ab timeout 120 $S attn1 $shape | grep -E "attention" | tail -1 | sed 's/.*nan=[0-9]* //'
    done
  done
} > gpurun_out/attn_ablate_r2j.log 2>&1
echo done
{
  for mode in 0 1 2; do
    for shape in "2 128 128 320 320" "2 64 64 640 640" "2 32 32 1280 1280"; do
      echo "== conv1 $shape bn=0 mode=$mode (0 temb, 1 residual, 2 plain)"; timeout 120 $S conv1 $shape 0 $mode | grep -E "conv3x3" | tail -1
    done
  done
} > gpurun_out/conv_modes_r2j.log 2>&1
echo done2
