#!/bin/bash
# round-2 GPU call 14: per-warp timeline of the attention pipeline, full and with every stage switched off
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for ab in 0 255 63; do
    echo "== ST_ATTN_ABLATE=$ab attn1 2 10 4096 4096"; ST_ATTN_ABLATE=$ab timeout 120 $S attn1 2 10 4096 4096 | grep -v "^device"
  done
} > gpurun_out/attn_trace_r2n.log 2>&1
echo done
