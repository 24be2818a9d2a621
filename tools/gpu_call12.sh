#!/bin/bash
# round-2 GPU call 12: per-warp clock64 timeline of the pipelined attention kernel (is the S -> max -> m_ready chain the block period?)
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  echo "== attn1 2 10 4096 4096"; timeout 120 $S attn1 2 10 4096 4096
  echo "== attn1 2 20 1024 1024"; timeout 120 $S attn1 2 20 1024 1024
} > gpurun_out/attn_trace_r2l.log 2>&1
echo done
