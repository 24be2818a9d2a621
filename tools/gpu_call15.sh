#!/bin/bash
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{ echo "== attn1 2 10 4096 4096"; timeout 120 $S attn1 2 10 4096 4096 | grep -v "^device"; } > gpurun_out/attn_trace_r2o.log 2>&1
echo done
