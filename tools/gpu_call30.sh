#!/bin/bash
# round-2 GPU call 30: resident (small-CTA, multi-block) attention kernel -- correctness + timing vs the pipelined kernel
mkdir -p gpurun_out
O=gpurun_out/attn_resident_r2ad.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
echo "== default ==" > $O
timeout 300 $S attn >> $O 2>&1
for v in 0 1 2; do
  echo "== resident variant $v ==" >> $O
  ST_ATTN_IMPL=resident ST_ATTN_RES_VARIANT=$v timeout 300 $S attn >> $O 2>&1
  echo "rc=$?" >> $O
done
echo done
