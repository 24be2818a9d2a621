#!/bin/bash
# round-2 GPU call 33: resident attention kernel, split-row softmax warps
mkdir -p gpurun_out
O=gpurun_out/attn_resident_r2ag.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
$S occ > $O 2>&1
for v in 5 6; do
  echo "== resident variant $v ==" >> $O
  ST_ATTN_IMPL=resident ST_ATTN_RES_VARIANT=$v timeout 300 $S attn 2>&1 | grep -v "PASS.*nan=0 worst@[0-9]*(got [-0-9.]* ref [-0-9.]*)$" >> $O
  echo "rc=$?" >> $O
done
echo done
