#!/bin/bash
# round-2 GPU call 37: resident kernel, shorter issuer / max phases, unordered exponentials
mkdir -p gpurun_out
O=gpurun_out/attn_chain_r2ak.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
for v in 0 1; do
    echo "== variant $v trace ==" >> $O
    ST_ATTN_IMPL=resident ST_ATTN_RES_VARIANT=$v timeout 100 $S attn1 2 18 1024 1024 2>&1 | grep "TFLOP\|resident slot" | sed 's/.*worst@[^ ]* ref [-0-9.]*)//' >> $O
    echo "== variant $v suite ==" >> $O
    ST_ATTN_IMPL=resident ST_ATTN_RES_VARIANT=$v timeout 300 $S attn 2>&1 | grep -v "PASS.*nan=0 worst@[0-9]*(got [-0-9.]* ref [-0-9.]*)$" | sed 's/worst@[^ ]* ref [-0-9.]*)//' >> $O
done
echo done
