#!/bin/bash
# round-2 GPU call 48: CTA pairs on / off for the 3x3 convolutions inside the real launch sequence (same box, two passes)
mkdir -p gpurun_out
O=gpurun_out
for pass in a b; do
  timeout 600 python tools/quick_bench.py > $O/qb_r2av_pairs_$pass.log 2>&1
  ST_CONV_CLUSTER=0 timeout 600 python tools/quick_bench.py > $O/qb_r2av_nopairs_$pass.log 2>&1
done
echo done
