#!/bin/bash
# round-2 GPU call 26: spin-polling the GEMM ring barriers (ST_GEMM_SPIN bit 0: MMA warp, bit 1: TMA producer)
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for sp in 0 1 2 3; do
    echo "== ST_GEMM_SPIN=$sp ctrace 2 32 32 1280 1280 -160"; ST_GEMM_SPIN=$sp timeout 60 $S ctrace 2 32 32 1280 1280 -160 | grep -E "t\[[23]\]"
    echo "== ST_GEMM_SPIN=$sp ctrace 2 32 32 1280 1280 160"; ST_GEMM_SPIN=$sp timeout 60 $S ctrace 2 32 32 1280 1280 160 | grep -E "t\[[23]\]"
    echo "== ST_GEMM_SPIN=$sp trace 2048 10240 1280 6 -256"; ST_GEMM_SPIN=$sp timeout 60 $S trace 2048 10240 1280 6 -256 | grep -E "t\[[23]\]"
    for shape in "2048 1280 1280 4 0 1 1" "2048 1280 5120 4 0 1 1" "2048 3840 1280 4 0 0 0" "2048 10240 1280 6 0 1 0" "8192 640 640 4 0 1 1"; do
      echo "== ST_GEMM_SPIN=$sp gemm1 $shape"; ST_GEMM_SPIN=$sp timeout 60 $S gemm1 $shape | grep -E "gemm M" | tail -1 | sed 's/.*worst@[^ ]* *//'
    done
    for shape in "2 32 32 1280 1280 0" "2 64 64 640 640 0" "2 128 128 320 320 0"; do
      echo "== ST_GEMM_SPIN=$sp conv1 $shape"; ST_GEMM_SPIN=$sp timeout 60 $S conv1 $shape | grep -E "conv3x3" | tail -1 | sed 's/.*worst@[^ ]* *//'
    done
  done
} > $O/gemm_spin_r2y.log 2>&1
for pass in a b; do
  for sp in 0 3 1; do ST_GEMM_SPIN=$sp timeout 600 python tools/quick_bench.py > $O/qb_r2y_spin${sp}_$pass.log 2>&1; done
done
echo done
