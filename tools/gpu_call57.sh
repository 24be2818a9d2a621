#!/bin/bash
# round-2 GPU call 57: phase timeline of the M = 8192 level GEMMs (two tiles per CTA, short K)
mkdir -p gpurun_out
O=gpurun_out/gemm_m8192_r2bc.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
for res in 1 0; do
  for bn in 0 160 128 192 256 64; do
    echo "== trace 8192 640 640 bn=$bn res=$res ==" >> $O
    timeout 100 $S trace 8192 640 640 4 $bn $res 2>&1 | grep "t\[[1-6]\]" >> $O
    timeout 100 $S gemm1 8192 640 640 4 $bn 1 $res 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
  done
done
echo "== 8192 640 2560 ==" >> $O
timeout 100 $S trace 8192 640 2560 4 0 1 2>&1 | grep "t\[[1-6]\]" >> $O
timeout 100 $S gemm1 8192 640 2560 4 0 1 1 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
echo "== 8192 1920 640 ==" >> $O
timeout 100 $S trace 8192 1920 640 4 0 0 2>&1 | grep "t\[[1-6]\]" >> $O
timeout 100 $S gemm1 8192 1920 640 4 0 1 0 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
echo done
