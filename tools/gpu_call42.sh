#!/bin/bash
# round-2 GPU call 42: what does the residual operand cost the single-wave GEMM?  (phase timeline with / without)
mkdir -p gpurun_out
O=gpurun_out/gemm_res_r2ap.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
for res in 1 0; do
  echo "== trace 2048 1280 1280 res=$res ==" >> $O
  timeout 100 $S trace 2048 1280 1280 4 0 $res >> $O 2>&1
  echo "== gemm1 2048 1280 1280 bias=1 res=$res ==" >> $O
  timeout 100 $S gemm1 2048 1280 1280 4 0 1 $res 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
  echo "== gemm1 2048 1280 5120 bias=1 res=$res ==" >> $O
  timeout 100 $S gemm1 2048 1280 5120 4 0 1 $res 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
  echo "== gemm1 8192 640 640 bias=1 res=$res ==" >> $O
  timeout 100 $S gemm1 8192 640 640 4 0 1 $res 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
done
echo done
