#!/bin/bash
# round-2 GPU call 47: phase timelines of the 3x3 convolutions at the SDXL levels (pair kernel vs one-CTA kernel)
mkdir -p gpurun_out
O=gpurun_out/conv_trace_r2au.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
for cl in auto 0 1; do
  if [ $cl = auto ]; then unset ST_GEMM_CLUSTER; else export ST_GEMM_CLUSTER=$cl; fi
  for shape in "2 128 128 320 320" "2 64 64 640 640" "2 32 32 1280 1280" "2 128 128 640 320"; do
    echo "== ST_GEMM_CLUSTER=$cl ctrace $shape ==" >> $O
    timeout 100 $S ctrace $shape 0 2>&1 | grep -v "^device" >> $O
    timeout 100 $S conv1 $shape 0 0 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
  done
done
echo done
