#!/bin/bash
# round-2 GPU call 23: main-loop pace of the implicit-GEMM conv (4-D TMA A operand) against the plain GEMM of the same shape
mkdir -p gpurun_out
S=stabletriton_b200/csrc/selftest
{
  for bn in 0 -160 -256 160; do
    echo "== ctrace 2 32 32 1280 1280 bn=$bn"; timeout 60 $S ctrace 2 32 32 1280 1280 $bn | grep -v "^device"
    echo "== trace 2048 1280 11520 4 bn=$bn (same GEMM, 2-D A)"; timeout 60 $S trace 2048 1280 11520 4 $bn | grep -E "t\[[1-6]\]"
  done
  for bn in 0 -160 -256; do
    echo "== ctrace 2 128 128 320 320 bn=$bn"; timeout 60 $S ctrace 2 128 128 320 320 $bn | grep -v "^device"
    echo "== trace 32768 320 2880 4 bn=$bn"; timeout 60 $S trace 32768 320 2880 4 $bn | grep -E "t\[[1-6]\]"
  done
  for bn in 0 -160 -256; do
    echo "== ctrace 2 64 64 640 640 bn=$bn"; timeout 60 $S ctrace 2 64 64 640 640 $bn | grep -v "^device"
    echo "== trace 8192 640 5760 4 bn=$bn"; timeout 60 $S trace 8192 640 5760 4 $bn | grep -E "t\[[1-6]\]"
  done
} > gpurun_out/conv_trace_r2w.log 2>&1
echo done
