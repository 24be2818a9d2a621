#!/bin/bash
# round-2 GPU call 3: 160-wide tiles + CTA pairs per shape; new finalize kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -s ) > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c.log
S=stabletriton_b200/csrc/selftest
g() { echo "-- gemm $*"; timeout 120 $S gemm1 "$@" | grep -E "gemm M|FAIL"; }
c() { echo "-- conv $*"; timeout 120 $S conv1 "$@" | grep -E "conv|FAIL"; }
{
  export ST_GEMM_CLUSTER=0
  echo "===== plain (ST_GEMM_CLUSTER=0, launcher's own width) vs forced"
  for bn in 0 160 192 -160 -192 -256; do g 2048 1280 1280 4 $bn 1 1; done
  for bn in 0 160 192 -160 -192 -256; do g 2048 1280 5120 4 $bn 1 1; done
  for bn in 0 256 -256 -192; do g 2048 3840 1280 4 $bn 0 0; done
  for bn in 0 -256; do g 2048 10240 1280 6 $bn 1 0; done
  for bn in 0 -256; do g 8192 5120 640 6 $bn 1 0; done
  for bn in 0 160 -160 -192 -256; do g 8192 640 640 4 $bn 1 1; done
  for bn in 0 -160 -192 -256; do g 8192 640 2560 4 $bn 1 1; done
  for bn in 0 -256 -192; do g 8192 1920 640 4 $bn 0 0; done
  for bn in 0 -256; do g 154 166400 2048 4 $bn 0 0; done
  for bn in 0 160 -160 -192 -256; do c 2 128 128 320 320 $bn; done
  for bn in 0 -160 -192 -256; do c 2 128 128 640 640 $bn; done
  for bn in 0 160 -160 -192 -256; do c 2 64 64 640 640 $bn; done
  for bn in 0 160 -160 -192 -256; do c 2 32 32 1280 1280 $bn; done
  for bn in 0 -160 -256; do c 2 32 32 2560 1280 $bn; done
  echo "===== traces"
  $S trace 8192 640 640 4 0
  $S trace 2048 1280 1280 4 -160
  $S trace 2048 1280 5120 4 -160
} > gpurun_out/tiles_r2c.log 2>&1
unset ST_GEMM_CLUSTER
ST_GEMM_CLUSTER=0 timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2c_plain.log 2>&1
timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2c_auto.log 2>&1
echo done
