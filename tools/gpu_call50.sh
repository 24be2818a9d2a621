#!/bin/bash
# round-2 GPU call 50: residual chunk reads hoisted ahead of the epilogue math (TMA-residual path) -- parity + timing
mkdir -p gpurun_out
O=gpurun_out/gemm_res_r2ax.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
timeout 600 $S gemm 2>&1 | grep -v "PASS" | tail -3 >> $O
timeout 600 $S conv 2>&1 | grep -v "PASS" | tail -3 >> $O
timeout 100 $S trace 2048 1280 1280 4 0 1 2>&1 | grep "t\[" >> $O
for shape in "2048 1280 1280" "2048 1280 5120" "8192 640 640" "300 328 192"; do
  echo -n "gemm1 $shape +res: " >> $O
  timeout 100 $S gemm1 $shape 4 0 1 1 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
done
( timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear or conv or gemm or geglu" ) > gpurun_out/pytest_gpu_r2ax.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2ax.log
for pass in a b; do timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2ax_$pass.log 2>&1; done
echo done
