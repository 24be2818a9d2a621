#!/bin/bash
# round-2 GPU call 54: new residual-by-TMA test + the whole kernel test file
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q ) > gpurun_out/pytest_gpu_r2ba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2ba.log
echo done
