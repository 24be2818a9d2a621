#!/bin/bash
# round-2 GPU call 27 (--gpus 2): the multi-GPU tests as the driver will run them (whole -m gpu suite on a 2-GPU box)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s ) > $O/pytest_gpu_r2z_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2z_2gpu.log
echo done
