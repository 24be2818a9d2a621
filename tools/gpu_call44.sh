#!/bin/bash
# round-2 GPU call 44 (--gpus 2): multi-GPU tests + 2-GPU bench line with the final library (resident attention kernel, TMA residual)
mkdir -p gpurun_out
O=gpurun_out
export NCCL_DEBUG=WARN
( time timeout 420 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -s ) > $O/pytest_gpu_r2ar_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2ar_2gpu.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 30 --warmup 3 --sustained-seconds 0 ) > $O/bench_r2_2gpu_v4.json 2> $O/bench_r2_2gpu_v4.err; echo "bench rc=$?" >> $O/bench_r2_2gpu_v4.err
echo done
