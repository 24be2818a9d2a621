#!/bin/bash
# round-2 GPU call 8 (--gpus 2): the multi-GPU data path on hardware -- NCCL tests + 2-GPU bench (weak headline,
# strong-scaling arm with the latent all-gather, CFG-split arm with the per-step all-gather inside the step graph)
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/r2h_gpus.txt 2>&1
( time timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_cabi.py -m gpu -x -q -s ) > $O/pytest_gpu_r2h_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2h_2gpu.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 30 --warmup 3 ) > $O/bench_r2_2gpu_v2.json 2> $O/bench_r2_2gpu_v2.err; echo "bench rc=$?" >> $O/bench_r2_2gpu_v2.err
echo done
