#!/bin/bash
# round-2 GPU call 53 (--gpus 8): weak headline + strong-scaling arm (8 prompts over 8 ranks, NCCL all-gather of the latents timed)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 30 --warmup 3 --sustained-seconds 0 ) > $O/bench_r2_8gpu_v4.json 2> $O/bench_r2_8gpu_v4.err; echo "bench rc=$?" >> $O/bench_r2_8gpu_v4.err
echo done
