#!/bin/bash
# round-2 GPU call 55: N-split of the QKV projection GEMM (whole rounds of CTA pairs + remainder)
mkdir -p gpurun_out
timeout 600 python tools/split_probe.py > gpurun_out/split_probe_r2bb.log 2>&1; echo "rc=$?" >> gpurun_out/split_probe_r2bb.log
echo done
