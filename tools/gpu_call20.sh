#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_probe.py > gpurun_out/conv_probe_r2t.log 2>&1
echo done
