#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/pipebench > gpurun_out/pipebench_r2p.log 2>&1
echo done
