#!/bin/bash
# round-2 GPU call 56: ncu --set full of the single-wave +residual GEMM, residual by TMA vs by row loads
mkdir -p gpurun_out
O=gpurun_out
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
timeout 300 $S gemm1 2048 1280 1280 4 0 1 1 > $O/plain_gemm_res.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 1 -c 1 -f -o $O/r02_ncu_gemm_res_tma $S gemm1 2048 1280 1280 4 0 1 1 > $O/ncu_gemm_res_tma.log 2>&1
ST_GEMM_RES_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 1 -c 1 -f -o $O/r02_ncu_gemm_res_rows $S gemm1 2048 1280 1280 4 0 1 1 > $O/ncu_gemm_res_rows.log 2>&1
ls -la $O/r02_ncu_gemm_res_*.ncu-rep
echo done
