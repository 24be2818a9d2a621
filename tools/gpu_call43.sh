#!/bin/bash
# round-2 GPU call 43: residual tile by TMA into the idle operand ring (last tile of a CTA) -- parity + timing A/B
mkdir -p gpurun_out
O=gpurun_out/gemm_res_r2aq.log
S=stabletriton_b200/csrc/selftest
export LD_LIBRARY_PATH=stabletriton_b200/csrc:$LD_LIBRARY_PATH
: > $O
timeout 600 $S gemm 2>&1 | grep -v "PASS" | tail -5 >> $O
timeout 600 $S conv 2>&1 | grep -v "PASS" | tail -5 >> $O
for mode in 1 0; do
  export ST_GEMM_RES_TMA=$mode
  echo "== ST_GEMM_RES_TMA=$mode trace 2048 1280 1280 res=1 ==" >> $O
  timeout 100 $S trace 2048 1280 1280 4 0 1 2>&1 | grep "t\[" >> $O
  for shape in "2048 1280 1280" "2048 1280 5120" "8192 640 640" "8192 640 2560" "2048 1280 1920" "300 328 192"; do
    echo -n "RES_TMA=$mode gemm1 $shape +res: " >> $O
    timeout 100 $S gemm1 $shape 4 0 1 1 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
  done
  echo -n "RES_TMA=$mode conv1 2 64 64 640 640 res: " >> $O
  timeout 100 $S conv1 2 64 64 640 640 0 1 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
  echo -n "RES_TMA=$mode conv1 2 32 32 1280 1280 res: " >> $O
  timeout 100 $S conv1 2 32 32 1280 1280 0 1 2>&1 | grep TFLOP | sed 's/worst@.*)//' >> $O
done
unset ST_GEMM_RES_TMA
( timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "linear or conv or gemm or geglu" ) > gpurun_out/pytest_gpu_r2aq.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2aq.log
for pass in a b; do
  ST_GEMM_RES_TMA=0 timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2aq_regs_$pass.log 2>&1
  timeout 600 python tools/quick_bench.py > gpurun_out/qb_r2aq_tma_$pass.log 2>&1
done
echo done
