#!/bin/bash
mkdir -p gpurun_out
MMR_GEMM_CLUSTER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_pair16_kernel|gemm_ln_kernel' -c 8 -f -o gpurun_out/r01f_gemm python tools/ncu_gemm_shapes.py 2 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gemm.log; ls -la gpurun_out/*.ncu-rep
