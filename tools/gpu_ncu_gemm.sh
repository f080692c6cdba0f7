#!/bin/bash
mkdir -p gpurun_out
M1="python tools/conv_micro.py 192 384 5 1 0 433 128 0 tf32 3"
$M1 > gpurun_out/r2_gemm1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o gpurun_out/r2_gemm_wn_in $M1 > gpurun_out/r2_gemm1_ncu.log 2>&1
M2="python tools/conv_micro.py 1536 192 1 1 0 866 128 0 tf32 3"
$M2 > gpurun_out/r2_gemm2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o gpurun_out/r2_gemm_ffn_w2 $M2 > gpurun_out/r2_gemm2_ncu.log 2>&1
ls -la gpurun_out/r2_gemm*
