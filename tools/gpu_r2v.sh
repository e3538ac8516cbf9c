#!/bin/bash
# ncu --set full of the short-K conv (layer1.1.conv3, split epilogue) and a tower conv; raw + source pages as csv
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_gemm_kernel' -f -o gpurun_out/r2v_conv_full python tools/profile_conv.py 7 3 64 > gpurun_out/r2v_ncu.log 2>&1
ncu -i gpurun_out/r2v_conv_full.ncu-rep --page raw --csv > gpurun_out/r2v_conv_full_raw.csv 2>> gpurun_out/r2v_ncu.log
ncu -i gpurun_out/r2v_conv_full.ncu-rep --page source --csv --kernel-id ::regex:conv_gemm_kernel:1 > gpurun_out/r2v_conv_full_source_k1.csv 2>> gpurun_out/r2v_ncu.log
ls -la gpurun_out/r2v_*; tail -5 gpurun_out/r2v_ncu.log
