#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_diag.py > gpurun_out/r2i_conv_diag.txt 2>&1
cat gpurun_out/r2i_conv_diag.txt
