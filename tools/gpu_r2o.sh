#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/conv_diag3x3.py > gpurun_out/r2o_conv_diag3x3.txt 2>&1
cat gpurun_out/r2o_conv_diag3x3.txt
