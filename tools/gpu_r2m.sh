#!/bin/bash
mkdir -p gpurun_out
SMB_DIAG_QUICK=1 timeout 300 python tools/conv_diag.py > gpurun_out/r2m_conv_diag.txt 2>&1
cat gpurun_out/r2m_conv_diag.txt
export SMB_LIB_PATH=$PWD/tools/_trace/libsipmask_b200_trace.so
out=gpurun_out/r2m_conv_trace.txt
: > $out
for dbg in 0 248 760 256; do
  SMB_CONV_DEBUG=$dbg timeout 200 python tools/conv_trace.py layer1.0.downsample 48 >> $out 2>&1
done
SMB_CONV_DEBUG=0 timeout 200 python tools/conv_trace.py layer1.1.conv3 48 >> $out 2>&1
SMB_CONV_DEBUG=760 timeout 200 python tools/conv_trace.py layer1.1.conv3 48 >> $out 2>&1
grep -v "^pairs above" $out
