#!/bin/bash
# per-role trace of the short-K convolutions (CTA 0), plain and with the epilogue work disabled, full grid and capped
mkdir -p gpurun_out
export SMB_LIB_PATH=$PWD/tools/_trace/libsipmask_b200_trace.so
out=gpurun_out/r2l_conv_trace.txt
: > $out
for pat in layer1.0.downsample layer1.1.conv3 layer2.1.conv3 layer1.1.conv1; do
  for cap in 0 48; do
    timeout 200 python tools/conv_trace.py $pat $cap >> $out 2>&1
    SMB_CONV_DEBUG=248 timeout 200 python tools/conv_trace.py $pat $cap >> $out 2>&1
  done
done
SMB_CONV_PAIR=0 timeout 200 python tools/conv_trace.py layer1.0.downsample 0 >> $out 2>&1
grep -v "^pairs above" $out | head -150
