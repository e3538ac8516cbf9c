#!/bin/bash
# which side of the conv pipeline is slow?  (profiling aid; results are wrong in debug modes)
for mode in 0 1 2; do
  echo "== SMB_CONV_PAIR=0 SMB_CONV_DEBUG=$mode"
  SMB_CONV_PAIR=0 SMB_CONV_DEBUG=$mode timeout 300 python tools/time_convs.py 2>&1 | grep -E "layer1.0.conv2|layer3.1.conv2|layer4.1.conv2|fpn_convs.4|multi-level x5  |sum warm|stem" | cut -c1-100 | head -12
done
