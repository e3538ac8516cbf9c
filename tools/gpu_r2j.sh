#!/bin/bash
# tuning sweep of workload A: forwards in flight x grid cap x planner min_tiles, and batch-B forwards
mkdir -p gpurun_out
out=gpurun_out/r2j_sweep.txt; : > $out
for cfg in "6 48 16" "6 56 16" "6 40 16" "8 40 16" "8 48 16" "5 56 16" "4 64 16" "6 48 32" "6 64 16" "10 36 16"; do
  set -- $cfg; timeout 200 python tools/time_inflight.py $1 $2 $3 >> $out 2>&1
done
SMB_BATCH=2 timeout 200 python tools/time_inflight.py 3 64 16 >> $out 2>&1
SMB_BATCH=3 timeout 200 python tools/time_inflight.py 2 74 16 >> $out 2>&1
SMB_BATCH=6 timeout 200 python tools/time_inflight.py 1 0 48 >> $out 2>&1
SMB_BATCH=6 timeout 200 python tools/time_inflight.py 2 74 16 >> $out 2>&1
SMB_BATCH=8 timeout 200 python tools/time_inflight.py 1 0 48 >> $out 2>&1
cat $out | grep in_flight
