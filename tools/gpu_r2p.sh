#!/bin/bash
# split epilogue with 2 / 3 / 4 warp groups: correctness, per-conv timings, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2p_pytest.txt
tail -5 gpurun_out/r2p_pytest.txt
for g in 4 3 2; do
  SMB_CONV_EPI_GROUPS=$g timeout 300 python tools/time_convs.py > gpurun_out/r2p_time_convs_g$g.txt 2>&1
  grep "sum warm" gpurun_out/r2p_time_convs_g$g.txt
done
for g in 4 3 2; do
  SMB_CONV_EPI_GROUPS=$g timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2p_bench_g$g.json 2> gpurun_out/r2p_bench_g$g.err
  head -c 120 gpurun_out/r2p_bench_g$g.json; echo
done
paste <(awk '{printf "%-45s %6s %5s %5s %8s\n", $1, $2, $3, $4, $6}' gpurun_out/r2p_time_convs_g4.txt) <(awk '{print $6}' gpurun_out/r2p_time_convs_g3.txt) <(awk '{print $6}' gpurun_out/r2p_time_convs_g2.txt) | sed -n 2,64p
