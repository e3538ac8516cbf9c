#!/bin/bash
# short-K 1x1 kernel (one tile per CTA): correctness, per-conv timings with / without it, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2k_pytest.txt
tail -5 gpurun_out/r2k_pytest.txt
timeout 300 python tools/time_convs.py > gpurun_out/r2k_time_convs_small1.txt 2>&1
SMB_CONV_SMALL=0 timeout 300 python tools/time_convs.py > gpurun_out/r2k_time_convs_small0.txt 2>&1
grep "sum warm" gpurun_out/r2k_time_convs_small1.txt gpurun_out/r2k_time_convs_small0.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2k_bench_small1.json 2> gpurun_out/r2k_bench_small1.err
SMB_CONV_SMALL=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2k_bench_small0.json 2> gpurun_out/r2k_bench_small0.err
head -c 200 gpurun_out/r2k_bench_small1.json; echo; head -c 200 gpurun_out/r2k_bench_small0.json; echo; tail -2 gpurun_out/r2k_bench_small1.err
