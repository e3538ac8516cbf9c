#!/bin/bash
# conv epilogue with packed f32x2 adds + fused relu pack + division-free tile decode: correctness, timings, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_head_api.py tests/test_gpu_vis.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2h_pytest.txt
timeout 600 python tools/time_convs.py > gpurun_out/r2h_time_convs.txt 2>&1
timeout 600 python tools/conv_mix.py > gpurun_out/r2h_conv_mix.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2h_bench_A.json 2> gpurun_out/r2h_bench_A.err
tail -3 gpurun_out/r2h_pytest.txt; grep "sum warm" gpurun_out/r2h_time_convs.txt; head -3 gpurun_out/r2h_conv_mix.txt; head -c 300 gpurun_out/r2h_bench_A.json; tail -2 gpurun_out/r2h_bench_A.err
