#!/bin/bash
# split-group epilogue: correctness, per-conv timings on / off, bench on / off, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2n_pytest.txt
tail -5 gpurun_out/r2n_pytest.txt
timeout 300 python tools/time_convs.py > gpurun_out/r2n_time_convs_split1.txt 2>&1
SMB_CONV_EPI_SPLIT=0 timeout 300 python tools/time_convs.py > gpurun_out/r2n_time_convs_split0.txt 2>&1
grep "sum warm" gpurun_out/r2n_time_convs_split1.txt gpurun_out/r2n_time_convs_split0.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2n_bench_split1.json 2> gpurun_out/r2n_bench_split1.err
SMB_CONV_EPI_SPLIT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2n_bench_split0.json 2> gpurun_out/r2n_bench_split0.err
head -c 200 gpurun_out/r2n_bench_split1.json; echo; head -c 200 gpurun_out/r2n_bench_split0.json; echo; tail -2 gpurun_out/r2n_bench_split1.err
export SMB_LIB_PATH=$PWD/tools/_trace/libsipmask_b200_trace.so
timeout 200 python tools/conv_trace.py layer1.1.conv3 48 > gpurun_out/r2n_conv_trace.txt 2>&1
timeout 200 python tools/conv_trace.py layer1.0.downsample 48 >> gpurun_out/r2n_conv_trace.txt 2>&1
grep -v "^pairs" gpurun_out/r2n_conv_trace.txt | head -40
paste <(awk '{print $1, $(NF-3)}' gpurun_out/r2n_time_convs_split1.txt) <(awk '{print $(NF-3)}' gpurun_out/r2n_time_convs_split0.txt) | head -70
