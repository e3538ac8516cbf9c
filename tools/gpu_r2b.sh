#!/bin/bash
# round-2 GPU run B: HBM-bound x tensor-bound conv overlap experiment, bench lines of workloads B / C / A101, failed tests again
mkdir -p gpurun_out
timeout 600 python tools/conv_mix.py > gpurun_out/r2b_conv_mix.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_vis.py tests/test_gpu_fullsize.py tests/test_gpu_postproc.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2b_pytest.txt
timeout 600 python bench.py --workload B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_B.json 2> gpurun_out/r2b_bench_B.err
timeout 600 python bench.py --workload C --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_C.json 2> gpurun_out/r2b_bench_C.err
timeout 600 python bench.py --workload A101 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_A101.json 2> gpurun_out/r2b_bench_A101.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
cat gpurun_out/r2b_conv_mix.txt; tail -5 gpurun_out/r2b_pytest.txt; head -c 300 gpurun_out/r2b_bench_B.json; tail -2 gpurun_out/r2b_bench_B.err; head -c 300 gpurun_out/r2b_bench_C.json; tail -2 gpurun_out/r2b_bench_C.err
