#!/bin/bash
# GroupNorm-statistics form of the split epilogue: correctness, timings, bench (same box: GN split on / off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2t_pytest.txt
tail -4 gpurun_out/r2t_pytest.txt
timeout 300 python tools/time_convs.py > gpurun_out/r2t_time_convs.txt 2>&1
grep "sum warm\|multi-level" gpurun_out/r2t_time_convs.txt
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2t_bench_A.json 2> gpurun_out/r2t_bench_A.err
timeout 400 python bench.py --workload C --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_C.json 2> gpurun_out/r2t_bench_C.err
timeout 400 python bench.py --workload B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_B.json 2> gpurun_out/r2t_bench_B.err
for w in A C B; do python -c "
import json; d=json.load(open('gpurun_out/r2t_bench_$w.json')); print('$w', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d['clocks'])"; done
