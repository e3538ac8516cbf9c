#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --workload C --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_C.json 2> gpurun_out/r2u_bench_C.err
timeout 400 python bench.py --workload B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_B.json 2> gpurun_out/r2u_bench_B.err
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py tests/test_gpu_vis.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for w in C B; do python -c "
import json; d=json.load(open('gpurun_out/r2u_bench_$w.json')); print('$w', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d['clocks'])"; done
