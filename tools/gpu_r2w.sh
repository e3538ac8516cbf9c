#!/bin/bash
# tensor-core mask kernels + row-stationary GroupNorm apply + bench step = one round of the forwards in flight:
# correctness of both mask families, bench A with the tensor family on, ncu of the mask kernels, then the engine-level suites
mkdir -p gpurun_out
export SMB_MASK_MMA=1
timeout 420 python -m pytest tests/test_gpu_postproc.py tests/test_gpu_conv.py -k "mask or multi_level" -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2w_pytest_mask.txt
tail -3 gpurun_out/r2w_pytest_mask.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2w_bench_A.json 2> gpurun_out/r2w_bench_A.err
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench_A.json')); print('A', round(d['value'],1), round(d['e2e']['value'],1), d['ms_per_step'], d['roofline']['frac'], d['clocks'])
for k in ('roofline_mask_assembly','roofline_mask_fused'): print(k, d[k]['ms'], d[k]['frac'], d[k]['other_family'])
print(d['serial'], d['e2e_dropin']['ms_per_image'])"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'mask_' -f -o gpurun_out/r2w_aux python tools/ncu_aux.py > gpurun_out/r2w_ncu.log 2>&1
ncu -i gpurun_out/r2w_aux.ncu-rep --page raw --csv > gpurun_out/r2w_aux_raw.csv 2>> gpurun_out/r2w_ncu.log
python tools/ncu_summary.py gpurun_out/r2w_aux_raw.csv > gpurun_out/r2w_aux_summary.txt 2>&1; cat gpurun_out/r2w_aux_summary.txt | cut -c1-160
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_head_api.py tests/test_gpu_vis.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2w_pytest_engine.txt
tail -3 gpurun_out/r2w_pytest_engine.txt
timeout 200 python tools/time_ops.py > gpurun_out/r2w_op_times.txt 2>&1; tail -16 gpurun_out/r2w_op_times.txt
