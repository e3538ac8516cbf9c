#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_postproc.py tests/test_gpu_ref_cuda.py tests/test_head_api.py tests/test_gpu_engine.py tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2g_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench_A.json 2> gpurun_out/r2g_bench_A.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mask_assemble_kernel|mask_fused_pack_kernel|deform_im2col_multi_kernel|gn_apply_multi_kernel' --launch-skip 4 --launch-count 4 -f -o gpurun_out/r02_aux python tools/ncu_aux.py > gpurun_out/r2g_ncu_aux.log 2>&1
ncu -i gpurun_out/r02_aux.ncu-rep --page raw --csv > gpurun_out/r02_aux_raw.csv 2>> gpurun_out/r2g_ncu_aux.log
tail -4 gpurun_out/r2g_pytest.txt; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_A.json')); print(d['value'], d['e2e']['value'], d['roofline_mask_assembly'], d['roofline_mask_fused']['ms'])"; tail -2 gpurun_out/r2g_bench_A.err
