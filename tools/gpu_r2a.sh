#!/bin/bash
# round-2 GPU run A: full -m gpu test suite, bench line, ncu --set full of the non-conv roofline kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc > gpurun_out/r2a_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2a_host.txt 2>&1; python -c "import os;print(len(os.sched_getaffinity(0)), os.cpu_count())" >> gpurun_out/r2a_host.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mask_assemble_kernel|mask_fused_pack_kernel|deform_im2col_multi_kernel|gn_apply_multi_kernel' --launch-skip 4 --launch-count 4 -f -o gpurun_out/r02_aux python tools/ncu_aux.py > gpurun_out/r2a_ncu.log 2>&1
ncu -i gpurun_out/r02_aux.ncu-rep --page raw --csv > gpurun_out/r02_aux_raw.csv 2>> gpurun_out/r2a_ncu.log
tail -5 gpurun_out/r2a_pytest.txt; head -c 600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
