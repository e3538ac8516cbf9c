#!/bin/bash
# round-2 GPU run B+C: overlap experiment, full -m gpu suite, bench lines of all workloads, ncu evidence
mkdir -p gpurun_out
timeout 600 python tools/conv_mix.py > gpurun_out/r2b_conv_mix.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2b_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_A.json 2> gpurun_out/r2b_bench_A.err
SMB_STEM_S2D=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2b_bench_A_stem448.json 2> gpurun_out/r2b_bench_A_stem448.err
timeout 600 python bench.py --workload B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_B.json 2> gpurun_out/r2b_bench_B.err
timeout 600 python bench.py --workload C --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_C.json 2> gpurun_out/r2b_bench_C.err
timeout 600 python bench.py --workload A101 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_A101.json 2> gpurun_out/r2b_bench_A101.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
timeout 600 python tools/time_convs.py > gpurun_out/r2b_time_convs.txt 2>&1
timeout 600 python tools/time_ops.py > gpurun_out/r2b_time_ops.txt 2>&1
# ncu: (1) launch list of a short bench run, (2) --set full of the aux kernels, (3) --set full of one whole engine step's convs
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/r2b_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mask_assemble_kernel|mask_fused_pack_kernel|deform_im2col_multi_kernel|gn_apply_multi_kernel' --launch-skip 4 --launch-count 4 -f -o gpurun_out/r02_aux python tools/ncu_aux.py > gpurun_out/r2b_ncu_aux.log 2>&1
ncu -i gpurun_out/r02_aux.ncu-rep --page raw --csv > gpurun_out/r02_aux_raw.csv 2>> gpurun_out/r2b_ncu_aux.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'conv_gemm_kernel' --launch-skip 146 --launch-count 73 -f -o gpurun_out/r02_conv_step python tools/profile_step.py > gpurun_out/r2b_ncu_conv.log 2>&1
ncu -i gpurun_out/r02_conv_step.ncu-rep --page raw --csv > gpurun_out/r02_conv_step_raw.csv 2>> gpurun_out/r2b_ncu_conv.log
rm -f gpurun_out/r02_conv_step.ncu-rep
cat gpurun_out/r2b_conv_mix.txt; tail -5 gpurun_out/r2b_pytest.txt; head -c 300 gpurun_out/r2b_bench_A.json; tail -2 gpurun_out/r2b_bench_A.err; head -c 200 gpurun_out/r2b_bench_B.json; tail -2 gpurun_out/r2b_bench_B.err; head -c 200 gpurun_out/r2b_bench_C.json; tail -2 gpurun_out/r2b_bench_C.err
