#!/bin/bash
# round-2 verification run: full -m gpu suite, smoke, bench lines of all workloads + reference arm, per-kernel timings, ncu evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2q_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2q_smoke.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2q_bench_A.json 2> gpurun_out/r2q_bench_A.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err
timeout 600 python bench.py --workload B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_B.json 2> gpurun_out/r2q_bench_B.err
timeout 600 python bench.py --workload C --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_C.json 2> gpurun_out/r2q_bench_C.err
timeout 600 python bench.py --workload A101 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_A101.json 2> gpurun_out/r2q_bench_A101.err
timeout 600 python tools/time_convs.py > gpurun_out/r2q_time_convs.txt 2>&1
timeout 600 python tools/time_ops.py > gpurun_out/r2q_time_ops.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/r2q_ncu_launches.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'conv_gemm_kernel' --launch-skip 146 --launch-count 73 -f -o gpurun_out/r2q_conv_step python tools/profile_step.py > gpurun_out/r2q_ncu_conv.log 2>&1
ncu -i gpurun_out/r2q_conv_step.ncu-rep --page raw --csv > gpurun_out/r2q_conv_step_raw.csv 2>> gpurun_out/r2q_ncu_conv.log
rm -f gpurun_out/r2q_conv_step.ncu-rep
tail -5 gpurun_out/r2q_pytest.txt; cat gpurun_out/r2q_smoke.txt | tail -2; head -c 300 gpurun_out/r2q_bench_A.json; echo; tail -2 gpurun_out/r2q_bench_A.err; head -c 200 gpurun_out/r2q_bench_B.json; echo; head -c 200 gpurun_out/r2q_bench_C.json; echo; head -c 200 gpurun_out/r2q_bench_A101.json; echo; head -c 300 gpurun_out/r2q_bench_ref.json; echo
