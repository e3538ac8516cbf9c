#!/bin/bash
# does a register-capped conv kernel (co-residency with the elementwise kernels) raise the in-flight throughput?
mkdir -p gpurun_out
for v in base lb512 lb608 base; do
  if [ $v = base ]; then unset SMB_LIB_PATH; else export SMB_LIB_PATH=$PWD/tools/_trace/libsipmask_b200_$v.so; fi
  timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_$v.json')); print('$v', round(d['value'],1), round(d['e2e']['value'],1), round(d['serial']['value'],1), round(d['roofline']['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
unset SMB_LIB_PATH
