#!/bin/bash
# multi-GPU scaling of workload A on one 8-GPU box: N = 1, 2, 8 back to back
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2e_gpus.txt
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2e_scale_1.json 2> gpurun_out/r2e_scale_1.err
for N in 2 8; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2e_scale_$N.json 2> gpurun_out/r2e_scale_$N.err
done
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload C --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_scale_C8.json 2> gpurun_out/r2e_scale_C8.err
for f in 1 2 8 C8; do head -c 160 gpurun_out/r2e_scale_$f.json; echo; tail -1 gpurun_out/r2e_scale_$f.err; done
