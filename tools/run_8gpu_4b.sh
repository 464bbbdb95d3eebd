#!/bin/bash
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --model 4b --steps 3 --warmup 3 > gpurun_out/bench_8gpu_4b.json 2> gpurun_out/bench_8gpu_4b.err; echo "bench8-4b rc=$?"; cut -c1-330 gpurun_out/bench_8gpu_4b.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_8gpu_4b.err | tail -5
