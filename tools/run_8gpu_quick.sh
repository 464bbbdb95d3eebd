#!/bin/bash
# 8-GPU check of the weak-scaling bench line (short: 3 + 3 steps), under a tight timeout
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu_b.json 2> gpurun_out/bench_8gpu_b.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_8gpu_b.json; tail -3 gpurun_out/bench_8gpu_b.err | cut -c1-300
