#!/bin/bash
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"; cut -c1-330 gpurun_out/bench_8gpu.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_8gpu.err | tail -5
