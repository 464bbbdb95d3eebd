#!/bin/bash
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"; cut -c1-330 gpurun_out/bench_4gpu.json; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_4gpu.err | tail -3
