#!/bin/bash
# BASELINE configs[2]: 4096 preference pairs through PairScorer on N GPUs (strong scaling). usage: run_pairs4096.sh N [pairs]
N=${1:-1}; P=${2:-4096}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --workload pairs4096 --pairs $P > gpurun_out/pairs${P}_${N}gpu.json 2> gpurun_out/pairs${P}_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload pairs4096 --pairs $P > gpurun_out/pairs${P}_${N}gpu.json 2> gpurun_out/pairs${P}_${N}gpu.err
fi
echo "rc=$?"; cat gpurun_out/pairs${P}_${N}gpu.json | cut -c1-1800; tail -n 3 gpurun_out/pairs${P}_${N}gpu.err
