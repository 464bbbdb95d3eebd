#!/bin/bash
# 2-GPU session: bench with the C-ABI all-gather, then the configs[2] driver on 2 GPUs
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_b.json 2> gpurun_out/bench_2gpu_b.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_2gpu_b.json; tail -3 gpurun_out/bench_2gpu_b.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 examples/score_pairs.py --pairs 96 > gpurun_out/score_pairs_2gpu.json 2> gpurun_out/score_pairs_2gpu.err; echo "pairs rc=$?"; cat gpurun_out/score_pairs_2gpu.json; tail -3 gpurun_out/score_pairs_2gpu.err
timeout 600 python examples/score_pairs.py --pairs 48 > gpurun_out/score_pairs_1gpu.json 2> gpurun_out/score_pairs_1gpu.err; echo "pairs1 rc=$?"; cat gpurun_out/score_pairs_1gpu.json; tail -3 gpurun_out/score_pairs_1gpu.err
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -m gpu -k "pair_scorer" 2>&1 | tail -3
