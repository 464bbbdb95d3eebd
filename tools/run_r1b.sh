#!/bin/bash
# GPU session: new Llama / config-5 tests, then benches for configs[1], [4] and [3]
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py -x -q -m gpu -k "llama or config5" > gpurun_out/t_llama.log 2>&1; echo "llama tests rc=$?" 
tail -5 gpurun_out/t_llama.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2b.json 2> gpurun_out/bench_2b.err; echo "bench2b rc=$?"; cat gpurun_out/bench_2b.json | cut -c1-400
timeout 900 python bench.py --model 4b --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_4b.json > gpurun_out/bench_4b.json 2> gpurun_out/bench_4b.err; echo "bench4b rc=$?"; cat gpurun_out/bench_4b.json | cut -c1-400; tail -3 gpurun_out/bench_4b.err
timeout 900 python bench.py --workload long16x7 --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_long.json > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; echo "benchlong rc=$?"; cat gpurun_out/bench_long.json | cut -c1-400; tail -3 gpurun_out/bench_long.err
