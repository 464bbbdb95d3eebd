#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 1200 python -m pytest tests/test_forward_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_fold.json > gpurun_out/bench_fold.json 2> gpurun_out/bench_fold.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_fold.json; tail -3 gpurun_out/bench_fold.err
