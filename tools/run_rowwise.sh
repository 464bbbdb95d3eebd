#!/bin/bash
timeout 900 python -m pytest tests/test_frames_gpu.py tests/test_forward_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown gpurun_out/breakdown_rw.json > gpurun_out/bench_rw.json 2> gpurun_out/bench_rw.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_rw.json
