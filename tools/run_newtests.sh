#!/bin/bash
# GPU call: the tests added last (full-size multi-tile row, configs[4] at full depth), with timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_forward_gpu.py -x -q -m gpu -k "multi_tile_row or full_depth" --durations=5 > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; tail -n 14 gpurun_out/pytest_new.log
free -g | head -2
