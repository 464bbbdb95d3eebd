#!/bin/bash
# GPU call: frame-path tests (golden + fused vs two-pass), frame-path timing, plus the Phi-3 / tail-attention tests added this round
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frames_gpu.py -x -q -m gpu > gpurun_out/pytest_frames.log 2>&1; echo "frames pytest rc=$?"; tail -n 3 gpurun_out/pytest_frames.log
timeout 300 python tools/frames_bench.py > gpurun_out/frames_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/frames_bench.log | tail -n 8
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_kernels_gpu.py -x -q -m gpu -k "phi3 or attention_bidirectional" > gpurun_out/pytest_new.log 2>&1; echo "new pytest rc=$?"; tail -n 5 gpurun_out/pytest_new.log
