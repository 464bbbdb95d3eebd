#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fp32_mode_gpu.py -x -q -m gpu -s > gpurun_out/pytest_fp32.log 2>&1; echo "fp32 rc=$?"
tail -n 25 gpurun_out/pytest_fp32.log | cut -c1-400
timeout 900 python -m pytest tests/test_decisions_gpu.py -x -q -m gpu -s > gpurun_out/pytest_decisions.log 2>&1; echo "decisions rc=$?"
grep -E "DECISION_PARITY|passed|failed|Error|assert" gpurun_out/pytest_decisions.log | cut -c1-2500
timeout 600 python -m pytest tests/test_forward_gpu.py -x -q -m gpu -k "dropin or fp32_pixel" > gpurun_out/pytest_fwd_sel.log 2>&1; echo "fwd-sel rc=$?"
tail -n 4 gpurun_out/pytest_fwd_sel.log
