#!/bin/bash
mkdir -p gpurun_out
echo "== impl1"; MJV_ATTN_IMPL=1 timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -k "full_size_2b_row_vs_oracle" 2>&1 | tail -n 4
echo "== impl2"; timeout 900 python -m pytest tests/test_forward_gpu.py -q -m gpu -k "full_size_2b_row_vs_oracle" 2>&1 | tail -n 4
