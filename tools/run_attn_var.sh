#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4
for v in 0 1 5; do MJV_ATTN_VARIANT=$v timeout 300 python tools/exp_attn_var.py 2>&1 | tail -2; done
