#!/bin/bash
# GPU call: step timeline of attention2 (tools/attn_trace.py) only, plus the llm / mha timing
mkdir -p gpurun_out
timeout 300 python tools/attn_trace.py llm > gpurun_out/attn_trace.log 2>&1; echo "trace rc=$?"
timeout 300 python tools/attn_bench.py llm mha > gpurun_out/attn_big.log 2>&1; echo "big rc=$?"
grep -v flash gpurun_out/attn_big.log | tail -n 6
