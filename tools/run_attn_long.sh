#!/bin/bash
# GPU call: causal attention timing on default-shaped and configs[3]-shaped rows
mkdir -p gpurun_out
timeout 600 python tools/attn_bench.py llm llmlong mha > gpurun_out/attn_big.log 2>&1; echo "big rc=$?"
cat gpurun_out/attn_big.log | tail -n 12
