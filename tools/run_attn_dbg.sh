#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_trace.py vit > gpurun_out/attn_trace.log 2>&1; echo "trace vit rc=$?"
timeout 300 python tools/attn_trace.py llm >> gpurun_out/attn_trace.log 2>&1; echo "trace llm rc=$?"
cat gpurun_out/attn_trace.log | tail -60
