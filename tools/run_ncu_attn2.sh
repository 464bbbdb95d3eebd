#!/bin/bash
# ncu --set full with source counters on attention2 (ViT shape and LLM causal shape); reports come back in gpurun_out/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 4 -c 1 -f -o gpurun_out/prof_attn2_vit \
    python tools/attn_bench.py vit > gpurun_out/ncu_attn2_vit.log 2>&1; echo "vit rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention2 -s 4 -c 1 -f -o gpurun_out/prof_attn2_llm \
    python tools/attn_bench.py llm > gpurun_out/ncu_attn2_llm.log 2>&1; echo "llm rc=$?"
ls -la gpurun_out/*.ncu-rep
