#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/decision_probe.py 64 24 > gpurun_out/decision_probe.log 2>&1; echo "probe rc=$?"
tail -n 3 gpurun_out/decision_probe.log | cut -c1-1500
timeout 400 python bench.py --no-cpu-baseline --breakdown gpurun_out/breakdown_attn2.json > gpurun_out/bench_attn2.json 2> gpurun_out/bench_attn2.err; echo "bench2 rc=$?"
MJV_ATTN_IMPL=1 timeout 400 python bench.py --no-cpu-baseline --breakdown gpurun_out/breakdown_attn1.json > gpurun_out/bench_attn1.json 2> gpurun_out/bench_attn1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in ("attn2","attn1"):
    try:
        b=json.load(open(f"gpurun_out/bench_{n}.json")); k=json.load(open(f"gpurun_out/breakdown_{n}.json"))["kernels"]
        print(n, "value", round(b["value"],2), "ms", round(b["ms_per_step"],1), "e2e", round(b["e2e"]["value"],2), "clk", b["clocks"]["sm_mhz"],
              "attn_bidir", round(k["attn_bidir"]["ms"],1), "attn_causal", round(k["attn_causal"]["ms"],1), "frac", round(b["roofline"]["frac"],3))
    except Exception as e: print(n, "failed", e)
PY
