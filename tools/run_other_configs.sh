#!/bin/bash
mkdir -p gpurun_out
timeout 700 python bench.py --model 4b --no-cpu-baseline --breakdown gpurun_out/breakdown_4b.json > gpurun_out/bench_4b.json 2> gpurun_out/bench_4b.err; echo "4b rc=$?"
timeout 700 python bench.py --workload long16x7 --no-cpu-baseline --breakdown gpurun_out/breakdown_long.json > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; echo "long rc=$?"
python - <<'PY'
import json
for n in ("4b","long"):
    try:
        b=json.load(open(f"gpurun_out/bench_{n}.json")); k=json.load(open(f"gpurun_out/breakdown_{n}.json"))["kernels"]
        print(n, "value", round(b["value"],2), "ms", round(b["ms_per_step"],1), "e2e", round(b["e2e"]["value"],2), "clk", b["clocks"]["sm_mhz"], "step_tflops", round(b["step_tflops"],1),
              "attn_bidir", round(k["attn_bidir"]["ms"],1), round(k["attn_bidir"]["rate_T_per_s"],1), "attn_causal", round(k["attn_causal"]["ms"],1), round(k["attn_causal"]["rate_T_per_s"],1))
    except Exception as e: print(n, "failed", e)
PY
tail -n 2 gpurun_out/bench_4b.err gpurun_out/bench_long.err
