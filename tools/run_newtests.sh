#!/bin/bash
# GPU call: the tests of the native head_dim-96 path, then the configs[4] bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_forward_gpu.py -x -q -m gpu -k "96 or config5 or narrow or qkv_llm" > gpurun_out/pytest_new.log 2>&1; echo "rc=$?"; tail -n 14 gpurun_out/pytest_new.log
timeout 700 python bench.py --model 4b --no-cpu-baseline --breakdown gpurun_out/breakdown_4b.json > gpurun_out/bench_4b.json 2> gpurun_out/bench_4b.err; echo "4b rc=$?"
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_4b.json")); k=json.load(open("gpurun_out/breakdown_4b.json"))["kernels"]
print("4b value", round(b["value"],2), "ms", round(b["ms_per_step"],1), "clk", b["clocks"]["sm_mhz"])
print({a:(round(v["ms"],2), round(v.get("rate_T_per_s",0),1)) for a,v in k.items() if v["ms"]>5})
PY
