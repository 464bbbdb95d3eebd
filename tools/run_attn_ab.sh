#!/bin/bash
# GPU call: correctness of attention2 on odd shapes, the attention unit tests, then the A/B timing against generation 1 and flash-attn 2
mkdir -p gpurun_out
timeout 300 python tools/attn_bench.py small > gpurun_out/attn_small.log 2>&1; echo "small rc=$?"
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k attention > gpurun_out/attn_pytest.log 2>&1; echo "pytest-attn rc=$?"
timeout 600 python tools/attn_bench.py vit llm mha > gpurun_out/attn_big.log 2>&1; echo "big rc=$?"
timeout 300 python tools/attn_trace.py vit llm > gpurun_out/attn_trace.log 2>&1; echo "trace rc=$?"
grep -E "impl2" gpurun_out/attn_small.log | awk '{print $1,$2,$3,$4,$9,$10}' | tr '\n' ';'; echo; tail -n 2 gpurun_out/attn_pytest.log; grep -v flash gpurun_out/attn_big.log | tail -n 8; python - <<'PY'
import json
d=json.load(open('gpurun_out/attn_trace.json'))
for k,v in d.items():
    print(k, {r:{a:int(b) for a,b in x.items()} for r,x in v.items()})
PY
