#!/bin/bash
for cfg in "513 1 1" "296,77 1 1" "300,77 1 1" "301,77 1 1" "300,77,513 1 1" "304,80,513 4 2"; do
  set -- $cfg
  echo "=== lens=$1 H=$2 Hkv=$3"
  CUDA_LAUNCH_BLOCKING=1 timeout 60 python tools/diag_attn2.py $1 $2 $3 2>&1 | grep -v "^  \|Traceback\|warn" | tail -3
done
