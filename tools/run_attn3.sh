for cfg in "1 128 1" "1 192 1" "1 129 1" "1 200 1" "1 256 1" "1 257 1" "1 320 2"; do
  CUDA_LAUNCH_BLOCKING=1 timeout 60 python tools/diag_attn3.py $cfg 2>&1 | tail -3
done
