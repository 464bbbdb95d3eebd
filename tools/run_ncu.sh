#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one real bench step, (2) full-set captures of the hot kernels and of the
# memory-bound kernels at their bench shapes. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_step.csv python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1
for k in ${KERNELS:-swiglu attn_vit attn_llm layernorm rmsnorm gather_ln frames im2col}; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -c 1 -f \
      -o gpurun_out/prof_$k python tools/profile_step.py --kernel $k > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -20
