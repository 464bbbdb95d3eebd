#!/bin/bash
for k in "$@"; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -c 1 -f \
      -o gpurun_out/prof_$k python tools/profile_step.py --kernel $k > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
