#!/bin/bash
# GPU call: one ncu --set full capture of a single kernel shape from tools/profile_step.py; KERNEL=<name>
mkdir -p gpurun_out
k=${KERNEL:-resize720}
ncu --profile-from-start off --set full --clock-control none --import-source on -c 1 -f \
    -o gpurun_out/prof_$k python tools/profile_step.py --kernel $k > gpurun_out/ncu_$k.log 2>&1
echo "$k rc=$?"; ls -la gpurun_out/prof_$k.ncu-rep
