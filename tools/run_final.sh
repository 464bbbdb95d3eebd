#!/bin/bash
# final evidence refresh: smoke, launch list of one bench step under ncu, full captures, then clean bench runs
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_step.csv python tools/profile_step.py > gpurun_out/ncu_step.log 2>&1
for k in swiglu gelu attn_vit attn_llm; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -c 1 -f \
      -o gpurun_out/prof_$k python tools/profile_step.py --kernel $k > gpurun_out/ncu_$k.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 --breakdown gpurun_out/breakdown_final.json > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
