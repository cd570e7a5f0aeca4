#!/bin/bash
# One GPU-box visit for the evidence under profiles/: the launch list of the bench command, one ncu --set full capture
# of the general kernel on wavenet_a2_max.nam (BASELINE.json config 3) and one of the lock-step kernel on a single
# 96,000-frame stream.  Reports land in gpurun_out/ and are summarised here with tools/ncu_summary.py.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
tail -2 gpurun_out/bench_under_ncu_$TAG.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wavenet_generic_kernel -c 1 -f \
  -o gpurun_out/prof_${TAG}_generic_a2_max python bench.py --model wavenet_a2_max --batch 4096 --frames 1024 --steps 1 \
  --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_generic_$TAG.log 2>&1
tail -2 gpurun_out/ncu_generic_$TAG.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel -c 1 -f \
  -o gpurun_out/prof_${TAG}_lockstep_1x96000 python bench.py --batch 1 --frames 96000 --steps 1 --warmup 3 \
  --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_lockstep_$TAG.log 2>&1
tail -2 gpurun_out/ncu_lockstep_$TAG.log | cut -c1-300
ls -la gpurun_out | tail -8
