#!/bin/bash
# One GPU-box visit for the evidence under profiles/: the launch list of the bench command, one ncu --set full capture
# of the general kernel on wavenet_a2_max.nam (BASELINE.json config 3) and one of the lock-step kernel on a single
# 96,000-frame stream.  The reports (~50 MB each) are summarised ON THE BOX with tools/ncu_summary.py and deleted:
# gpurun brings back at most 64 MiB.
# usage: gpurun --timeout 1200 -- 'bash tools/gpu_profile.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
tail -1 gpurun_out/bench_under_ncu_$TAG.log | cut -c1-200
capture() { # name, kernel regex, bench arguments...
  local name=$1 regex=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex --launch-skip ${SKIP:-0} -c 1 -f -o /tmp/prof_$name \
    python bench.py "$@" --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_${name}_$TAG.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$name.ncu-rep gpurun_out/${TAG}_$name > /dev/null 2>> gpurun_out/ncu_${name}_$TAG.log
  rm -f /tmp/prof_$name.ncu-rep
}
SKIP=3 capture generic_a2_max wavenet_generic_kernel --model wavenet_a2_max --batch 4096 --frames 1024
capture lockstep_1x96000 wavenet_fused_kernel --batch 1 --frames 96000
ls -la gpurun_out | tail -8
