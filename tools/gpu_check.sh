#!/bin/bash
# One GPU-box visit while developing: selected tests, then a short bench line.  usage: bash tools/gpu_check.sh "<pytest args>" "<bench args>"
mkdir -p gpurun_out
timeout 1500 python -m pytest $1 -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_tail.log
timeout 600 python bench.py --steps 20 --warmup 3 $2 2>gpurun_out/bench_err.log | tee gpurun_out/bench_line.json | cut -c1-1500
tail -5 gpurun_out/bench_err.log
