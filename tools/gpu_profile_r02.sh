#!/bin/bash
# Round-2 evidence in one GPU-box visit: the launch list of the bench command and an `ncu --set full` capture of the headline
# kernel at the bench shape (4096 x 4096), summarised on the box (reports are ~50 MB).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_spec_kernel --launch-skip 3 -c 1 -f -o /tmp/prof_head \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_head.log 2>&1
python tools/ncu_summary.py /tmp/prof_head.ncu-rep gpurun_out/r02_spec_kernel_bench_shape > /dev/null 2>> gpurun_out/ncu_head.log
tail -2 gpurun_out/ncu_head.log
