#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_lat2 --launch-skip 300 -c 1 -f -o /tmp/prof_lat2 \
  python tools/latency_probe.py wavenet_a1_standard 2 > gpurun_out/ncu_lat2.log 2>&1
python tools/ncu_summary.py /tmp/prof_lat2.ncu-rep gpurun_out/r02o_lat2_kernel_one_64_frame_call > /dev/null 2>> gpurun_out/ncu_lat2.log
tail -2 gpurun_out/ncu_lat2.log
