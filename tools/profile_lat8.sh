#!/bin/bash
# ncu --set full of the compiled low-latency kernel with 8 channel groups (one 64-frame call of one stream); r02f_* is the
# 4-group kernel.  --cache-control none keeps L2 warm between the replayed passes, as in real use (code is L2-resident there)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:wavenet_lat_kernel --launch-skip 300 -c 1 -f -o /tmp/prof_lat8 \
  python tools/latency_probe.py wavenet_a1_standard 3 > gpurun_out/ncu_lat8.log 2>&1
python tools/ncu_summary.py /tmp/prof_lat8.ncu-rep gpurun_out/r02ze_lat_kernel_8_groups_one_64_frame_call > /dev/null 2>> gpurun_out/ncu_lat8.log
python -c "
import json
j=json.load(open('gpurun_out/r02ze_lat_kernel_8_groups_one_64_frame_call.json'))
print(j['stall_cycles_per_issued_instruction']); print(j['warp_instructions_executed'], j.get('block'))
m=j['metrics']
for k in m:
    if any(s in k for s in ('issue_active','warps_active','duration','cycles_elapsed.max')): print(k, m[k]['value'])
"
tail -2 gpurun_out/ncu_lat8.log
