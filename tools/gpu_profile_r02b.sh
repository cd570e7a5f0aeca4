#!/bin/bash
# Round-2 evidence, second visit: the launch list of the bench command after the short-call work and an `ncu --set full`
# capture of the short-call kernel WITH the L2 evict-first policy (profiles/r02t_* is the same kernel before it)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02z_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_spec_short_kernel --launch-skip 5 -c 1 -f -o /tmp/prof_sshort \
  python bench.py --batch 4096 --frames 64 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_sshort.log 2>&1
python tools/ncu_summary.py /tmp/prof_sshort.ncu-rep gpurun_out/r02z_spec_short_kernel_with_l2_policy > /dev/null 2>> gpurun_out/ncu_sshort.log
ncu -i /tmp/prof_sshort.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2]
for k in ('gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):
    if k in h: print(k, v[h.index(k)])
" | tee gpurun_out/r02z_spec_short_extra.txt
python -c "
import json
j=json.load(open('gpurun_out/r02z_spec_short_kernel_with_l2_policy.json'))
print(j['stall_cycles_per_issued_instruction'])"
tail -2 gpurun_out/ncu_sshort.log
