#!/bin/bash
# ncu --set full of the model-specialised short-call kernel: batch 4096 x 64-frame calls
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_spec_short_kernel --launch-skip 5 -c 1 -f -o /tmp/prof_sshort \
  python bench.py --batch 4096 --frames 64 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_sshort.log 2>&1
python tools/ncu_summary.py /tmp/prof_sshort.ncu-rep gpurun_out/r02t_spec_short_kernel_batch4096_64frames > /dev/null 2>> gpurun_out/ncu_sshort.log
ncu -i /tmp/prof_sshort.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2]
for k in ('gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):
    if k in h: print(k, v[h.index(k)])
" | tee gpurun_out/r02t_spec_short_extra.txt
cat gpurun_out/r02t_spec_short_kernel_batch4096_64frames.json | head -c 3000
tail -2 gpurun_out/ncu_sshort.log
