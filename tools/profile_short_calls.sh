#!/bin/bash
# ncu --set full of the short-call geometry: batch 4096 x 64-frame calls (the reference's calling convention at scale)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_fused_kernel --launch-skip 67 -c 1 -f -o /tmp/prof_short \
  python bench.py --batch 4096 --frames 64 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --jit 2 > gpurun_out/ncu_short.log 2>&1
python tools/ncu_summary.py /tmp/prof_short.ncu-rep gpurun_out/r02i_short_call_kernel_batch4096_64frames > /dev/null 2>> gpurun_out/ncu_short.log
ncu -i /tmp/prof_short.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[2]
for k in ('lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_bytes.sum','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed'):
    if k in h: print(k, v[h.index(k)])
" | tee gpurun_out/r02i_short_call_extra.txt
tail -2 gpurun_out/ncu_short.log
