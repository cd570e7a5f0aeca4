#!/bin/bash
# ncu --set full of the specialised general kernel on wavenet_a2_max.nam, 4096 streams x 1024 frames (BASELINE.json config 3)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_generic_spec --launch-skip 3 -c 1 -f -o /tmp/prof_gen \
  python bench.py --model wavenet_a2_max --batch 4096 --frames 1024 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/ncu_gen.log 2>&1
python tools/ncu_summary.py /tmp/prof_gen.ncu-rep gpurun_out/r02l_general_spec_kernel_a2_max_batch4096 > /dev/null 2>> gpurun_out/ncu_gen.log
tail -2 gpurun_out/ncu_gen.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | cut -c1-200
