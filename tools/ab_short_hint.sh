#!/bin/bash
# A/B on one box: the short-call variant (4096 streams x 64-frame calls) before / with L2 evict-first hints / with hints and
# next-layer L2 prefetch (the shipped source)
mkdir -p gpurun_out
run() { python bench.py --batch 4096 --frames 64 --steps 300 --warmup 20 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('$1', round(j['value'],1), 'Msamples/s', round(j['ms_per_step']*1000,1), 'us/step')"; }
for rep in 1 2; do
NAM_B200_SPEC_SOURCE=$PWD/tools/spec_proto/ab_before.cuh run before
run hint_only_shipped
NAM_B200_SPEC_SOURCE=$PWD/tools/spec_proto/ab_bulk_prefetch.cuh run hint_and_bulk_prefetch_evict_first
done 2>&1 | tee gpurun_out/r02u_ab_short_call_l2_hints_2.log
python -m pytest tests/test_spec_kernel_gpu.py -x -q -m gpu 2>&1 | tail -3
