#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_spec_kernel_gpu.py -x -q -m gpu -k "short" 2>&1 | tail -3
for fr in 64 128 256 512; do
python bench.py --batch 4096 --frames $fr --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('frames $fr:', round(j['value'],1), 'Msamples/s', round(j['ms_per_step']*1000,1), 'us/step')"
done 2>&1 | tee gpurun_out/r02v_short_call_variants.log
