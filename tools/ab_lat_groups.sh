#!/bin/bash
# A/B on one box: the low-latency kernel with 4 and with 8 channel groups (instruction streams)
mkdir -p gpurun_out
python -m pytest tests/test_spec_kernel_gpu.py tests/test_dropin_tools_gpu.py -x -q -m gpu -k "lat or dropin or tool" 2>&1 | tail -3
python - <<'PY'
import json
from tests import nam_fixtures as fx
for n in ("wavenet_a1_standard",):
    open(f"/tmp/{n}.nam", "w").write(json.dumps(fx.load_model(n)))
PY
for rep in 1 2; do
for g in 4 8; do
echo "groups $g: $(NAM_B200_LAT_GROUPS=$g python tools/latency_probe.py wavenet_a1_standard 3 2>&1 | tail -1)"
echo "groups $g benchmodel: $(NAM_B200_LAT_GROUPS=$g build/ref_tools/benchmodel /tmp/wavenet_a1_standard.nam | tail -1)"
done; done 2>&1 | tee gpurun_out/r02y_ab_lat_groups.log
