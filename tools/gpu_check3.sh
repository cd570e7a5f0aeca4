#!/bin/bash
# latency work: lat-kernel tests, the C++ multi-GPU host, the probe and the reference's benchmodel on the shim
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_tail3.log
python tools/latency_probe.py wavenet_a1_standard 3 | tee gpurun_out/latency_probe_after.log
python tools/latency_probe.py wavenet_a1_standard 2 | tee -a gpurun_out/latency_probe_after.log
python - <<'PY'
import json
from tests import nam_fixtures as fx
for n in ("wavenet_a1_standard", "wavenet", "lstm"):
    open(f"/tmp/{n}.nam", "w").write(json.dumps(fx.load_model(n)))
PY
for rep in 1 2 3; do build/ref_tools/benchmodel /tmp/wavenet_a1_standard.nam | tail -1; done | tee gpurun_out/benchmodel_a1.log

build/ref_tools/benchmodel /tmp/lstm.nam | tail -1 | sed 's/^/lstm: /' | tee -a gpurun_out/benchmodel_a1.log
NAM_B200_LAT_KERNEL=precompiled build/ref_tools/benchmodel /tmp/wavenet_a1_standard.nam | tail -1 | sed 's/^/precompiled lat2: /' | tee -a gpurun_out/benchmodel_a1.log
NAM_B200_LAT_KERNEL=off NAM_B200_JIT=0 build/ref_tools/benchmodel /tmp/wavenet_a1_standard.nam | tail -1 | sed 's/^/neither (128x1 geometry): /' | tee -a gpurun_out/benchmodel_a1.log
