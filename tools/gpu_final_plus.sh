#!/bin/bash
# round-end validation (tools/gpu_final.sh) followed by the reference's benchmodel on the shim, three times
bash tools/gpu_final.sh
python - <<'PY'
import json
from tests import nam_fixtures as fx
open("/tmp/wavenet_a1_standard.nam", "w").write(json.dumps(fx.load_model("wavenet_a1_standard")))
PY
for rep in 1 2 3; do echo "benchmodel: $(build/ref_tools/benchmodel /tmp/wavenet_a1_standard.nam | tail -1)"; done | tee gpurun_out/final_benchmodel.log
echo "precompiled low-latency kernel: $(NAM_B200_LAT_KERNEL=precompiled build/ref_tools/benchmodel /tmp/wavenet_a1_standard.nam | tail -1)" | tee -a gpurun_out/final_benchmodel.log
