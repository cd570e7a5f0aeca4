#!/bin/bash
# One GPU-box visit: the new tests first, the whole GPU suite, the default bench line.  Everything lands in gpurun_out/.
# usage (from the build container): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tile_parallel_gpu.py tests/test_multichannel_gpu.py tests/test_convnet_gpu.py tests/test_slimmable_wavenet_gpu.py tests/test_generic_kernel_gpu.py -q -m gpu 2>&1 | tail -30 > gpurun_out/tests_new_$TAG.log
cat gpurun_out/tests_new_$TAG.log
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/tests_all_$TAG.log
cat gpurun_out/tests_all_$TAG.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"])
for k, v in d["secondary"].items():
    print(k, v.get("Msamples_per_s"), v.get("error"))
PY
tail -3 gpurun_out/bench_$TAG.err
# the reference's own benchmodel (1 stream, 1500 x 64-frame process() calls, tools/benchmodel.cpp) on the shim
if [ -x build/ref_tools/benchmodel ]; then
  python - <<PY
import json
from tests import nam_fixtures as fx
json.dump(fx.load_model("wavenet_a1_standard"), open("/tmp/a1.nam", "w"))
PY
  ./build/ref_tools/benchmodel /tmp/a1.nam 2>&1 | tail -3
  ./build/ref_tools/benchmodel_bufsize /tmp/a1.nam 1024 20 2>&1 | tail -1   # (one iteration = 2 s of audio)
fi
