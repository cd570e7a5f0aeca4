#!/bin/bash
# A/B on one box: the specialised kernel with the trimmed vs the literal fast-tanh denominator
mkdir -p gpurun_out
python - <<'PY'
s = open("neuralampmodelercore_b200/csrc/wavenet_spec.cuh").read()
a = "const u64 s = fma2(dup2(0.814642734961073f), x2, ax);"
b = "const u64 s = fma2(dup2(0.814642734961073f), mul2(x, ax), x) & kAbs;"
assert a in s
open("/tmp/wavenet_spec_literal.cuh", "w").write(s.replace(a, b))
PY
for rep in 1 2; do
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('trimmed', d['value'], d['ms_per_step'])"
  NAM_B200_SPEC_SOURCE=/tmp/wavenet_spec_literal.cuh python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('literal', d['value'], d['ms_per_step'])"
done 2>&1 | tee gpurun_out/ab_tanh.log
python tools/latency_probe.py wavenet_a1_standard 0 | tee gpurun_out/latency_probe.log
python tools/latency_probe.py lstm 0 | tee -a gpurun_out/latency_probe.log
python - <<'PY' | tee gpurun_out/lstm_probe.log
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import neuralampmodelercore_b200 as nb
from tests import nam_fixtures as fx
for b in (32, 4096, 16384):
    m = nb.get_dsp(fx.load_model("lstm"), batch=b, fast_tanh=True, jit=1)
    m.Reset(48000.0, 4096)
    x = torch.from_numpy(fx.synthetic_batch(b, 4096, seed=7)).cuda(); y = torch.empty_like(x)
    s = torch.cuda.Stream()
    for _ in range(3): m.process_batch_device(x.data_ptr(), y.data_ptr(), b, 4096, 4096, 4096, s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(4): m.process_batch_device(x.data_ptr(), y.data_ptr(), b, 4096, 4096, 4096, s.cuda_stream)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print({"lstm_batch": b, "ms_per_4096_frames": ms, "ns_per_step": ms * 1e6 / 4096, "Gsamples_per_s": b * 4096 / ms / 1e6})
PY
