#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_spec_kernel_gpu.py tests/test_parity_vs_reference_gpu.py tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_tail4.log
python - <<'PY' | tee gpurun_out/lstm_probe2.log
import sys, torch
sys.path.insert(0, ".")
import neuralampmodelercore_b200 as nb
from tests import nam_fixtures as fx
for geom in (1, 2):
  for b in (32, 4096, 16384):
    m = nb.get_dsp(fx.load_model("lstm"), batch=b, fast_tanh=True, jit=1, kernel_geometry=geom)
    m.Reset(48000.0, 4096)
    x = torch.from_numpy(fx.synthetic_batch(b, 4096, seed=7)).cuda(); y = torch.empty_like(x)
    s = torch.cuda.Stream()
    for _ in range(3): m.process_batch_device(x.data_ptr(), y.data_ptr(), b, 4096, 4096, 4096, s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(4): m.process_batch_device(x.data_ptr(), y.data_ptr(), b, 4096, 4096, 4096, s.cuda_stream)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print({"lstm_geometry": "gate_split" if geom == 1 else "thread_per_stream", "batch": b, "ns_per_step": ms * 1e6 / 4096, "Gsamples_per_s": b * 4096 / ms / 1e6})
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_err.log | tee gpurun_out/bench_line.json | cut -c1-300
