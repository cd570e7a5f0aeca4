"""One stream, the plugin protocol (tools/benchmodel.cpp:116-133: Reset(sr, 64), 1500 x 64-frame process() calls):
wall time per call through nam_b200_process_f32_planar and the kernel's own time (CUDA events of the handle)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import neuralampmodelercore_b200 as nb
from tests import nam_fixtures as fx

name = sys.argv[1] if len(sys.argv) > 1 else "wavenet_a1_standard"
jit = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nam = fx.load_model(name)
d = nb.get_dsp(nam, batch=1, fast_tanh=True, jit=jit)
d.Reset(48000.0, 64)
x = fx.synthetic_batch(1, 64 * 1500, seed=3)[0]
out = np.zeros(64, np.float32)
for i in range(50):
    d.process([x[i * 64:(i + 1) * 64]], [out], 64)
kms = []
t0 = time.perf_counter()
for i in range(1500):
    d.process([x[i * 64:(i + 1) * 64]], [out], 64)
    if i % 100 == 0:
        kms.append(d.last_kernel_ms())
dt = time.perf_counter() - t0
print({"model": name, "jit": d.jit_state, "wall_us_per_call": dt / 1500 * 1e6, "kernel_us_per_call": float(np.median(kms)) * 1e3,
       "total_ms_for_1500_calls": dt * 1e3})
