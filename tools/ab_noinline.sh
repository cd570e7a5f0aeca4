#!/bin/bash
# A/B on one box: the specialised kernel with the warp-0 history copies inlined (shipped) vs out of line
mkdir -p gpurun_out
python - <<'PY'
s = open("neuralampmodelercore_b200/csrc/wavenet_spec.cuh").read()
for name in ("hist_load(", "hist_store("):
    a = "template <int P, int L, int RMASK, int W>\n__device__ __forceinline__ void " + name
    assert a in s
    s = s.replace(a, "template <int P, int L, int RMASK, int W>\n__device__ __noinline__ void " + name)
open("/tmp/wavenet_spec_noinline.cuh", "w").write(s)
PY
for rep in 1 2; do
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('inlined', d['value'], d['ms_per_step'])"
  NAM_B200_SPEC_SOURCE=/tmp/wavenet_spec_noinline.cuh python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('out_of_line', d['value'], d['ms_per_step'], d['roofline']['jit'])"
done 2>&1 | tee gpurun_out/ab_noinline.log
