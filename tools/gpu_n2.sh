#!/usr/bin/env bash
# two GPUs: the bench line with the NCCL gather, with and without reserved SMs' effect visible in value_with_gather
mkdir -p gpurun_out
python -m pytest tests/test_spec_kernel_gpu.py -x -q -m gpu -k "reserved" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -c 1500 gpurun_out/n2_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/n2_bench.json"):
    l = l.strip()
    if l.startswith("{"):
        j = json.loads(l)
        print({k: j.get(k) for k in ("value", "ms_per_step", "value_with_gather", "nccl_gather", "n_gpus")})
PY
