#!/usr/bin/env bash
# two GPUs, launched like the driver does: the reference arm (rank 0 alone works) and the bench line with the NCCL gather pass
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/n2_bench_reference.json 2> gpurun_out/n2_bench_reference.err
echo "reference arm rc=$?"; cut -c1-200 gpurun_out/n2_bench_reference.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/n2_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/n2_bench.json"):
    l = l.strip()
    if l.startswith("{"):
        j = json.loads(l)
        print({k: j.get(k) for k in ("value", "ms_per_step", "value_with_gather", "n_gpus", "e2e")})
        print(j.get("nccl_gather"))
PY
