#!/bin/bash
# run every prototype variant on the GPU box; one JSON line each
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
for b in spec_proto_*x*; do
  timeout 120 ./$b 4096 4096 5 2>&1 | tail -2
done | tee ../../gpurun_out/spec_proto.jsonl
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
