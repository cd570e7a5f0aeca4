#!/bin/bash
# ncu --set full capture of one prototype variant, summarised on the box (reports are ~50 MB)
# usage: gpurun --timeout 900 -- 'bash tools/spec_proto/profile.sh spec_proto_512x1x2 TAG'
BIN=${1:-spec_proto_512x1x2}; TAG=${2:-r02a}
mkdir -p gpurun_out
cd tools/spec_proto
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_spec --launch-skip 2 -c 1 -f -o /tmp/prof_spec \
  ./$BIN 1184 4096 1 > ../../gpurun_out/ncu_spec_$TAG.log 2>&1
cd ../..
python tools/ncu_summary.py /tmp/prof_spec.ncu-rep gpurun_out/${TAG}_spec_proto > /dev/null 2>> gpurun_out/ncu_spec_$TAG.log
ncu -i /tmp/prof_spec.ncu-rep --page raw --csv > gpurun_out/${TAG}_spec_proto_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
