#!/bin/bash
mkdir -p gpurun_out
cd tools/spec_proto
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wavenet_lat --launch-skip 50 -c 1 -f -o /tmp/prof_lat ./lat_proto_fw2 1 64 20 > ../../gpurun_out/ncu_lat.log 2>&1
cd ../..
python tools/ncu_summary.py /tmp/prof_lat.ncu-rep gpurun_out/r02f_lat_proto > /dev/null 2>> gpurun_out/ncu_lat.log
tail -3 gpurun_out/ncu_lat.log
