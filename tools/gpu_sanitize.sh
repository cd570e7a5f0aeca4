#!/bin/bash
# compute-sanitizer over the kernels that changed late in round 2: short-call entry points (64 / 128 / 256 frames, L2 policy),
# the 8-group low-latency kernel, reserved SMs
mkdir -p gpurun_out
SEL='short_call_variant or short_call_variants or lat_kernel_plugin or reserved_sms'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_spec_kernel_gpu.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_spec_kernel_gpu.py -x -q -m gpu -k "short_call_variant_many or lat_kernel_plugin" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
