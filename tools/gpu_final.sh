#!/bin/bash
# what the driver does at round end, in one visit: GPU tests, smoke(), the N = 1 bench line of both arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -10 | tee gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/final_bench_reference.json | cut -c1-300
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 3 2>gpurun_out/final_bench_err.log | tee gpurun_out/final_bench.json | cut -c1-300
