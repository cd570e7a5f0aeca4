#!/bin/bash
# 2-GPU visit: spec tests on GPU 0, then the scaling bench line at N = 2 (with the NCCL gather pass)
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_spec_kernel_gpu.py tests/test_sharding_gloo.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_tail2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/bench2_err.log | tee gpurun_out/bench_line_n2.json | cut -c1-600
tail -3 gpurun_out/bench2_err.log
