#!/bin/bash
# first contact with the B200: parity tests, chunk-size micro-benchmark, a timing of the n=30 layer
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/chunk_bw.cu -o gpurun_out/chunk_bw && timeout 300 ./gpurun_out/chunk_bw > gpurun_out/chunk_bw.log 2>&1
cat gpurun_out/chunk_bw.log
timeout 900 python tools/quick_time.py > gpurun_out/quick_time.log 2>&1
cat gpurun_out/quick_time.log
