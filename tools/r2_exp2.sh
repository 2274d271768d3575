#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/parity.log
XYK_DEBUG_PASSES=1 timeout 300 python tools/prof_layer.py 30 3 > gpurun_out/passes.log 2>&1
timeout 600 python tools/trace_pass.py 30 0,1 100 4 > gpurun_out/trace.log 2>&1
cat gpurun_out/parity.log; tail -14 gpurun_out/passes.log; tail -3 gpurun_out/trace.log
