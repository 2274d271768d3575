#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 tools/_build/overlap_probe > gpurun_out/overlap_probe.log 2>&1
timeout 600 python tools/trace_pass.py 30 0,1,3,5 100 4 > gpurun_out/trace.log 2>&1
XYK_DEBUG_PASSES=1 timeout 300 python tools/prof_layer.py 30 3 > gpurun_out/passes.log 2>&1
tail -14 gpurun_out/passes.log; cat gpurun_out/trace.log | tail -5; tail -6 gpurun_out/overlap_probe.log
