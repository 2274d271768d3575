#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/parity.log; cat gpurun_out/parity.log
bash tools/ab.sh grp1 grp3
XYK_DEBUG_PASSES=1 timeout 300 python tools/prof_layer.py 30 3 2>&1 | tail -14 > gpurun_out/passes.log; cat gpurun_out/passes.log
