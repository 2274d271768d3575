#!/bin/bash
mkdir -p gpurun_out
{
if [ "${PARITY:-0}" = "1" ]; then timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3; fi
XYK_DEBUG_PASSES=1 python tools/prof_layer.py 30 3 2>&1 | tail -${TAILN:-15}
} > gpurun_out/exp_cur.log 2>&1
cat gpurun_out/exp_cur.log
