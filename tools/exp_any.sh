#!/bin/bash
mkdir -p gpurun_out
{
if [ "${PARITY:-0}" = "1" ]; then
echo "== parity, default config"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "== parity, min_left 2 + beam 3"; XYK_SPLIT_MIN_LEFT=2 XYK_BEAM=3 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
fi
echo "== timing default"; XYK_DEBUG_PASSES=1 python tools/prof_layer.py 30 2 2>&1 | tail -15
echo "== timing min_left 2 beam 3"; XYK_SPLIT_MIN_LEFT=2 XYK_BEAM=3 XYK_DEBUG_PASSES=1 python tools/prof_layer.py 30 2 2>&1 | tail -15
echo "== timing forced UNB kernel, default plan"; XYK_FORCE_UNB=1 XYK_DEBUG_PASSES=1 python tools/prof_layer.py 30 2 2>&1 | tail -15
} > gpurun_out/exp_any.log 2>&1
cat gpurun_out/exp_any.log
