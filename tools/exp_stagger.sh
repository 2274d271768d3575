#!/bin/bash
# experiment: de-phase the three co-resident CTAs of an SM (results stay correct)
mkdir -p gpurun_out
{
python tools/ablate_time.py 30 0
for ns in 1500 3000 4000 5000 6500 8000; do
  echo "stagger ns=$ns"; XYK_STAGGER_NS=$ns python tools/ablate_time.py 30 256
done
XYK_STAGGER_NS=4000 XYK_ABLATE=256 XYK_DEBUG_PASSES=1 python tools/prof_layer.py 30 3 2>&1 | tail -15
} > gpurun_out/exp_stagger.log 2>&1
cat gpurun_out/exp_stagger.log
