#!/bin/bash
mkdir -p gpurun_out
{
for ab in ${ABLATES:-0 1}; do
  echo "== ablate $ab $EXTRA"; XYK_ABLATE=$ab XYK_DEBUG_PASSES=1 python tools/prof_layer.py 30 2 2>&1 | tail -${TAILN:-24}
done
} > gpurun_out/exp_ablate.log 2>&1
cat gpurun_out/exp_ablate.log
