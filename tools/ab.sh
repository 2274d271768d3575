#!/bin/bash
# same-box A/B timing of library builds: tools/ab.sh name1 name2 ... (tools/_build/libxyk_<name>.so), two interleaved rounds
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    XYK_LIB_AB=tools/_build/libxyk_$v.so timeout 300 python tools/prof_layer_ab.py 30 2>&1 | tail -1
  done
done | tee gpurun_out/ab.log
