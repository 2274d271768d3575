#!/bin/bash
# ncu full-set capture of one tile pass (launch index $1 among tile_pass launches) + source-page CSV
mkdir -p gpurun_out
SKIP=${1:-13}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass -s $SKIP -c 1 -o gpurun_out/tp_one -f python tools/prof_layer.py 30 2 > gpurun_out/ncu_one.log 2>&1
ncu -i gpurun_out/tp_one.ncu-rep --page source --csv --print-source sass > gpurun_out/tp_one_src.csv 2>/dev/null
ncu -i gpurun_out/tp_one.ncu-rep --page raw --csv > gpurun_out/tp_one_raw.csv 2>/dev/null
ls -la gpurun_out/tp_one*; tail -2 gpurun_out/ncu_one.log
