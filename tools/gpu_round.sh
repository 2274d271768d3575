#!/bin/bash
# one GPU-box visit: parity tests, bench line, per-pass timing, ncu launch list, ncu full-set capture of tile passes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
XYK_DEBUG_PASSES=1 timeout 300 python tools/prof_layer.py 30 3 > gpurun_out/passes.log 2>&1
tail -16 gpurun_out/passes.log
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_pass -s ${NCU_SKIP:-28} -c ${NCU_COUNT:-4} -o gpurun_out/tile_pass_full -f python tools/prof_layer.py 30 3 > gpurun_out/ncu_full.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
