#!/bin/bash
# DRAM bytes + duration of the tile pass launches of one layer under different XYK_ABLATE settings
n=${1:-28}; shift
for ab in "$@"; do
  XYK_ABLATE=$ab timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tile_pass -s 12 -c 12 --csv --log-file gpurun_out/dram_ab$ab.csv python tools/prof_layer.py $n 2 > /dev/null 2>&1
  python - "$ab" <<'PY'
import csv, sys
ab = sys.argv[1]
rows = [r for r in csv.reader(open(f"gpurun_out/dram_ab{ab}.csv")) if len(r) > 10]
hdr = rows[0]; i_m = hdr.index("Metric Name"); i_v = hdr.index("Metric Value"); i_id = hdr.index("ID")
acc = {}
for r in rows[1:]:
    acc.setdefault(r[i_m], []).append(float(r[i_v].replace(",", "")))
n = len(acc["gpu__time_duration.sum"])
print(f"ablate={ab}: launches={n} read={sum(acc['dram__bytes_read.sum'])/n:.3f} write={sum(acc['dram__bytes_write.sum'])/n:.3f} (units as ncu prints) time={sum(acc['gpu__time_duration.sum'])/n:.4f}")
PY
done
