"""Summarise an ncu --set full capture (exported with `ncu -i X.ncu-rep --page raw --csv`) into a small JSON for profiles/."""
import csv, json, sys
src, dst, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg"] + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
idx = [hdr.index(w) for w in want if w in hdr]
out = {"note": note, "units": {hdr[i]: units[i] for i in idx}, "launches": []}
for r in rows[2:]:
    d = {}
    for i in idx:
        try: d[hdr[i]] = float(r[i])
        except ValueError: d[hdr[i]] = r[i]
    out["launches"].append(d)
json.dump(out, open(dst, "w"), indent=1)
print("wrote", dst, len(out["launches"]), "launches")
