"""Dynamic instruction mix of one profiled launch (ncu --page source --csv --print-source sass export)."""
import csv, re, collections, sys
path = sys.argv[1]; ntile = float(sys.argv[2]) if len(sys.argv) > 2 else 2**18 * 4
rows = list(csv.reader(open(path)))
hdr = rows[1]
iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples")
iW = hdr.index("L1 Wavefronts Shared"); iWI = hdr.index("L1 Wavefronts Shared Ideal")
ops = collections.Counter(); samp = collections.Counter(); tot = 0
wf = collections.Counter(); wfi = collections.Counter()
for r in rows[2:]:
    if len(r) <= iE: continue
    s = r[iS].strip(); m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', s)
    if not m: continue
    op = m.group(2); e = int(r[iE] or 0); ops[op] += e; tot += e; samp[op] += int(r[iSamp] or 0)
    wf[op] += int(r[iW] or 0); wfi[op] += int(r[iWI] or 0)
print("total", tot, tot / ntile, "samples", sum(samp.values()))
for op, c in ops.most_common(32):
    print(f"{op:10s} {c/ntile:9.1f} per warp-tile  samples {samp[op]:7d}  wf {wf[op]/ntile:8.1f} ideal {wfi[op]/ntile:8.1f}")
