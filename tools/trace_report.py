"""Summarise a trace written by tools/trace_pass.py (host only)."""
import sys
import numpy as np

def load(path):
    a = np.load(path)
    out = {}
    for slot in range(a.shape[0]):
        ev = a[slot]
        if ev[2] == 0:
            continue
        sm = int(ev[0])
        cta, warp = slot // 4, slot % 4
        e = [(int(x >> np.uint64(8)), int(x & np.uint64(255))) for x in ev[2:] if x != 0]
        out.setdefault(sm, {})[(cta, warp)] = e
    return out

def phases(e):
    """[(name, t0, t1)] from an event list"""
    ph = []
    prev_t, prev_c = None, None
    for t, c in e:
        if prev_t is not None:
            name = {(1, 2): "tilewait", (2, 4): "load0", (3, 4): "load", (4, 5): "gates", (5, 6): "store", (6, 3): "barrier",
                    (6, 1): "tail"}.get((prev_c, c), f"{prev_c}->{c}")
            ph.append((name, prev_t, t))
        prev_t, prev_c = t, c
    return ph

if __name__ == "__main__":
    tr = load(sys.argv[1])
    sm_show = int(sys.argv[2]) if len(sys.argv) > 2 else sorted(tr)[0]
    agg = {}
    for sm, d in tr.items():
        for key, e in d.items():
            for name, t0, t1 in phases(e):
                agg.setdefault(name, []).append(t1 - t0)
    print("phase durations over all warps (cycles): mean / p10 / p50 / p90 / count")
    for name, v in sorted(agg.items()):
        v = np.array(v)
        print(f"  {name:10s} {v.mean():8.0f} {np.percentile(v, 10):8.0f} {np.percentile(v, 50):8.0f} {np.percentile(v, 90):8.0f} {len(v):7d}")
    # per tile span (code 1 -> next code 1)
    spans = []
    for sm, d in tr.items():
        for key, e in d.items():
            ts = [t for t, c in e if c == 1]
            spans += list(np.diff(ts))
    if spans:
        print(f"tile period per CTA: mean {np.mean(spans):.0f} cycles (x3 CTAs/SM -> {np.mean(spans) / 3:.0f} per tile)")
    # concurrency on every SM: fraction of time k CTAs are in a shared-memory phase / in the gate phase (warp 0 of each CTA)
    conc = {"lsu": np.zeros(5), "gates": np.zeros(5)}
    for sm, d in tr.items():
        iv = {"lsu": [], "gates": []}
        lo, hi = None, None
        for (cta, warp), e in d.items():
            if warp != 0:
                continue
            ph = phases(e)
            if not ph:
                continue
            s, t = ph[0][1], ph[-1][2]
            lo = s if lo is None else max(lo, s)
            hi = t if hi is None else min(hi, t)
            for name, t0, t1 in ph:
                if name in ("load", "load0", "store"):
                    iv["lsu"].append((t0, t1))
                elif name == "gates":
                    iv["gates"].append((t0, t1))
        if lo is None or hi <= lo:
            continue
        for k, lst in iv.items():
            pts = []
            for t0, t1 in lst:
                t0, t1 = max(t0, lo), min(t1, hi)
                if t1 > t0:
                    pts += [(t0, 1), (t1, -1)]
            pts.sort()
            cur, last = 0, lo
            for t, dlt in pts:
                conc[k][cur] += t - last
                last = t
                cur += dlt
            conc[k][cur] += hi - last
    for k, v in conc.items():
        tot = v.sum()
        print(f"{k:6s}: fraction of time with 0/1/2/3 CTAs in the phase: " + " ".join(f"{x / tot:.2f}" for x in v[:4]))
    # timeline of one SM
    d = tr[sm_show]
    t0 = min(e[0][0] for e in d.values() if e)
    print(f"SM {sm_show}: CTAs {sorted(set(c for c, w in d))}")
    for key in sorted(d):
        if key[1] != 0:
            continue
        line = " ".join(f"{name[:2]}{t1 - t0_:d}" for name, t0_, t1 in phases(d[key]))
        print(f"  cta {key[0]} w{key[1]} start {d[key][0][0] - t0}: {line}")
