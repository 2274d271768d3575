"""Host-only: shared-memory wavefronts per round of the compiled pass program (bank-conflict audit).
Replays the address arithmetic of tile_pass_kernel (pad_slot layout; arrival layout for round 0 of a
fused load) for warp 0..3 and counts 128-byte-window wavefronts per warp instruction vs the ideal."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import plan_describe

def pad_slot(x): return x + (x >> 3) + (x >> 6) + (x >> 9)
def arr_stride(p): return (16 << p) if p < 9 else (16 << p) + (16 << (p - 9))

def wavefronts(addrs, width):
    """addrs: byte address per lane (32), width 8 or 16.  Hardware model: a wavefront serves distinct banks
    (4-byte banks, 32 of them); same 4-byte word may be shared.  128-bit accesses go quarter-warp at a time,
    64-bit half-warp at a time."""
    group = 8 if width == 16 else 16
    total = 0
    for g0 in range(0, 32, group):
        words = {}
        for a in addrs[g0:g0 + group]:
            for w in range(a // 4, (a + width) // 4):
                words.setdefault(w % 32, set()).add(w)
        total += max(len(s) for s in words.values())
    return total

def audit(plan, TB=12, RB=5):
    out = []
    for pi, p in enumerate(plan["passes"]):
        rs = p["rounds"]
        row = []
        for ri, r in enumerate(rs):
            split = r["split"]
            nreg = len(r["regq"])
            thrpos = r["thrpos"]
            first_fused = ri == 0 and p["fuse_load"]
            stride = (lambda q: arr_stride(q)) if first_fused else (lambda q: 16 * pad_slot(1 << q))
            tot = ideal = 0
            for warp in range(4):
                addrs = []
                for lane in range(32):
                    tid = warp * 32 + lane
                    if split:
                        half, tq = tid & 1, tid >> 1
                        b = half * 8
                        for k, pos in enumerate(thrpos):
                            if (tq >> k) & 1: b += stride(pos)
                    else:
                        b = 0
                        for k, pos in enumerate(thrpos):
                            if (tid >> k) & 1: b += stride(pos)
                    addrs.append(b)
                width = 8 if split else 16
                # register index only adds a warp-uniform offset; in split rounds odd-parity elements use the
                # other component (+8 - 16*half): check both parities
                for parity in ((0, 1) if split else (0,)):
                    a2 = [a + (8 - 16 * ((a // 8) & 1)) if parity else a for a in addrs] if split else addrs
                    tot += wavefronts(a2, width); ideal += (2 if split else 4)
            row.append((("S" if split else "C") + str(len(r["gates"])), tot / ideal, r["cta_sync_after"]))
        out.append(row)
    return out

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    K, w = oc.random_problem(n, 20260700 + n)
    plan = plan_describe(K, w, 0.02, order, 1)
    for pi, row in enumerate(audit(plan)):
        p = plan["passes"][pi]
        print(pi, "shared", p["n_shared"], "gates", p["ngates"], " ".join(f"{g}:x{c:.2f}{'B' if s else 'w'}" for g, c, s in row))
