"""NumPy emulation of a SHARDED program (test helper): the whole state is kept as one array in
physical index order (rank bits on top); a local pass acts on every rank's shard at once, a classic
exchange swaps the top local bits with the rank bits, and a peer pass applies the full n-bit output
permutation (its store scatters across ranks)."""

from __future__ import annotations

import numpy as np

from oracle import xy_oracle as oc
from tests.plan_sim import permute_bits


def simulate_sharded(prog: dict, n: int, n_local: int, psi_canonical: np.ndarray, omega: np.ndarray):
    """Returns (state in final physical order, perm logical->physical, number of exchanges, fused exchanges)."""
    g = n - n_local
    perm = list(range(n))  # canonical start: logical q on physical bit q (the g highest on the rank bits)
    psi = psi_canonical.copy()
    n_ex = n_fused = 0

    def relayout(target):
        nonlocal psi, perm
        dst_of_src = list(range(n))
        for q in range(n):
            dst_of_src[perm[q]] = target[q]
        psi = permute_bits(psi, n, dst_of_src)
        perm = list(target)

    for op in prog["ops"]:
        if op["kind"] == "exchange":
            out, inc = op["outgoing"], op["incoming"]
            assert [perm[q] - n_local for q in inc] == list(range(g)), "incoming qubits must sit on the rank bits in order"
            # stage: non-outgoing locals keep their relative order, outgoing go to the top local bits
            locs = sorted((perm[q], q) for q in range(n) if perm[q] < n_local and q not in out)
            target = list(perm)
            pos = 0
            for _, q in locs:
                target[q] = pos
                pos += 1
            for q in out:
                target[q] = pos
                pos += 1
            relayout(target)
            target = list(perm)
            for k in range(g):
                target[out[k]] = n_local + k
                target[inc[k]] = n_local - g + k
            relayout(target)  # the all-to-all is exactly this bit swap
            n_ex += 1
            continue
        plan, loc2log = op["plan"], op["loc2log"]
        assert all(perm[q] < n_local for q in loc2log)
        target = list(perm)
        for l, q in enumerate(loc2log):
            target[q] = plan["start_perm"][l]
        relayout(target)
        TB = plan["TB"]
        passes = plan["passes"]
        for pi, p in enumerate(passes):
            inv = {perm[q]: q for q in range(n)}
            assert [inv[b] for b in range(TB)] == [loc2log[l] for l in p["tileq"][:TB]]
            if p["has_phase"]:
                z_terms = [(perm[q], float(omega[q])) for q in range(n) if abs(omega[q]) > 1e-15]
                oc.apply_z_phase(psi, z_terms, p["ztau"])
            for r in p["rounds"]:
                for i, j, phi in r["gates"]:
                    oc.apply_pair(psi, perm[loc2log[i]], perm[loc2log[j]], phi)
            out_of_in = p["out_of_in"]
            peer = bool(op.get("peer_last")) and pi == len(passes) - 1
            target = list(perm)
            for q in loc2log:
                target[q] = out_of_in[perm[q]]
            if peer:
                for k, q in enumerate(op["incoming"]):
                    assert perm[q] == n_local + k
                    target[q] = op["rank_out"][k]
                for k, q in enumerate(op["outgoing"]):
                    assert target[q] == n_local + k, "outgoing qubit must land on rank bit k"
                n_fused += 1
            else:
                assert all(target[q] < n_local for q in loc2log)
            assert sorted(target) == list(range(n)), "output layout is not a permutation"
            relayout(target)
        if plan["trailing_ztau"] != 0.0:
            z_terms = [(perm[q], float(omega[q])) for q in range(n) if abs(omega[q]) > 1e-15]
            oc.apply_z_phase(psi, z_terms, plan["trailing_ztau"])
    return psi, perm, n_ex, n_fused
