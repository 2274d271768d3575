"""NumPy emulation of a compiled pass program (test helper): applies the passes exactly the way
the tile kernel does -- gates on physical bit positions, then the bit permutation of the store --
so that schedule, layouts and gate order can be checked on the CPU against the oracle."""

from __future__ import annotations

import numpy as np

from oracle import xy_oracle as oc


def permute_bits(psi: np.ndarray, n: int, dst_of_src) -> np.ndarray:
    """out[index with bit dst_of_src[b] = bit b of src index] = psi[src index]."""
    view = psi.reshape((2,) * n)  # axis a <-> bit n-1-a
    # new axis for destination bit d must come from source bit b with dst_of_src[b] = d
    src_of_dst = [0] * n
    for b, d in enumerate(dst_of_src):
        src_of_dst[d] = b
    axes = [n - 1 - src_of_dst[n - 1 - a] for a in range(n)]
    return np.ascontiguousarray(np.transpose(view, axes)).reshape(-1)


def simulate_plan(plan: dict, psi_canonical: np.ndarray, omega: np.ndarray) -> tuple[np.ndarray, list[int]]:
    """Returns (state in the plan's end layout, end perm logical->physical)."""
    n = plan["nq"]
    TB = plan["TB"]
    perm = list(plan["start_perm"])
    psi = permute_bits(psi_canonical.copy(), n, perm)  # source bit q (logical) -> physical perm[q]
    for p in plan["passes"]:
        inv = [0] * n
        for q in range(n):
            inv[perm[q]] = q
        assert [inv[b] for b in range(TB)] == p["tileq"][:TB], "tile qubits are not on the low bits"
        if p["has_phase"]:
            z_terms = [(perm[q], float(omega[q])) for q in range(n) if abs(omega[q]) > 1e-15]
            oc.apply_z_phase(psi, z_terms, p["ztau"])
        for r in p["rounds"]:
            regq, regpos, thrpos = r["regq"], r["regpos"], r["thrpos"]
            assert sorted(regpos + thrpos) == list(range(TB))
            for q, pos in zip(regq, regpos):
                assert perm[q] == pos
            for i, j, phi in r["gates"]:
                assert i in regq and j in regq
                oc.apply_pair(psi, perm[i], perm[j], phi)
        out_of_in = p["out_of_in"]
        psi = permute_bits(psi, n, out_of_in)
        perm = [out_of_in[perm[q]] for q in range(n)]
    return psi, perm
