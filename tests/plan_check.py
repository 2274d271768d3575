"""Structural invariants of a compiled pass program that the NumPy emulation (plan_sim) cannot see because it has
no threads: barrier kinds, warp-local rounds, the phase round, split-round bipartitions, descriptor limits."""

from __future__ import annotations

MAX_ROUNDS = 14  # kMaxRounds of the pass descriptor


def warp_positions(r: dict) -> list[int]:
    """Tile positions on the two warp-index bits (thread bits 5, 6) of a round, in bit order."""
    thr = r["thrpos"]
    return thr[4:6] if r["split"] else thr[5:7]


def check_plan(plan: dict) -> None:
    TB = plan["TB"]
    for pi, p in enumerate(plan["passes"]):
        rounds = p["rounds"]
        assert 1 <= len(rounds) <= MAX_ROUNDS, f"pass {pi}: {len(rounds)} rounds"
        assert rounds[-1]["cta_sync_after"] == 1, f"pass {pi}: the store after the last round needs a block-wide barrier"
        if p["has_phase"]:
            assert not rounds[0]["split"], f"pass {pi}: the fused diagonal phase needs a complex first round"
        unb = {r["nleft"] for r in rounds if r["split"] and r["nleft"] != 3}
        assert len(unb) <= 1, f"pass {pi}: mixes unbalanced split kinds {unb}"
        for ri, r in enumerate(rounds):
            nreg = len(r["regq"])
            assert nreg == (6 if r["split"] else min(5, TB)), f"pass {pi} round {ri}: {nreg} register qubits"
            assert sorted(r["regpos"] + r["thrpos"]) == list(range(TB))
            assert [p["tileq"][pos] for pos in r["regpos"]] == r["regq"]
            if r["split"]:
                assert r["nleft"] in (1, 2, 3)
                left, right = set(r["regq"][: r["nleft"]]), set(r["regq"][r["nleft"]:])
                for i, j, _ in r["gates"]:
                    assert (i in left and j in right) or (i in right and j in left), f"pass {pi} round {ri}: gate ({i},{j}) inside a group"
            else:
                for i, j, _ in r["gates"]:
                    assert i in r["regq"] and j in r["regq"]
            if r["cta_sync_after"] == 0:
                # a warp-level barrier is enough only if both rounds put the SAME qubits on the warp-index bits, in the
                # same order, and neither holds them in registers: then every warp reads back what it wrote itself
                assert ri + 1 < len(rounds)
                nxt = rounds[ri + 1]
                if TB >= 12:
                    assert warp_positions(r) == warp_positions(nxt), f"pass {pi} round {ri}: warp qubits change under a warp barrier"
        if p["fuse_store"]:
            low = sorted(range(TB), key=lambda pos: p["out_of_in"][pos])[:3]
            assert [p["out_of_in"][b] for b in low] == [0, 1, 2]
            assert all(b in rounds[-1]["thrpos"] for b in low), f"pass {pi}: a low output bit sits in a register of the last round"
