"""Unit tests of the schedule compiler's stages (csrc/xyk_sched.cpp, PlanCompiler), host only: the tile search, the round
search and the layout assignment are checked on their own intermediate results (xyk_plan_stages), independently of the
emitted pass descriptors that tests/test_schedule_cpu.py executes."""
import numpy as np
import pytest

from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import plan_stages


def _problems():
    out = []
    for n, order in ((13, 1), (14, 2), (16, 1), (16, 2), (20, 1), (24, 1), (30, 1)):
        K, w = oc.random_problem(n, 4000 + n)
        out.append((f"all-to-all n={n} order={order}", K, w, order))
    K, w = oc.ring_K(18), oc.random_problem(18, 1)[1]
    out.append(("ring n=18", K, w, 1))
    rng = np.random.default_rng(3)
    K, w = oc.random_problem(22, 2)
    K = K * (rng.uniform(size=K.shape) < 0.3)
    K = np.triu(K, 1) + np.triu(K, 1).T
    out.append(("sparse n=22", K, w, 2))
    return out


PROBLEMS = _problems()


@pytest.mark.parametrize("name,K,w,order", PROBLEMS, ids=[p[0] for p in PROBLEMS])
def test_stage1_tiles_cover_the_gate_list_in_dependency_order(name, K, w, order):
    d = plan_stages(K, w, 0.05, order, 1, upto_stage=1)
    TB, n = d["TB"], d["nq"]
    passes = d["passes"]
    where = {}   # (segment, gate index) -> pass
    for p, ps in enumerate(passes):
        tile = set(ps["tile"])
        assert 2 <= len(tile) <= min(TB, n)   # a tile may stay below TB qubits (the layout fills the low bits up)
        assert ps["gates"], "a pass without gates"
        for idx, i, j, _sweep in ps["gates"]:
            assert i in tile and j in tile
            assert (ps["seg"], idx) not in where, "gate scheduled twice"
            where[(ps["seg"], idx)] = p
        if p:
            assert ps["seg"] >= passes[p - 1]["seg"]
            assert ps["has_phase"] == (ps["seg"] != passes[p - 1]["seg"])   # the phase rides on a segment's first pass
    for s, cnt in enumerate(d["segments"]):
        assert sorted(idx for (seg, idx) in where if seg == s) == list(range(cnt)), "a gate is missing"
    # two gates that share a qubit keep their order: the earlier one is never in a later pass, and inside a pass the
    # gate list is ascending in the sequence index
    for ps in passes:
        idxs = [g[0] for g in ps["gates"]]
        assert idxs == sorted(idxs)
    by_seg = {}
    for p, ps in enumerate(passes):
        for idx, i, j, _ in ps["gates"]:
            by_seg.setdefault(ps["seg"], []).append((idx, i, j, p))
    for seg, gl in by_seg.items():
        gl.sort()
        last_pass_of_qubit = {}
        for idx, i, j, p in gl:
            for q in (i, j):
                assert last_pass_of_qubit.get(q, -1) <= p, f"gate {idx} on qubit {q} overtakes an earlier gate"
                last_pass_of_qubit[q] = p
    # consecutive tiles (the program is cyclic) share enough qubits for contiguous output chunks
    if len(passes) > 1 and n > TB:
        for p in range(len(passes)):
            a, b = set(passes[p]["tile"]), set(passes[(p + 1) % len(passes)]["tile"])
            assert len(a & b) >= d["need_shared"]


@pytest.mark.parametrize("name,K,w,order", PROBLEMS, ids=[p[0] for p in PROBLEMS])
def test_stage2_rounds_partition_every_pass(name, K, w, order):
    d = plan_stages(K, w, 0.05, order, 1, upto_stage=2)
    RB, RS = d["RB"], d["RS"]
    for ps in d["passes"]:
        gate = {g[0]: g for g in ps["gates"]}
        seen = []
        round_of = {}
        for r, rd in enumerate(ps["rounds"]):
            reg = set(rd["reg"])
            assert rd["gates"], "an empty round"
            if rd["split"]:
                left = set(rd["left"])
                assert len(reg) == RS and left < reg and 1 <= len(left) <= RS // 2
            else:
                assert len(reg) <= RB
            pairs, sweeps = set(), set()
            for idx in rd["gates"]:
                _, i, j, sweep = gate[idx]
                assert i in reg and j in reg
                if rd["split"]:
                    assert (i in left) != (j in left), "a split round only couples the left with the right group"
                assert (i, j) not in pairs, "a pair twice in one round"
                pairs.add((i, j))
                sweeps.add(sweep)
                round_of[idx] = r
            assert len(sweeps) == 1, "a round mixes two sweeps"
            seen += rd["gates"]
        assert sorted(seen) == sorted(gate), "the rounds do not partition the pass"
        last_round_of_qubit = {}
        for idx in sorted(gate):
            _, i, j, _ = gate[idx]
            for q in (i, j):
                assert last_round_of_qubit.get(q, -1) <= round_of[idx]
                last_round_of_qubit[q] = round_of[idx]


@pytest.mark.parametrize("name,K,w,order", PROBLEMS, ids=[p[0] for p in PROBLEMS])
def test_stage3_layouts_put_the_tile_on_the_low_bits(name, K, w, order):
    d = plan_stages(K, w, 0.05, order, 1, upto_stage=3)
    n, TB = d["nq"], d["TB"]
    for p, ps in enumerate(d["passes"]):
        lay = ps["layout"]
        assert sorted(lay) == list(range(n)), "a layout is a permutation of the physical bits"
        assert set(ps["tile"]) <= {q for q in range(n) if lay[q] < TB}
        if p and n > TB:
            assert ps["n_shared_in"] >= d["need_shared"]
    assert d["final_layout"] == d["passes"][0]["layout"], "a cyclic program ends in its start layout"


def test_stage_results_do_not_depend_on_the_step_size():
    """tile sequences and round decompositions are functions of the gate PATTERN (they are memoised by it)"""
    K, w = oc.random_problem(18, 7)
    a = plan_stages(K, w, 0.05, 2, 1, upto_stage=3)
    b = plan_stages(K, w, 0.31, 2, 1, upto_stage=3)
    assert a == b
