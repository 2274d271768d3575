"""The schedule compiler (host C++ inside libxyk.so) checked on the CPU against the oracle:
executing the compiled passes in their order, with their layouts, must reproduce the reference
gate order (oracle Mode T) exactly."""

import numpy as np
import pytest

from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import plan_describe
from tests.plan_sim import permute_bits, simulate_plan


def _check(n, K, omega, delta, order, reps):
    plan = plan_describe(K, omega, delta, order, reps)
    K, omega = oc.canonicalise(n, K, omega)
    z_terms, pair_terms = oc.term_list(K, omega)
    psi0 = oc.initial_state(omega)
    ref = oc.evolve_product_formula(psi0.copy(), z_terms, pair_terms, delta, reps, order)
    got, perm = simulate_plan(plan, psi0, omega)
    if plan["trailing_ztau"] != 0.0:
        zt = [(perm[q], float(omega[q])) for q in range(n) if abs(omega[q]) > 1e-15]
        oc.apply_z_phase(got, zt, plan["trailing_ztau"])
    assert perm == plan["start_perm"], "program must end in its own start layout"
    inv = [0] * n
    for q in range(n):
        inv[perm[q]] = q
    back = permute_bits(got, n, inv)
    assert 1.0 - oc.fidelity(back, ref) < 1e-12
    assert np.abs(back - ref).max() < 1e-12
    ngates = sum(len(r["gates"]) for p in plan["passes"] for r in p["rounds"])
    assert ngates == plan["total_gates"]
    return plan


@pytest.mark.parametrize("n", [12, 13, 14, 16])
@pytest.mark.parametrize("order", [1, 2])
def test_all_to_all_plan_matches_reference_order(n, order):
    K, w = oc.random_problem(n, 100 + n)
    plan = _check(n, K, w, 0.1, order, 2)
    npairs = n * (n - 1) // 2
    assert plan["total_gates"] == (npairs if order == 1 else 2 * npairs - 1) * 2


def test_order4_plan():
    K, w = oc.random_problem(13, 5)
    _check(13, K, w, 0.1, 4, 1)


def test_sparse_ring_plan():
    n = 14
    K = oc.ring_K(n)
    w = np.tile(oc.OMEGA_N_16, 2)[:n]
    plan = _check(n, K, w, 0.05, 1, 3)
    assert plan["total_gates"] == 3 * n


def test_no_pairs_plan_is_phase_only():
    n = 13
    K = np.zeros((n, n))
    w = np.linspace(0.5, 2.0, n)
    plan = plan_describe(K, w, 0.1, 2, 2)
    assert plan["passes"] == []
    assert plan["trailing_ztau"] == pytest.approx(0.1)


def test_shared_low_bits_between_consecutive_tiles():
    """Every pass must leave >= 3 of its tile qubits on the lowest output bits (>= 128-byte chunks)."""
    n = 20
    K, w = oc.random_problem(n, 3)
    plan = plan_describe(K, w, 0.1, 1, 1)
    for p in plan["passes"]:
        assert p["n_shared"] >= 3


def test_unbalanced_split_rounds_and_piece_barriers():
    """n=20, order 2: the plan uses unbalanced (2|4 or 1|5) split rounds, a pass longer than the descriptor's round limit is cut
    into pieces, and every pass (piece) ends with a block-wide barrier -- its store reads the whole tile.
    (Regression: the piece boundary used to inherit a warp-level barrier from the unsplit pass.)"""
    n = 20
    K, w = oc.random_problem(n, 70 + n)
    plan = _check(n, K, w, 0.1, 2, 1)
    kinds = {(r["split"], r["nleft"]) for p in plan["passes"] for r in p["rounds"]}
    assert ((1, 2) in kinds or (1, 1) in kinds) and (1, 3) in kinds and (0, 0) in kinds
    assert all(r["nleft"] in (1, 2, 3) for p in plan["passes"] for r in p["rounds"] if r["split"])
    for p in plan["passes"]:
        # a pass never mixes 1|5 and 2|4 rounds (each kind has its own kernel instantiation)
        assert len({r["nleft"] for r in p["rounds"] if r["split"] and r["nleft"] != 3}) <= 1
        assert len(p["rounds"]) <= 14
        assert p["rounds"][-1]["cta_sync_after"] == 1
        for r in p["rounds"]:
            if r["split"]:  # every gate joins the left group with the right group
                left, right = set(r["regq"][: r["nleft"]]), set(r["regq"][r["nleft"]:])
                assert all((i in left) != (j in left) and (i in right) != (j in right) for i, j, _ in r["gates"])


def test_bank_conflict_free_lane_bits():
    """The three lowest lane bits of (almost) every round sit on tile positions of three different bank classes of the
    padded shared-memory layout (position mod 3); n=30 all-to-all: at most one conflicted round per layer."""
    n = 30
    K, w = oc.random_problem(n, 20260730)
    plan = plan_describe(K, w, 0.02, 1, 1)
    bad = 0
    for p in plan["passes"]:
        for ri, r in enumerate(p["rounds"]):
            if ri == 0 and p["fuse_load"]:
                continue  # arrival layout: positions {b, 9 + b}
            if len({pos % 3 for pos in r["thrpos"][:3]}) != 3:
                bad += 1
    assert bad <= 1


def test_plan_invariants_headline_programs():
    """Barrier kinds, warp-local rounds, phase round, split bipartitions, fused-store lanes (tests/plan_check.py)
    of the programs the bench and the GPU tests run."""
    from tests.plan_check import check_plan

    for n, order, reps in ((30, 1, 1), (30, 2, 1), (24, 1, 2), (16, 6, 1), (13, 2, 2), (31, 1, 1)):
        K, w = oc.random_problem(n, 20260700 + n)
        check_plan(plan_describe(K, w, 0.02, order, reps))


def test_random_sparse_couplings_plan_and_invariants():
    """Irregular problems: random coupling graphs of several densities, some omega_i = 0, orders 1 and 2: the
    compiled program reproduces the reference gate order (NumPy emulation) and keeps the structural invariants."""
    from tests.plan_check import check_plan

    rng = np.random.default_rng(12345)
    for _ in range(16):
        n = int(rng.integers(12, 16))
        order, reps = int(rng.choice([1, 2])), int(rng.choice([1, 2]))
        dens = float(rng.choice([0.15, 0.3, 0.6, 1.0]))
        A = rng.uniform(0.05, 1.0, (n, n)) * (rng.random((n, n)) < dens)
        K = (A + A.T) / 2.0
        np.fill_diagonal(K, 0.0)
        w = rng.uniform(0.5, 4.0, n)
        w[rng.random(n) < 0.2] = 0.0
        check_plan(_check(n, K, w, 0.1, order, reps))
