"""CPU tests of the repetition-count rule behind ``evolution="reference"`` (Mode E through the 6th-order composition):
the halving check and the reduced-problem proxy used above _CALIBRATE_DIRECT_MAX_N qubits, validated with the oracle
against the exact propagator (the former norm heuristic missed exact_tol by 1-3 orders of magnitude)."""
import numpy as np
import pytest

from oracle import xy_oracle as oc
from scpn_quantum_control_b200.xy_kuramoto import halving_reps, reduced_problem_for_calibration


def _oracle_observer(n, K, w, step_dt):
    Kc, wc = oc.canonicalise(n, K, w)
    z, p = oc.term_list(Kc, wc)

    def observe(reps):
        psi = oc.evolve_product_formula(oc.initial_state(wc), z, p, step_dt, reps, 6)
        ex, ey = oc.xy_expectations(psi, n)
        return np.concatenate([ex, ey])

    return observe


def test_halving_reps_returns_the_coarser_count():
    seq = {1: 1.0, 2: 0.5, 4: 0.5 + 1e-13, 8: 0.5}
    assert halving_reps(lambda r: np.array([seq[r]]), 1e-10) == 2
    assert halving_reps(lambda r: np.array([float(r)]), 1e-10, max_reps=8) == 8


def test_reduced_problem_keeps_the_local_energy_scale():
    K, w = oc.random_problem(12, 7)
    Ks, ws = reduced_problem_for_calibration(K, w, 8)
    assert Ks.shape == (8, 8) and ws.shape == (8,)
    assert np.allclose(Ks, Ks.T) and np.all(np.diag(Ks) == 0)
    assert np.abs(Ks).sum(axis=1).max() == pytest.approx(np.abs(K).sum(axis=1).max())
    K2, w2 = reduced_problem_for_calibration(K, w, 12)
    assert np.array_equal(K2, K) and np.array_equal(w2, w)


@pytest.mark.parametrize("case", ["ring_strong", "paper27", "random"])
def test_proxy_rule_meets_exact_tol(case):
    """reps from the 8-qubit reduced problem (doubled) reproduce the exact R(t) of the full problem to 1e-9"""
    nsteps, exact_tol = 6, 2e-10
    if case == "ring_strong":
        n, dt = 11, 0.3
        K, w = oc.ring_K(n, 3.0), oc.OMEGA_N_16[:n]
    elif case == "paper27":
        n, dt = 12, 0.1
        K, w = oc.paper27_K(n), oc.OMEGA_N_16[:n]
    else:
        n, dt = 11, 0.1
        K, w = oc.random_problem(n, 20260711)
    Ks, ws = reduced_problem_for_calibration(K, w, 8)
    reps = 2 * halving_reps(_oracle_observer(8, Ks, ws, dt), exact_tol / max(4, nsteps))
    _, R6 = oc.run(n, K, w, nsteps * dt, dt, reps, order=6, mode="T")
    _, RE = oc.run(n, K, w, nsteps * dt, dt, mode="E")
    assert np.max(np.abs(R6 - RE)) < 1e-9, (reps, float(np.max(np.abs(R6 - RE))))
