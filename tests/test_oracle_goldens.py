"""Pin the oracle against every golden value the reference holds for this path
(tests/golden/reference_goldens.json, extracted by tests/golden/make_goldens.py)."""

import json
import os

import numpy as np
import pytest

from oracle import xy_oracle as oc

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))
VEC = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))["cases"]


@pytest.mark.parametrize("n", [2, 4, 6, 8, 10, 12])
def test_mode_e_R_dt01(n):
    """tests/test_classical_reference_regression.py:35-46 (abs 1e-10 there; we hold 1e-13)."""
    expected = REF["test_classical_reference_regression"]["R_dt01"][str(n)]
    assert expected == REF["classical_16q_reference"]["evo_R_dt0.1"][str(n)]
    _, R = oc.run(n, oc.paper27_K(n), oc.OMEGA_N_16[:n], 0.1, 0.1, mode="E")
    assert abs(R[-1] - expected) < 1e-13


def test_mode_e_14q():
    expected = REF["classical_16q_reference"]["evo_R_dt0.1"]["14"]
    _, R = oc.run(14, oc.paper27_K(14), oc.OMEGA_N_16[:14], 0.1, 0.1, mode="E")
    assert abs(R[-1] - expected) < 1e-12


def test_n16_initial_R():
    g = REF["classical_16q_reference"]["evo_16q_8step_dt0.05"]
    psi = oc.initial_state(oc.OMEGA_N_16)
    assert abs(oc.order_parameter(psi, 16)[0] - g[0]) < 1e-14


def test_stored_run_output_n8():
    """data/classical_quantum_comparison/reproducible_comparison_n8.json: the real run() output."""
    g = REF["reproducible_comparison_n8"]
    _, R = oc.run(8, oc.paper27_K(8), oc.OMEGA_N_16[:8], g["t_max"], g["dt"], mode="E")
    assert abs(R[-1] - g["quantum_trotter_r_final"]) < 1e-13
    assert abs(R[-1] - g["classical_exact_r_final"]) < 1e-13
    # the literal first-order product formula does NOT reproduce it
    _, Rt = oc.run(8, oc.paper27_K(8), oc.OMEGA_N_16[:8], g["t_max"], g["dt"], g["trotter_per_step"], order=1)
    assert abs(Rt[-1] - g["quantum_trotter_r_final"]) > 1e-4


def test_stored_run_output_n4_41_points():
    g = REF["qutip_comparison_qiskit_unitary"]
    idx = np.arange(4)
    K = 0.45 * np.exp(-0.3 * np.abs(idx[:, None] - idx[None, :]))
    _, R = oc.run(4, K, np.linspace(0.8, 1.2, 4), g["t_max"], g["dt"], mode="E")
    assert R.shape[0] == 41
    assert np.abs(R - np.array(g["R"])).max() < 5e-15


def test_mode_t_doc_traces():
    """docs/pipeline_performance.md:376,410 -- 3 printed digits of the order-1 x 5 product formula."""
    d = REF["doc_traces_3_digits"]
    K, w = oc.paper27_K(4), oc.OMEGA_N_16[:4]
    _, R = oc.run(4, K, w, 0.3, 0.1, 5, order=1)
    assert np.all(np.round(R[:3], 3) == np.array(d["run_t0.3_dt0.1_n4"]))
    _, R = oc.run(4, K, w, 0.25, 0.05, 5, order=1)
    assert np.all(np.round(R, 3) == np.array(d["upde_run_5steps_dt0.05_n4"]))
    # exact evolution prints different third digits: the doc traces are Mode T
    _, Re = oc.run(4, K, w, 0.3, 0.1, mode="E")
    assert not np.all(np.round(Re[:3], 3) == np.array(d["run_t0.3_dt0.1_n4"]))


def test_time_grid_example():
    g = REF["time_grid_example"]
    times, steps = oc.time_grid(g["t_max"], g["dt"])
    np.testing.assert_allclose(times, g["times"], atol=1e-15)
    np.testing.assert_allclose(steps, g["step_sizes"], atol=1e-15)
    times, steps = oc.time_grid(0.0, 0.1)
    assert times.tolist() == [0.0] and steps.size == 0


def test_order2_converges_faster_and_suzuki_reaches_exact():
    n = 6
    K, w = oc.paper27_K(n), oc.OMEGA_N_16[:n]
    _, Re, pe = oc.run(n, K, w, 0.5, 0.1, mode="E", return_state=True)
    errs = {}
    for order, reps in ((1, 5), (2, 5), (4, 2), (6, 1)):
        _, R, p = oc.run(n, K, w, 0.5, 0.1, reps, order=order, return_state=True)
        errs[order] = 1.0 - oc.fidelity(p, pe)
    assert errs[1] > errs[2] > errs[4] > 0 or errs[4] < 1e-11
    assert errs[6] < 1e-12


@pytest.mark.parametrize("name", ["n4_paper27_o1", "n4_ring_o2", "n8_paper27_o2", "n4_paper27_exact"])
def test_committed_oracle_vectors_are_reproducible(name):
    c = VEC[name]
    _, R = oc.run(c["n"], np.array(c["K"]), np.array(c["omega"]), c["t_max"], c["dt"], c["trotter_per_step"],
                  order=c["order"], mode=c["mode"])
    assert np.abs(R - np.array(c["R"])).max() < 1e-13


def test_observable_cross_checks():
    """sectors.rs:67-99 full-loop form and pauli.rs half-loop form agree; unitarity; <sum Z> conserved."""
    n = 7
    K, w = oc.random_problem(n, 9)
    Kc, wc = oc.canonicalise(n, K, w)
    z, p = oc.term_list(Kc, wc)
    psi0 = oc.initial_state(wc)
    psi = oc.evolve_product_formula(psi0.copy(), z, p, 0.4, 3, 2)
    assert abs(np.vdot(psi, psi).real - 1.0) < 1e-13
    idx = np.arange(1 << n)
    zr = 0.0 + 0.0j
    for q in range(n):
        zr += np.vdot(psi, psi[idx ^ (1 << q)] * 1.0).real + 1j * np.sum(
            (psi.conj() * (-1j) * (1 - 2 * ((idx >> q) & 1)) * psi[idx ^ (1 << q)])).real
    assert abs(abs(zr) / n - oc.order_parameter(psi, n)[0]) < 1e-13
    assert abs(oc.z_expectations(psi, n).sum() - oc.z_expectations(psi0, n).sum()) < 1e-12
    # energy of |0...0> is -sum omega (tests/test_knm_hamiltonian.py:71-84)
    e0 = np.zeros(1 << n, dtype=complex)
    e0[0] = 1.0
    assert abs(oc.energy(e0, n, Kc, wc) + wc.sum()) < 1e-13
    H = oc.hamiltonian_sparse(Kc, wc)
    assert abs(np.vdot(psi, H @ psi).real - oc.energy(psi, n, Kc, wc)) < 1e-12
