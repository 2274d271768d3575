"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the
oracle on identical seeded inputs.  Tolerances (complex128): state fidelity >= 1 - 1e-10 and
|dR| <= 1e-9, as BASELINE.json's north_star states."""

import ctypes
import json
import os

import numpy as np
import pytest

from oracle import xy_oracle as oc
from scpn_quantum_control_b200 import QuantumKuramotoSolver, QuantumUPDESolver, auto_solve
from scpn_quantum_control_b200.engine import XYKEngine, batch_run

pytestmark = pytest.mark.gpu

FID_TOL = 1e-10
R_TOL = 1e-9
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))["cases"]
REFGOLD = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))

_ORACLE_SO = os.path.join(os.path.dirname(HERE), "oracle", "_build", "libxyoracle.so")


def c_oracle_evolve(n, K, omega, delta, reps, order):
    """Final state of the C oracle after one interval (fast enough up to n ~ 24)."""
    lib = ctypes.CDLL(_ORACLE_SO)
    dp = ctypes.POINTER(ctypes.c_double)
    Kc, wc = oc.canonicalise(n, K, omega)
    Ks = np.ascontiguousarray((Kc + Kc.T) / 2.0)
    psi = np.zeros(2 << n)
    lib.xyo_init_state(n, wc.ctypes.data_as(dp), psi.ctypes.data_as(dp))
    lib.xyo_evolve.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, dp]
    lib.xyo_evolve(n, Ks.ctypes.data_as(dp), wc.ctypes.data_as(dp), float(delta), reps, order, psi.ctypes.data_as(dp))
    return psi.view(np.complex128)


def engine_evolve(n, K, omega, delta, reps, order, engine):
    with XYKEngine(n) as eng:
        eng.set_problem(*oc.canonicalise(n, K, omega))
        eng.set_engine(engine)
        eng.init_product_state()
        eng.apply_layers(delta, order, reps)
        R, ang = eng.order_parameter()
        psi = eng.get_state()
        st = eng.stats()
    return psi, R, st


@pytest.mark.parametrize("n", [2, 3, 4, 6, 8, 10])
@pytest.mark.parametrize("order", [1, 2])
def test_gate_by_gate_engine_vs_oracle(n, order):
    K, w = oc.random_problem(n, 40 + n)
    Kc, wc = oc.canonicalise(n, K, w)
    z, p = oc.term_list(Kc, wc)
    ref = oc.evolve_product_formula(oc.initial_state(wc), z, p, 0.1, 3, order)
    psi, R, _ = engine_evolve(n, K, w, 0.1, 3, order, "simple")
    assert 1.0 - oc.fidelity(psi, ref) < FID_TOL
    assert abs(R - oc.order_parameter(ref, n)[0]) < R_TOL


@pytest.mark.parametrize("n", [12, 13, 14, 16, 18, 20])
@pytest.mark.parametrize("order", [1, 2])
def test_tiled_engine_vs_oracle(n, order):
    K, w = oc.random_problem(n, 70 + n)
    ref = c_oracle_evolve(n, K, w, 0.1, 2, order)
    psi, R, st = engine_evolve(n, K, w, 0.1, 2, order, "tiled")
    assert st["pass_launches"] > 0
    assert 1.0 - oc.fidelity(psi, ref) < FID_TOL
    assert np.abs(psi - ref).max() < 1e-12
    assert abs(R - oc.order_parameter(ref, n)[0]) < R_TOL


def test_tiled_engine_order4_and_6():
    n = 13
    K, w = oc.random_problem(n, 11)
    Kc, wc = oc.canonicalise(n, K, w)
    z, p = oc.term_list(Kc, wc)
    for order in (4, 6):
        ref = oc.evolve_product_formula(oc.initial_state(wc), z, p, 0.1, 1, order)
        psi, R, _ = engine_evolve(n, K, w, 0.1, 1, order, "tiled")
        assert 1.0 - oc.fidelity(psi, ref) < FID_TOL


def test_tiled_sparse_ring_and_large_angles():
    n = 14
    K = 30.0 * oc.ring_K(n)  # angles 2*K*tau of order several radians: standard-rotation kernel
    w = np.tile(oc.OMEGA_N_16, 2)[:n]
    ref = c_oracle_evolve(n, K, w, 0.2, 2, 1)
    psi, R, _ = engine_evolve(n, K, w, 0.2, 2, 1, "tiled")
    assert 1.0 - oc.fidelity(psi, ref) < FID_TOL


def test_tiled_equals_gate_by_gate_n24():
    """Size-independent property at a size the numpy oracle cannot reach quickly."""
    n = 24
    K, w = oc.random_problem(n, 24)
    with XYKEngine(n) as a, XYKEngine(n) as b:
        for e, name in ((a, "tiled"), (b, "simple")):
            e.set_problem(K, w)
            e.set_engine(name)
            e.init_product_state()
            e.apply_layers(0.1, 1, 1)
        xa, ya = a.xy_expectations()
        xb, yb = b.xy_expectations()
        za, zb = a.z_expectations(), b.z_expectations()
        assert np.abs(xa - xb).max() < 1e-10 and np.abs(ya - yb).max() < 1e-10
        assert np.abs(za - zb).max() < 1e-10
        assert abs(a.norm2() - 1.0) < 1e-10
        # <sum Z> is conserved by every gate of the XY model
        w0 = np.mod(w, 2 * np.pi)
        assert abs(za.sum() - np.cos(w0).sum()) < 1e-9
        pa, pb = a.get_state(), b.get_state()
        assert 1.0 - oc.fidelity(pa, pb) < FID_TOL


def test_full_size_n30_properties():
    """BASELINE configs[3] at full size (16 GiB state), where no oracle reaches: norm and <sum Z> are conserved by
    a first-order layer, and an order-2 step followed by the same step with negative time (its exact inverse, gate
    by gate) returns the analytic product state: <Z_q> = cos(theta_q), <X_q> = sin(theta_q), <Y_q> = 0."""
    from scpn_quantum_control_b200 import _lib

    n = 30
    lib = _lib.load()
    import ctypes

    free_b, total_b = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.xyk_device_memory(0, ctypes.byref(free_b), ctypes.byref(total_b)) == 0
    if free_b.value < 40 * 2**30:
        pytest.skip("needs 2 x 16 GiB of free device memory")
    rng = np.random.default_rng(20260730)
    A = rng.uniform(0.05, 1.0, (n, n))
    K = (A + A.T) / 2.0
    np.fill_diagonal(K, 0.0)
    w = rng.uniform(0.5, 4.0, n)
    th = np.mod(w, 2 * np.pi)
    with XYKEngine(n) as e:
        e.set_problem(K, w)
        e.set_engine("tiled")
        e.init_product_state()
        e.apply_layers(0.02, 1, 1)
        st = e.stats()
        assert st["last_gates"] == n * (n - 1) // 2 and st["last_passes"] >= 2
        assert abs(e.norm2() - 1.0) < 1e-10
        assert abs(e.z_expectations().sum() - np.cos(th).sum()) < 1e-9
        x1, _ = e.xy_expectations()
        assert np.abs(x1 - np.sin(th)).max() > 1e-3  # the layer did something
        e.init_product_state()
        e.apply_layers(0.05, 2, 1)
        e.apply_layers(-0.05, 2, 1)
        x, y = e.xy_expectations()
        z = e.z_expectations()
        assert np.abs(x - np.sin(th)).max() < 1e-9
        assert np.abs(y).max() < 1e-9
        assert np.abs(z - np.cos(th)).max() < 1e-9
        assert abs(e.norm2() - 1.0) < 1e-10


@pytest.mark.parametrize("name", sorted(k for k, v in GOLD.items() if v["mode"] == "T"))
def test_solver_run_vs_golden_trotter_traces(name):
    c = GOLD[name]
    s = QuantumKuramotoSolver(c["n"], np.array(c["K"]), np.array(c["omega"]), trotter_order=c["order"],
                              evolution="trotter")
    res = s.run(c["t_max"], c["dt"], c["trotter_per_step"])
    np.testing.assert_allclose(res["times"], c["times"], rtol=0, atol=1e-15)
    assert np.abs(res["R"] - np.array(c["R"])).max() < R_TOL
    assert res.metadata["backend"] == "gpu_statevector"
    if "state_re" in c:
        eng = XYKEngine(c["n"])
        eng.set_problem(s.K, s.omega)
        eng.init_product_state()
        _, steps = oc.time_grid(c["t_max"], c["dt"])
        for d in steps:
            eng.apply_layers(float(d), c["order"], c["trotter_per_step"])
        psi = eng.get_state()
        eng.close()
        ref = np.array(c["state_re"]) + 1j * np.array(c["state_im"])
        assert 1.0 - oc.fidelity(psi, ref) < FID_TOL
    s.close()


@pytest.mark.parametrize("name", sorted(k for k, v in GOLD.items() if v["mode"] == "E"))
def test_solver_reference_mode_vs_exact_goldens(name):
    """Default numerics (evolution='reference') must reproduce exact per-step propagation."""
    c = GOLD[name]
    s = QuantumKuramotoSolver(c["n"], np.array(c["K"]), np.array(c["omega"]))
    res = s.run(c["t_max"], c["dt"])
    assert np.abs(res["R"] - np.array(c["R"])).max() < R_TOL
    s.close()


def test_reference_goldens_mode_e():
    """The reference's own golden values (tests/test_classical_reference_regression.py:35-46 and
    the stored outputs of its run())."""
    from scpn_quantum_control_b200 import OMEGA_N_16, build_knm_paper27

    for n_str, expected in REFGOLD["test_classical_reference_regression"]["R_dt01"].items():
        n = int(n_str)
        s = QuantumKuramotoSolver(n, build_knm_paper27(n), OMEGA_N_16[:n])
        res = s.run(0.1, 0.1)
        assert abs(res["R"][-1] - expected) < R_TOL, (n, res["R"][-1], expected)
        s.close()
    g = REFGOLD["reproducible_comparison_n8"]
    s = QuantumKuramotoSolver(8, build_knm_paper27(8), OMEGA_N_16[:8])
    res = s.run(g["t_max"], g["dt"], g["trotter_per_step"])
    assert abs(res["R"][-1] - g["quantum_trotter_r_final"]) < R_TOL
    s.close()
    g = REFGOLD["qutip_comparison_qiskit_unitary"]
    idx = np.arange(4)
    K = 0.45 * np.exp(-0.3 * np.abs(idx[:, None] - idx[None, :]))
    s = QuantumKuramotoSolver(4, K, np.linspace(0.8, 1.2, 4))
    res = s.run(g["t_max"], g["dt"])
    assert np.abs(res["R"] - np.array(g["R"])).max() < R_TOL
    s.close()


def test_n16_paper27_exact_trajectory():
    """results/classical_16q_reference.json: evo_16q_8step_dt0.05 (exact evolution, n=16)."""
    from scpn_quantum_control_b200 import OMEGA_N_16, build_knm_paper27

    g = REFGOLD["classical_16q_reference"]["evo_16q_8step_dt0.05"]
    s = QuantumKuramotoSolver(16, build_knm_paper27(16), OMEGA_N_16)
    res = s.run(0.4, 0.05)
    assert np.abs(res["R"] - np.array(g)).max() < R_TOL
    s.close()


def test_batch_kernel_vs_oracle():
    rng = np.random.default_rng(5)
    for n, order in ((4, 1), (8, 2), (12, 1)):
        B = 6
        Ks, ws = [], []
        for b in range(B):
            K, w = oc.random_problem(n, 1000 + b)
            Ks.append(K)
            ws.append(w)
        _, steps = oc.time_grid(0.25, 0.1)
        R, states = batch_run(np.array(Ks), np.array(ws), steps, 3, order, return_states=True)
        for b in range(B):
            _, Rref, psi = oc.run(n, Ks[b], ws[b], 0.25, 0.1, 3, order=order, return_state=True)
            assert np.abs(R[b] - Rref).max() < R_TOL
            assert 1.0 - oc.fidelity(states[b], psi) < FID_TOL


def test_observables_vs_oracle():
    n = 10
    K, w = oc.random_problem(n, 77)
    Kc, wc = oc.canonicalise(n, K, w)
    z, p = oc.term_list(Kc, wc)
    ref = oc.evolve_product_formula(oc.initial_state(wc), z, p, 0.3, 4, 2)
    with XYKEngine(n) as eng:
        eng.set_problem(Kc, wc)
        eng.set_state(ref)
        ex, ey = eng.xy_expectations()
        rx, ry = oc.xy_expectations(ref, n)
        assert np.abs(ex - rx).max() < 1e-12 and np.abs(ey - ry).max() < 1e-12
        assert np.abs(eng.z_expectations() - oc.z_expectations(ref, n)).max() < 1e-12
        assert np.abs(eng.correlation_matrix_xy() - oc.correlation_matrix_xy(ref, n)).max() < 1e-12
        assert abs(eng.energy() - oc.energy(ref, n, Kc, wc)) < 1e-10


def test_edge_cases():
    # n = 1: no pair terms, pure phase; zero time; zero coupling; zero omega entries
    s = QuantumKuramotoSolver(1, np.zeros((1, 1)), np.array([0.7]), evolution="trotter")
    res = s.run(0.2, 0.1)
    assert np.allclose(res["R"], abs(np.sin(0.7)), atol=1e-12)
    s.close()
    s = QuantumKuramotoSolver(2, np.zeros((2, 2)), np.array([0.2, 0.4]), evolution="trotter")
    res = s.run(0.35, 0.1, trotter_per_step=1)
    np.testing.assert_allclose(res["times"], [0.0, 0.1, 0.2, 0.3, 0.35])
    s.close()
    s = QuantumKuramotoSolver(3, oc.ring_K(3), np.array([0.0, 1.0, 0.0]), evolution="trotter")
    res = s.run(0.0, 0.1)
    assert res["times"].shape == (1,)
    _, Rref = oc.run(3, oc.ring_K(3), np.array([0.0, 1.0, 0.0]), 0.0, 0.1, 5, order=1)
    assert abs(res["R"][0] - Rref[0]) < 1e-12
    s.close()


def test_auto_solve_and_upde():
    K = oc.paper27_K(4)
    w = oc.OMEGA_N_16[:4]
    out = auto_solve(K, w, t_max=0.3, dt=0.1, evolution="trotter")
    assert out["backend_used"] == "gpu_statevector"
    _, Rref = oc.run(4, K, w, 0.3, 0.1, 5, order=1)
    assert np.abs(out["result"]["R"] - Rref).max() < R_TOL
    u = QuantumUPDESolver(K, w, evolution="trotter")
    vals = [u.step(0.05, 5)["R_global"] for _ in range(5)]
    _, Rref = oc.run(4, K, w, 0.25, 0.05, 5, order=1)
    assert np.abs(np.array(vals) - Rref[1:]).max() < R_TOL
    doc = REFGOLD["doc_traces_3_digits"]["upde_run_5steps_dt0.05_n4"]
    assert np.abs(np.round(np.concatenate([[Rref[0]], vals]), 3) - np.array(doc)).max() < 1.5e-3
    u.close()


@pytest.mark.parametrize("n", [22, 24])
def test_reference_mode_large_n_repetition_count(n):
    """evolution="reference" above n = 20: R(t) of the calibrated 6th-order composition equals the
    same run with doubled repetitions to 1e-9 -- for the direct halving check on the full problem and for the
    reduced-problem proxy that larger / sharded states use."""
    K, w = oc.random_problem(n, 4200 + n)
    dt, nsteps = 0.1, 2
    traces = {}
    for kind in ("direct", "proxy"):
        s = QuantumKuramotoSolver(n, K, w)  # default evolution="reference"
        if kind == "proxy":
            s._calibrate_direct_max_n = 0
        res = s.run(nsteps * dt, dt)
        meta = res.metadata["exact_composition"]
        assert meta["calibration"]["kind"] == ("direct" if kind == "direct" else "reduced_problem")
        reps = max(meta["reps"].values())
        traces[kind] = np.asarray(res["R"])
        # the same trajectory with twice the repetitions, straight on the engine
        eng = s.engine()
        eng.init_product_state()
        R2 = [eng.order_parameter()[0]]
        for _ in range(nsteps):
            eng.apply_layers(dt, 6, 2 * reps)
            R2.append(eng.order_parameter()[0])
        s.close()
        assert np.abs(traces[kind] - np.array(R2)).max() < R_TOL, (kind, reps)
    assert np.abs(traces["direct"] - traces["proxy"]).max() < R_TOL


def test_set_coupling_mid_trajectory_vs_oracle():
    """re-parameterising K on the LIVE state (realtime-feedback tick, reference control/realtime_feedback.py:98-137):
    two steps with K, then K scaled by 0.5 for two more steps, against the oracle applying the same sequence"""
    n = 14
    K, w = oc.random_problem(n, 515)
    Kc, wc = oc.canonicalise(n, K, w)
    u = QuantumUPDESolver(K, w, evolution="trotter")
    got = [u.step(0.1, 3)["R_global"] for _ in range(2)]
    u.set_coupling(0.5 * K)
    got += [u.step(0.1, 3)["R_global"] for _ in range(2)]
    psi_gpu = u.statevector()
    u.close()
    psi = oc.initial_state(wc)
    want = []
    for k in range(4):
        z, p = oc.term_list(Kc if k < 2 else 0.5 * Kc, wc)
        psi = oc.evolve_product_formula(psi, z, p, 0.1, 3, 1)
        want.append(oc.order_parameter(psi, n)[0])
    assert np.abs(np.array(got) - np.array(want)).max() < R_TOL
    assert 1.0 - oc.fidelity(psi_gpu, psi) < FID_TOL


def test_measuring_an_external_state_leaves_the_live_state_alone():
    n = 13
    K, w = oc.random_problem(n, 99)
    s = QuantumKuramotoSolver(n, K, w, evolution="trotter")
    s.prepare_initial_state()
    s.evolve(0.1, 2)
    R_live, _ = s.measure_order_parameter()
    other = oc.initial_state(oc.canonicalise(n, K, w)[1])
    R_ext, _ = s.measure_order_parameter(other)
    assert abs(R_ext - oc.order_parameter(other, n)[0]) < 1e-12
    assert abs(s.energy_expectation(other) - oc.energy(other, n, *oc.canonicalise(n, K, w))) < 1e-10
    assert s.measure_order_parameter()[0] == R_live
    s.close()


def test_batch_config3_subset_vs_oracle():
    """BASELINE configs[2]: 4096 independent n = 12 problems, 100 layers each; parity on a random 64-problem subset
     against the C oracle"""
    B, n = 4096, 12
    Ks, ws = np.empty((B, n, n)), np.empty((B, n))
    for b in range(B):
        Ks[b], ws[b] = oc.random_problem(n, 20260716 + b)
    steps = np.full(20, 0.1)
    R = batch_run(Ks, ws, steps, 5, 1)
    assert R.shape == (B, 21)
    lib = ctypes.CDLL(_ORACLE_SO)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.xyo_run.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp]
    for b in np.random.default_rng(3).choice(B, 64, replace=False):
        Kc, wc = oc.canonicalise(n, Ks[b], ws[b])
        Ksym = np.ascontiguousarray((Kc + Kc.T) / 2.0)
        psi, Rref = np.zeros(2 << n), np.zeros(21)
        lib.xyo_run(n, Ksym.ctypes.data_as(dp), wc.ctypes.data_as(dp), steps.ctypes.data_as(dp), 20, 5, 1,
                    psi.ctypes.data_as(dp), Rref.ctypes.data_as(dp))
        assert np.abs(R[b] - Rref).max() < R_TOL


@pytest.mark.parametrize("n", [13, 16, 21])
def test_tile_observables_after_evolution(n):
    """<X_q>, <Y_q> through the shared-memory tile sweeps (n_local >= 13) in the permuted layout the passes leave behind"""
    K, w = oc.random_problem(n, 300 + n)
    ref = c_oracle_evolve(n, K, w, 0.2, 2, 1)
    with XYKEngine(n) as eng:
        eng.set_problem(*oc.canonicalise(n, K, w))
        eng.set_engine("tiled")
        eng.init_product_state()
        eng.apply_layers(0.2, 1, 2)
        ex, ey = eng.xy_expectations()
        ez = eng.z_expectations()
    rx, ry = oc.xy_expectations(ref, n)
    assert np.abs(ex - rx).max() < 1e-12 and np.abs(ey - ry).max() < 1e-12
    assert np.abs(ez - oc.z_expectations(ref, n)).max() < 1e-12


@pytest.mark.parametrize("case", range(12))
def test_random_coupling_patterns_through_the_tiled_engine(case):
    """seeded random sparsity patterns (empty rows, isolated qubits, missing Z terms, negative steps) through the fused tile
    passes vs the oracle: partial rounds, unbalanced split rounds and masks that the all-to-all problems never produce"""
    rng = np.random.default_rng(7100 + case)
    n = int(rng.integers(13, 17))
    density = [0.05, 0.2, 0.5, 0.8][case % 4]
    A = rng.uniform(0.0, 2.0, (n, n)) * (rng.uniform(size=(n, n)) < density)
    K = np.triu(A, 1)
    K = K + K.T
    if case % 3 == 0:
        q = int(rng.integers(0, n))
        K[q, :] = 0.0
        K[:, q] = 0.0
    w = rng.uniform(0.1, 5.0, n)
    if case % 2:
        w[int(rng.integers(0, n))] = 0.0
    order = 1 + case % 2
    delta = [0.1, -0.07, 0.03][case % 3]
    ref = c_oracle_evolve(n, K, w, delta, 2, order)
    psi, R, st = engine_evolve(n, K, w, delta, 2, order, "tiled")
    assert 1.0 - oc.fidelity(psi, ref) < FID_TOL
    assert np.abs(psi - ref).max() < 1e-10
    assert abs(R - oc.order_parameter(ref, n)[0]) < R_TOL
