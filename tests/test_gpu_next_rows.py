"""GPU parity tests of the SURVEY 8(f) "next" rows, through the C ABI, against oracle/xy_oracle_next.py.
Tolerances as on the main path: fidelity >= 1 - 1e-10, |dR| <= 1e-9 (complex128)."""
import numpy as np
import pytest

from oracle import xy_oracle as oc
from oracle import xy_oracle_next as on
from scpn_quantum_control_b200.engine import XYKEngine
from scpn_quantum_control_b200.realtime_feedback import RealtimeFeedbackConfig, RealtimeSyncFeedbackController

pytestmark = pytest.mark.gpu

FID_TOL = 1e-10
R_TOL = 1e-9


@pytest.mark.parametrize("n,engine", [(5, "simple"), (13, "tiled"), (16, "tiled")])
def test_apply_ry_on_the_live_state(n, engine):
    """xyk_apply_ry after a tiled evolution (permuted layout) against the oracle's Ry layer"""
    K, w = oc.random_problem(n, 40 + n)
    Kc, wc = oc.canonicalise(n, K, w)
    rng = np.random.default_rng(n)
    ang = rng.uniform(-0.4, 0.4, n)
    ang[1] = 0.0
    with XYKEngine(n) as eng:
        eng.set_problem(Kc, wc)
        eng.set_engine(engine)
        eng.init_product_state()
        eng.apply_layers(0.1, 1, 2)
        eng.apply_ry(ang)
        eng.apply_layers(0.05, 2, 1)
        got = eng.get_state()
    z, p = oc.term_list(Kc, wc)
    psi = oc.evolve_product_formula(oc.initial_state(wc), z, p, 0.1, 2, 1)
    for q in range(n):
        on.apply_ry(psi, n, q, float(ang[q]))
    psi = oc.evolve_product_formula(psi, z, p, 0.05, 1, 2)
    assert 1.0 - oc.fidelity(got, psi) < FID_TOL
    assert np.abs(got - psi).max() < 1e-9


@pytest.mark.parametrize("n,evolution,order", [(4, "trotter", 1), (8, "trotter", 2), (13, "trotter", 1), (6, "reference", 1)])
def test_feedback_controller_history_vs_oracle(n, evolution, order):
    """RealtimeSyncFeedbackController.run on the live device state: every tick's exact and sampled quantities, the
    decisions, the Ry correction and the read-out sample equal the oracle's restatement of the reference tick
    (control/realtime_feedback.py:112-166) for the same seed"""
    K = oc.paper27_K(16)[:n, :n] * 2.0
    w = oc.OMEGA_N_16[:n]
    cfg = RealtimeFeedbackConfig(target_r=0.6, deadband=0.02, measurement_shots=64, correction_angle=0.2)
    ctl = RealtimeSyncFeedbackController(K, w, cfg, trotter_order=order, evolution=evolution)
    hist = ctl.run(6, seed=123)
    psi_gpu = ctl.statevector
    rows, psi = on.feedback_run(K, w, 6, 123, target_r=0.6, deadband=0.02, measurement_shots=64, correction_angle=0.2,
                                order=order, mode="T" if evolution == "trotter" else "E")
    assert {r["action"] for r in rows} != {"hold"}  # the correction branch is exercised
    for got, want in zip(hist, rows):
        assert got.action == want["action"] and got.index == want["index"]
        assert abs(got.r_statevector - want["r_statevector"]) < R_TOL
        assert abs(got.psi_statevector - want["psi_statevector"]) < 1e-8
        assert got.r_live == pytest.approx(want["r_live"], abs=1e-12)
        assert got.next_coupling_scale == pytest.approx(want["next_coupling_scale"], abs=1e-12)
        assert got.correction_angle == pytest.approx(want["correction_angle"], abs=1e-12)
        assert dict(got.readout_counts) == want["readout_counts"]
    assert 1.0 - oc.fidelity(psi_gpu, psi) < FID_TOL
    ctl.reset()
    assert ctl.coupling_scale == 1.0 and not ctl.history
    assert 1.0 - oc.fidelity(ctl.statevector, oc.initial_state(w)) < FID_TOL
    ctl.close()


def test_product_angle_states():
    """xyk_init_product_angles: |0...0>, |+>^n and a general Ry product state"""
    n = 13
    rng = np.random.default_rng(3)
    for ang in (np.zeros(n), np.full(n, np.pi / 2), rng.uniform(0, 2 * np.pi, n)):
        with XYKEngine(n) as eng:
            eng.init_product_angles(ang)
            got = eng.get_state()
        want = np.ones(1, dtype=np.complex128)
        for q in range(n):
            want = np.kron(np.array([np.cos(ang[q] / 2), np.sin(ang[q] / 2)]), want)
        assert np.abs(got - want).max() < 1e-14


def test_benchmark_callers_vs_oracle():
    """the four swap-in call sites of SURVEY 8(f) row 3, on the GPU solver, against the oracle"""
    from scpn_quantum_control_b200 import benchmarks as bm

    n = 10
    K, w = oc.paper27_K(n), oc.OMEGA_N_16[:n]
    out = bm.quantum_benchmark(n, t_max=0.3, dt=0.1, trotter_reps=2)
    assert out["n_trotter_steps"] == max(1, int(0.3 / 0.1)) * 2 and out["t_total_ms"] > 0.0  # the reference's int(t_max / dt)
    _, RE = oc.run(n, K, w, 0.3, 0.1, mode="E")
    row = bm.quantum_trotter_row(n, 0.3, 0.1, 5, r_exact=RE[-1])
    assert row.available and row.method == "quantum_trotter" and row.r_error_vs_exact < R_TOL
    assert set(row.to_dict()) == {"method", "backend", "available", "r_final", "r_error_vs_exact", "elapsed_ms",
                                  "unavailable_reason"}
    _, RT = oc.run(n, K, w, 0.3, 0.1, 5, order=1, mode="T")
    rowT = bm.quantum_trotter_row(n, 0.3, 0.1, 5, evolution="trotter")
    assert abs(rowT.r_final - RT[-1]) < R_TOL
    spec = {"K_nm": K.tolist(), "omega": w.tolist(), "t_max": 0.3, "dt": 0.1, "trotter_per_step": 5, "trotter_order": 1}
    sim = bm.simulate_execute(spec)
    assert sim["n_time_points"] == 4 and np.abs(np.array(sim["R"]) - RE).max() < R_TOL
    assert sim["R_final"] == sim["R"][-1] and sim["n_nodes"] == n
    # |0...0> is an eigenstate of every XX+YY term and of Z: R stays 0 exactly, in any mode
    srow = bm.statevector_scaling_row(14)
    assert srow["status"] == "ok" and abs(srow["metric_payload"]["final_R"]) < 1e-12
    assert bm.statevector_scaling_row(14, max_statevector_qubits=12)["status"] == "skipped"


def test_scaling_row_state_vs_oracle():
    """evolve(0.1, 1) from |0...0> with a superposed start instead: Ry(pi/3) product state through the same calls"""
    from scpn_quantum_control_b200 import QuantumKuramotoSolver

    n = 14
    K = oc.paper27_K(n)
    w = oc.OMEGA_N_16[:n]
    ang = np.full(n, np.pi / 3)
    s = QuantumKuramotoSolver(n, K, w, evolution="trotter")
    s.prepare_product_state(ang)
    s.evolve(0.1, trotter_steps=1)
    R, psi_ang = s.measure_order_parameter()
    got = s.statevector()
    s.close()
    psi = np.ones(1, dtype=np.complex128)
    for q in range(n):
        psi = np.kron(np.array([np.cos(ang[q] / 2), np.sin(ang[q] / 2)]), psi)
    Kc, wc = oc.canonicalise(n, K, w)
    z, p = oc.term_list(Kc, wc)
    psi = oc.evolve_product_formula(psi.astype(np.complex128), z, p, 0.1, 1, 1)
    assert 1.0 - oc.fidelity(got, psi) < FID_TOL
    assert abs(R - oc.order_parameter(psi, n)[0]) < R_TOL


@pytest.mark.parametrize("n,order", [(5, 1), (9, 2), (13, 1), (13, 2)])
def test_xxz_gate_by_gate_in_reference_term_order(n, order):
    """delta != 0, literal product formula: XX.YY then ZZ per pair (bridge/knm_hamiltonian.py:176-201) vs the oracle"""
    K, w = oc.random_problem(n, 900 + n)
    Kc, wc = oc.canonicalise(n, K, w)
    delta, dt, reps = 0.7, 0.1, 2
    with XYKEngine(n) as eng:
        eng.set_problem(Kc, wc, delta)
        eng.set_engine("simple")
        eng.init_product_state()
        eng.apply_layers(dt, order, reps)
        got = eng.get_state()
        R, _ = eng.order_parameter()
    z, p = on.term_list_xxz(Kc, wc, delta)
    psi = oc.initial_state(wc)
    for _ in range(reps):
        psi = (on.layer_order1_xxz if order == 1 else on.layer_order2_xxz)(psi, n, z, p, dt / reps)
    assert 1.0 - oc.fidelity(got, psi) < FID_TOL
    assert np.abs(got - psi).max() < 1e-9  # the global phase of the ZZ terms is kept
    assert abs(R - oc.order_parameter(psi, n)[0]) < R_TOL


@pytest.mark.parametrize("n", [13, 16, 18])
def test_xxz_tiled_grouped_diagonal_vs_oracle(n):
    """delta != 0 through the fused tile passes: the ZZ terms ride with the diagonal phase (one extra sweep per half
    layer, zz_diag_kernel) -- against the oracle's restatement of exactly that splitting"""
    K, w = oc.random_problem(n, 950 + n)
    Kc, wc = oc.canonicalise(n, K, w)
    delta, dt, reps = -0.45, 0.08, 2
    with XYKEngine(n) as eng:
        eng.set_problem(Kc, wc, delta)
        eng.set_engine("tiled")
        eng.init_product_state()
        eng.apply_layers(dt, 2, reps)
        got = eng.get_state()
        st = eng.stats()
    assert st["pass_launches"] > 0
    z, p = on.term_list_xxz(Kc, wc, delta)
    psi = oc.initial_state(wc)
    for _ in range(reps):
        psi = on.layer_order2_xxz_grouped(psi, n, z, p, dt / reps)
    assert 1.0 - oc.fidelity(got, psi) < FID_TOL
    assert np.abs(got - psi).max() < 1e-9


@pytest.mark.parametrize("n", [6, 13])
def test_xxz_solver_reference_mode_vs_exact(n):
    """QuantumKuramotoSolver(anisotropy_delta=...) in its default (exact per step) mode vs expm_multiply of the XXZ H"""
    from scpn_quantum_control_b200 import QuantumKuramotoSolver

    K = oc.paper27_K(16)[:n, :n]
    w = oc.OMEGA_N_16[:n]
    s = QuantumKuramotoSolver(n, K, w, anisotropy_delta=1.0)
    res = s.run(0.3, 0.1)
    s.close()
    _, RE = on.run_xxz(n, K, w, 1.0, 0.3, 0.1, mode="E")
    assert np.abs(res["R"] - RE).max() < R_TOL
    st = QuantumKuramotoSolver(n, K, w, anisotropy_delta=1.0, evolution="trotter", trotter_order=2)
    rt = st.run(0.3, 0.1, 3)
    st.close()
    _, RT = on.run_xxz(n, K, w, 1.0, 0.3, 0.1, 3, order=2, mode="T")
    assert np.abs(rt["R"] - RT).max() < R_TOL


@pytest.mark.parametrize("n", [4, 8, 13])
def test_floquet_evolve_vs_oracle(n):
    """floquet_evolve (phase/floquet_kuramoto.py:94-182) on the device vs the oracle's exact midpoint propagation"""
    from scpn_quantum_control_b200.floquet_kuramoto import floquet_evolve, scan_drive_amplitude

    K = oc.paper27_K(16)[:n, :n]
    top = K / K.max()
    w = oc.OMEGA_N_16[:n]
    periods, spp = (2, 8) if n < 13 else (1, 4)
    got = floquet_evolve(top, w, 0.5, 0.6, 2.0, periods, spp)
    want = on.floquet_evolve(top, w, 0.5, 0.6, 2.0, periods, spp)
    assert np.allclose(got.times, want["times"], rtol=0, atol=1e-14)
    assert np.allclose(got.drive_signal, want["drive_signal"], rtol=0, atol=1e-14)
    # the phase-coherence R divides by |<X>+i<Y>| per qubit: a little less well conditioned than the solver's R
    assert np.abs(got.R_values - want["R_values"]).max() < 1e-8
    assert got.subharmonic_ratio == pytest.approx(want["subharmonic_ratio"], rel=1e-6, abs=1e-9)
    assert got.is_dtc_candidate == (want["subharmonic_ratio"] > 0.1)
    if n == 4:
        sc = scan_drive_amplitude(top, w, 0.5, 2.0, amplitudes=[0.2, 1.0], n_periods=2, steps_per_period=8)
        assert sc["amplitude"] == [0.2, 1.0] and len(sc["mean_R"]) == 2
        ref = on.floquet_evolve(top, w, 0.5, 1.0, 2.0, 2, 8)
        assert sc["subharmonic_ratio"][1] == pytest.approx(ref["subharmonic_ratio"], rel=1e-6, abs=1e-9)


@pytest.mark.parametrize("n,delta,dt", [(4, 0.0, 0.1), (9, 0.0, 0.3), (14, 0.0, 0.1), (10, 0.8, 0.2), (14, 0.0, -0.15)])
def test_chebyshev_exact_step_vs_expm_multiply(n, delta, dt):
    """xyk_exact_step (Chebyshev series on the fused H.phi sweep) vs scipy's expm_multiply of the oracle's sparse H --
    the reference twin hardware/classical.py:201-208 -- after a tiled evolution (permuted layout)"""
    K, w = oc.random_problem(n, 300 + n)
    Kc, wc = oc.canonicalise(n, K, w)
    with XYKEngine(n) as eng:
        eng.set_problem(Kc, wc, delta)
        eng.init_product_state()
        if n >= 13 and delta == 0.0:
            eng.set_engine("tiled")
            eng.apply_layers(0.05, 2, 1)  # leaves a permuted layout behind
        terms = eng.exact_step(dt, 1e-14)
        got = eng.get_state()
        R, _ = eng.order_parameter()
    psi = oc.initial_state(wc)
    if n >= 13 and delta == 0.0:
        z, p = oc.term_list(Kc, wc)
        psi = oc.evolve_product_formula(psi, z, p, 0.05, 1, 2)
    psi = oc.exact_step(psi, on.hamiltonian_sparse_xxz(Kc, wc, delta), dt)
    assert terms > 3
    assert 1.0 - oc.fidelity(got, psi) < FID_TOL
    assert np.abs(got - psi).max() < 1e-10
    assert abs(R - oc.order_parameter(psi, n)[0]) < R_TOL


def test_chebyshev_solver_vs_goldens_and_composition():
    """evolution="chebyshev" reproduces the reference's stored n=8 run (Mode E) and agrees with the calibrated
    6th-order composition at n=18 where no CPU oracle is cheap"""
    import json
    import os

    from scpn_quantum_control_b200 import QuantumKuramotoSolver

    n = 8
    K, w = oc.paper27_K(n), oc.OMEGA_N_16[:n]
    s = QuantumKuramotoSolver(n, K, w, evolution="chebyshev")
    res = s.run(1.0, 0.1)
    s.close()
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")))
    _, RE = oc.run(n, K, w, 1.0, 0.1, mode="E")
    assert np.abs(res["R"] - RE).max() < R_TOL
    assert "chebyshev_terms_per_step" in res.metadata
    flat = json.dumps(gold)
    assert "0.25916072998431" in flat and abs(res["R"][-1] - 0.259160729984313) < R_TOL  # reproducible_comparison_n8.json r_final
    n = 18
    K, w = oc.random_problem(n, 77)
    a = QuantumKuramotoSolver(n, K, w, evolution="chebyshev")
    ra = a.run(0.2, 0.1)
    a.close()
    b = QuantumKuramotoSolver(n, K, w)
    rb = b.run(0.2, 0.1)
    b.close()
    assert np.abs(ra["R"] - rb["R"]).max() < R_TOL


C64_FID_TOL = 1e-5   # stated bound of the complex64 variant: fidelity >= 1 - 1e-5,
C64_R_TOL = 1e-4     # |dR| <= 1e-4 for runs of <= 25 000 gates; coefficients and reductions stay double


@pytest.mark.parametrize("n,reps,order,engine", [(14, 4, 1, "simple"), (14, 4, 1, "tiled"), (13, 2, 2, "tiled"), (16, 3, 2, "tiled"),
                                                 (20, 2, 1, "tiled"), (24, 1, 1, "tiled"), (24, 1, 1, "simple")])
def test_complex64_state_within_stated_bound(n, reps, order, engine):
    """complex64 state (8 B per amplitude in HBM; FP64 coefficients, on-chip arithmetic and reductions) vs the FP64 C
    oracle: gate-by-gate kernels and the fused tile passes (float2 tiles widened to FP64 in shared memory, narrowed in the
    fused store)"""
    import ctypes
    import os

    K, w = oc.random_problem(n, 640 + n)
    Kc, wc = oc.canonicalise(n, K, w)
    with XYKEngine(n, dtype="complex64") as eng:
        eng.set_problem(Kc, wc)
        eng.set_engine(engine)
        eng.init_product_state()
        eng.apply_layers(0.1, order, reps)
        R, _ = eng.order_parameter()
        ex, ey = eng.xy_expectations()
        got = eng.get_state()
        nrm = eng.norm2()
        st = eng.stats()
    assert got.dtype == np.complex64
    assert (st["pass_launches"] > 0) == (engine == "tiled")
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "libxyoracle.so")
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    psi = np.zeros(2 << n)
    lib.xyo_init_state(n, wc.ctypes.data_as(dp), psi.ctypes.data_as(dp))
    lib.xyo_evolve.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, dp]
    lib.xyo_evolve(n, np.ascontiguousarray(Kc).ctypes.data_as(dp), wc.ctypes.data_as(dp), 0.1, reps, order, psi.ctypes.data_as(dp))
    ref = psi.view(np.complex128)
    assert 1.0 - oc.fidelity(got.astype(np.complex128), ref) < C64_FID_TOL
    Rref, _ = oc.order_parameter(ref, n)
    assert abs(R - Rref) < C64_R_TOL
    assert abs(nrm - 1.0) < 1e-4


def test_complex64_solver_run():
    from scpn_quantum_control_b200 import QuantumKuramotoSolver

    n = 14
    K, w = oc.random_problem(n, 655)
    s = QuantumKuramotoSolver(n, K, w, dtype="complex64", evolution="trotter")
    res = s.run(0.4, 0.1, 3)
    s.close()
    _, RT = oc.run(n, K, w, 0.4, 0.1, 3, order=1, mode="T")
    assert res.metadata["dtype"] == "complex64"
    assert np.abs(res["R"] - RT).max() < C64_R_TOL


def test_next_rows_at_scale_properties():
    """size-independent properties where no CPU oracle is cheap: XXZ through the tiled engine at n = 26 (norm, <sum Z>
    conserved by every XX+YY+ZZ term, order-2 echo back to the product state), the Chebyshev step at n = 24 (norm, <sum Z>,
    forward-backward echo) and the complex64 tile passes at n = 27 (echo within the stated bound)"""
    n = 26
    K, w = oc.random_problem(n, 2600)
    theta = np.mod(w, 2 * np.pi)
    with XYKEngine(n) as eng:
        eng.set_problem(K, w, 0.6)
        eng.set_engine("tiled")
        eng.init_product_state()
        eng.apply_layers(0.05, 2, 1)
        assert abs(eng.norm2() - 1.0) < 1e-10
        assert abs(eng.z_expectations().sum() - np.cos(theta).sum()) < 1e-9
        eng.apply_layers(-0.05, 2, 1)
        ex, ey = eng.xy_expectations()
        assert max(np.abs(ex - np.sin(theta)).max(), np.abs(ey).max()) < 1e-9
        assert eng.stats()["pass_launches"] > 0
    n = 24
    K, w = oc.random_problem(n, 2400)
    theta = np.mod(w, 2 * np.pi)
    with XYKEngine(n) as eng:
        eng.set_problem(K, w)
        eng.init_product_state()
        terms = eng.exact_step(0.02, 1e-14)
        assert terms > 3 and abs(eng.norm2() - 1.0) < 1e-10
        assert abs(eng.z_expectations().sum() - np.cos(theta).sum()) < 1e-9
        eng.exact_step(-0.02, 1e-14)
        ex, ey = eng.xy_expectations()
        assert max(np.abs(ex - np.sin(theta)).max(), np.abs(ey).max()) < 1e-9
    n = 27
    K, w = oc.random_problem(n, 2700)
    theta = np.mod(w, 2 * np.pi)
    with XYKEngine(n, dtype="complex64") as eng:
        eng.set_problem(K, w)
        eng.set_engine("tiled")
        eng.init_product_state()
        eng.apply_layers(0.05, 2, 1)
        eng.apply_layers(-0.05, 2, 1)
        ex, ey = eng.xy_expectations()
        assert max(np.abs(ex - np.sin(theta)).max(), np.abs(ey).max()) < C64_R_TOL
        assert abs(eng.norm2() - 1.0) < 1e-4


def test_complex64_next_row_kernels():
    """the float instantiations of the ZZ diagonal sweep, the per-pair ZZ gate and the Ry layer (complex64 storage)"""
    n = 13
    K, w = oc.random_problem(n, 1300)
    Kc, wc = oc.canonicalise(n, K, w)
    delta = 0.5
    z, p = on.term_list_xxz(Kc, wc, delta)
    ang = np.linspace(-0.3, 0.3, n)
    for engine in ("tiled", "simple"):
        with XYKEngine(n, dtype="complex64") as eng:
            eng.set_problem(Kc, wc, delta)
            eng.set_engine(engine)
            eng.init_product_state()
            eng.apply_layers(0.1, 2, 1)
            eng.apply_ry(ang)
            got = eng.get_state().astype(np.complex128)
        psi = oc.initial_state(wc)
        psi = (on.layer_order2_xxz_grouped if engine == "tiled" else on.layer_order2_xxz)(psi, n, z, p, 0.1)
        for q in range(n):
            on.apply_ry(psi, n, q, float(ang[q]))
        assert 1.0 - oc.fidelity(got, psi) < C64_FID_TOL, engine
        assert np.abs(got - psi).max() < 1e-5, engine   # global phase of the ZZ terms kept
    with XYKEngine(n, dtype="complex64") as eng:
        eng.set_problem(Kc, wc)
        eng.init_product_state()
        with pytest.raises(RuntimeError):
            eng.exact_step(0.1)   # the Chebyshev step needs a complex128 state


@pytest.mark.parametrize("n,engine,dtype", [(7, "simple", "complex128"), (13, "tiled", "complex128"), (16, "tiled", "complex128"),
                                           (13, "tiled", "complex64")])
def test_xxz_energy_vs_oracle(n, engine, dtype):
    """<H> of the XXZ model: XY correlations + <Z> + the ZZ part from ONE sweep over the diagonal quadratic form
    (zz_energy_kernel), on a permuted layout, against <psi|H|psi> of the oracle's sparse XXZ Hamiltonian"""
    K, w = oc.random_problem(n, 3100 + n)
    Kc, wc = oc.canonicalise(n, K, w)
    delta = 0.8
    with XYKEngine(n, dtype=dtype) as eng:
        eng.set_problem(Kc, wc, delta)
        eng.set_engine(engine)
        eng.init_product_state()
        e0 = eng.energy()
        eng.apply_layers(0.06, 2, 2)
        e1 = eng.energy()
        psi = eng.get_state().astype(np.complex128)
    H = on.hamiltonian_sparse_xxz(Kc, wc, delta)
    psi0 = oc.initial_state(wc)
    tol = 1e-9 if dtype == "complex128" else 1e-3
    assert abs(e0 - float(np.real(np.vdot(psi0, H @ psi0)))) < tol
    assert abs(e1 - float(np.real(np.vdot(psi, H @ psi))) / float(np.real(np.vdot(psi, psi)))) < tol
    assert abs(e1 - e0) < 0.05 * max(1.0, abs(e0))   # a second-order step conserves the energy up to its splitting error


def test_next_rows_vs_committed_golden_vectors():
    """the GPU path against tests/golden/next_rows_vectors.json (feedback history, XXZ traces, Floquet drive)"""
    import json
    import os

    from scpn_quantum_control_b200 import QuantumKuramotoSolver
    from scpn_quantum_control_b200.floquet_kuramoto import floquet_evolve

    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next_rows_vectors.json")))
    fb = g["feedback_n4_seed123_trotter_order1"]
    ctl = RealtimeSyncFeedbackController(oc.paper27_K(16)[:4, :4] * 2.0, oc.OMEGA_N_16[:4], RealtimeFeedbackConfig(**fb["config"]),
                                         trotter_order=1, evolution="trotter")
    hist = ctl.run(fb["n_steps"], seed=fb["seed"])
    ctl.close()
    for got, want in zip(hist, fb["history"]):
        assert got.action == want["action"] and dict(got.readout_counts) == want["readout_counts"]
        assert abs(got.r_statevector - want["r_statevector"]) < R_TOL
        assert got.r_live == pytest.approx(want["r_live"], abs=1e-12)
        assert got.next_coupling_scale == pytest.approx(want["next_coupling_scale"], abs=1e-12)
    x = g["xxz_n6"]
    K, w = oc.random_problem(6, 11)
    for order in (1, 2):
        s = QuantumKuramotoSolver(6, K, w, anisotropy_delta=x["delta"], evolution="trotter", trotter_order=order)
        R = s.run(x["t_max"], x["dt"], 4)["R"]
        s.close()
        assert np.abs(R - np.array(x[f"R_trotter_order{order}_4_per_step"])).max() < R_TOL
    s = QuantumKuramotoSolver(6, K, w, anisotropy_delta=x["delta"])
    R = s.run(x["t_max"], x["dt"])["R"]
    s.close()
    assert np.abs(R - np.array(x["R_exact"])).max() < R_TOL
    f = g["floquet_n4"]
    Kp = oc.paper27_K(16)[:4, :4]
    fl = floquet_evolve(Kp / Kp.max(), oc.OMEGA_N_16[:4], f["K_base"], f["drive_amplitude"], f["drive_frequency"], f["n_periods"],
                        f["steps_per_period"])
    assert np.abs(fl.R_values - np.array(f["R_values"])).max() < 1e-8
    assert np.allclose(fl.drive_signal, f["drive_signal"], rtol=0, atol=1e-14)


def test_mcwf_building_blocks_on_the_device():
    """xyk_apply_occupation_decay and xyk_apply_jump on a permuted layout against NumPy"""
    n = 13
    K, w = oc.random_problem(n, 1313)
    Kc, wc = oc.canonicalise(n, K, w)
    with XYKEngine(n) as eng:
        eng.set_problem(Kc, wc)
        eng.set_engine("tiled")
        eng.init_product_state()
        eng.apply_layers(0.05, 1, 1)
        ref = eng.get_state()
        eng.apply_occupation_decay(0.3, 0.1, 0.7)
        got = eng.get_state()
        idx = np.arange(1 << n)
        ones = np.array([bin(int(k)).count("1") for k in idx])
        want = ref * (0.7 * np.exp(-0.3 * (n - ones) - 0.1 * ones))
        assert np.abs(got - want).max() < 1e-14
        eng.apply_jump(5, 0)
        m = 1 << 5
        w2 = np.zeros_like(want)
        lower = idx[(idx & m) == 0]
        w2[lower | m] = want[lower]
        assert np.abs(eng.get_state() - w2).max() < 1e-14
        eng.apply_jump(11, 1)
        w3 = np.where(idx & (1 << 11), -w2, w2)
        assert np.abs(eng.get_state() - w3).max() < 1e-14


@pytest.mark.parametrize("n,gamma_amp,gamma_deph,seed", [(4, 0.6, 0.4, 7), (6, 1.5, 0.5, 11), (13, 0.8, 0.3, 2)])
def test_mcwf_trajectory_vs_oracle(n, gamma_amp, gamma_deph, seed):
    """mcwf_trajectory on the device (exact step + bit-count damping + jump kernels) vs the oracle's literal restatement
    of phase/tensor_jump.py:120-219 for the same seed: R history, jump count, final state"""
    from scpn_quantum_control_b200.tensor_jump import mcwf_ensemble, mcwf_trajectory

    K = oc.paper27_K(16)[:n, :n] * 1.5
    w = oc.OMEGA_N_16[:n]
    t_max, dt = (1.0, 0.1) if n < 13 else (0.3, 0.1)
    got = mcwf_trajectory(K, w, gamma_amp, gamma_deph, t_max, dt, seed)
    want = on.mcwf_trajectory(K, w, gamma_amp, gamma_deph, t_max, dt, seed)
    assert got["n_jumps"] == want["n_jumps"]
    assert np.abs(got["R"] - want["R"]).max() < R_TOL
    assert 1.0 - oc.fidelity(got["psi_final"], want["psi_final"]) < FID_TOL
    if n == 4:
        ens = mcwf_ensemble(K, w, gamma_amp, gamma_deph, 0.5, 0.1, n_trajectories=3, seed=5)
        seeds = np.random.default_rng(5).integers(0, 2**31, size=3)
        ref = np.array([on.mcwf_trajectory(K, w, gamma_amp, gamma_deph, 0.5, 0.1, int(sd))["R"] for sd in seeds])
        assert np.abs(ens["R_trajectories"] - ref).max() < R_TOL
        assert np.abs(ens["R_mean"] - ref.mean(axis=0)).max() < R_TOL and ens["n_trajectories"] == 3
