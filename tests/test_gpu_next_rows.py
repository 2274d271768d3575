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


@pytest.mark.parametrize("n,engine", [(5, "gates"), (13, "tiled"), (16, "tiled")])
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
