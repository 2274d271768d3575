"""CPU tests of the SURVEY 8(f) "next" rows: host logic of the feedback controller mirror against the oracle's
restatement of the reference policy (control/realtime_feedback.py:316-341), and the next-rows oracle against itself
(XXZ Mode T converges to Mode E, delta = 0 reduces to the path oracle, Floquet with zero drive = static exact run)."""
import numpy as np
import pytest

from oracle import xy_oracle as oc
from oracle import xy_oracle_next as on
from scpn_quantum_control_b200.realtime_feedback import RealtimeFeedbackConfig, feedback_policy_numpy


def test_feedback_policy_matches_oracle_scalar_policy():
    rng = np.random.default_rng(5)
    r = np.concatenate([rng.uniform(0, 1, 200), [0.75, 0.71, 0.79, 0.0, 1.0]])
    for (t, d, g, m) in [(0.75, 0.04, 0.8, 2.0), (0.5, 0.0, 3.0, 1.5), (0.9, 0.2, 0.0, 1.0)]:
        a, gains, e = feedback_policy_numpy(r, t, d, g, m)
        assert a.dtype == np.int32 and gains.dtype == np.float64
        for k in range(r.shape[0]):
            code, gain, err = on.feedback_policy(float(r[k]), t, d, g, m)
            assert (int(a[k]), float(gains[k]), float(e[k])) == (code, gain, err)


def test_feedback_config_validation_messages():
    with pytest.raises(ValueError, match=r"target_r must be finite and in \[0.0, 1.0\]"):
        RealtimeFeedbackConfig(target_r=1.5)
    with pytest.raises(ValueError, match="base_dt must be finite and positive"):
        RealtimeFeedbackConfig(base_dt=0.0)
    with pytest.raises(ValueError, match="trotter_steps must be a positive integer"):
        RealtimeFeedbackConfig(trotter_steps=0)
    with pytest.raises(ValueError, match="max_gain must be finite and at least 1.0"):
        RealtimeFeedbackConfig(max_gain=0.5)
    with pytest.raises(ValueError, match="feedback_correction_sign must be either -1.0 or 1.0"):
        RealtimeFeedbackConfig(feedback_correction_sign=0.5)
    with pytest.raises(ValueError, match="r_values must contain only finite values"):
        feedback_policy_numpy(np.array([np.nan]), 0.5, 0.1, 1.0, 2.0)


def test_xxz_oracle_reduces_and_converges():
    n = 6
    K, w = oc.random_problem(n, 11)
    t0, R0 = oc.run(n, K, w, 0.3, 0.1, 4, order=2, mode="T")
    t1, R1 = on.run_xxz(n, K, w, 0.0, 0.3, 0.1, 4, order=2, mode="T")
    assert np.array_equal(t0, t1) and np.max(np.abs(R0 - R1)) < 1e-14
    _, RE = on.run_xxz(n, K, w, 0.7, 0.3, 0.1, mode="E")
    err = [np.max(np.abs(on.run_xxz(n, K, w, 0.7, 0.3, 0.1, reps, order=2, mode="T")[1] - RE)) for reps in (2, 4, 8)]
    assert err[0] > err[1] > err[2] and err[1] / err[2] > 3.5  # second order


def test_floquet_oracle_zero_drive_is_static_exact_evolution():
    n = 4
    K = oc.paper27_K(n)
    w = oc.OMEGA_N_16[:n]
    out, psi = on.floquet_evolve(K / K.max(), w, 0.45, 0.0, 2.0, 1, 6, return_state=True)
    assert np.allclose(out["drive_signal"], 0.45)
    Ks = 0.45 * K / K.max()
    ref = np.full(1 << n, 0.25, dtype=np.complex128)
    H = oc.hamiltonian_sparse(Ks, w)
    for _ in range(6):
        ref = oc.exact_step(ref, H, (2 * np.pi / 2.0) / 6)
    assert 1 - oc.fidelity(psi, ref) < 1e-12
    assert out["R_values"][0] == pytest.approx(1.0)


def test_floquet_host_helpers_match_the_oracle():
    from scpn_quantum_control_b200.floquet_kuramoto import DTC_SUBHARMONIC_THRESHOLD, _phase_coherence, _subharmonic_ratio

    rng = np.random.default_rng(8)
    assert DTC_SUBHARMONIC_THRESHOLD == 0.1
    for m in (3, 16, 37, 160):
        sig = rng.uniform(0, 1, m) + 0.3 * np.cos(np.arange(m) * 0.7)
        assert _subharmonic_ratio(sig, 2.0, 0.1) == pytest.approx(on.subharmonic_ratio(sig, 2.0, 0.1), rel=1e-14, abs=0)
    assert _subharmonic_ratio(np.ones(10), 2.0, 0.1) == 0.0
    psi = oc.initial_state(oc.OMEGA_N_16[:5])
    ex, ey = oc.xy_expectations(psi, 5)
    assert _phase_coherence(ex, ey) == pytest.approx(on.phase_coherence_R(psi, 5), abs=1e-15)


def test_benchmark_caller_shims_validate_like_the_reference():
    from scpn_quantum_control_b200 import benchmarks as bm

    with pytest.raises(ValueError, match="n_oscillators must be an integer >= 2"):
        bm.quantum_trotter_row(1)
    with pytest.raises(ValueError, match="dt must be finite, positive and not exceed t_max"):
        bm.quantum_trotter_row(4, t_max=0.1, dt=0.2)
    with pytest.raises(ValueError, match="trotter_per_step must be an integer >= 1"):
        bm.quantum_trotter_row(4, trotter_per_step=0)
    row = bm.ComparisonMethodRow("quantum_trotter", "x", False, None, None, 0.0, "why")
    assert row.to_dict()["unavailable_reason"] == "why" and list(row.to_dict())[0] == "method"
    skipped = bm.statevector_scaling_row(20, max_statevector_qubits=16)
    assert skipped["status"] == "skipped" and skipped["notes"] == ["size gate"] and skipped["n_qubits"] == 20
    assert {"baseline", "wall_time_ms", "memory_bytes", "metric_payload", "machine"} <= set(skipped)   # the reference's row schema


def test_solver_rejects_bad_next_row_arguments():
    from scpn_quantum_control_b200 import QuantumKuramotoSolver

    K, w = oc.paper27_K(4), oc.OMEGA_N_16[:4]
    with pytest.raises(ValueError, match="evolution must be"):
        QuantumKuramotoSolver(4, K, w, evolution="magic")
    with pytest.raises(ValueError, match="anisotropy_delta must be finite"):
        QuantumKuramotoSolver(4, K, w, anisotropy_delta=float("nan"))
    with pytest.raises(ValueError, match="evolution='chebyshev' needs a single-device complex128 state"):
        QuantumKuramotoSolver(4, K, w, evolution="chebyshev", dtype="complex64")
    assert QuantumKuramotoSolver(4, K, w, anisotropy_delta=1e-16).anisotropy_delta == 0.0  # below the reference's sparsity threshold


def _next_rows_goldens():
    import json
    import os

    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next_rows_vectors.json")))


def test_next_rows_oracle_reproduces_the_committed_vectors():
    g = _next_rows_goldens()
    fb = g["feedback_n4_seed123_trotter_order1"]
    K, w = oc.paper27_K(16)[:4, :4] * 2.0, oc.OMEGA_N_16[:4]
    rows, _ = on.feedback_run(K, w, fb["n_steps"], fb["seed"], order=1, mode="T", **fb["config"])
    for got, want in zip(rows, fb["history"]):
        assert got["action"] == want["action"] and got["readout_counts"] == want["readout_counts"]
        for key in ("r_live", "r_statevector", "psi_statevector", "error", "next_coupling_scale", "correction_angle"):
            assert got[key] == pytest.approx(want[key], abs=1e-13)
    x = g["xxz_n6"]
    K, w = oc.random_problem(6, 11)
    for order in (1, 2):
        _, R = on.run_xxz(6, K, w, x["delta"], x["t_max"], x["dt"], 4, order=order, mode="T")
        assert np.abs(R - np.array(x[f"R_trotter_order{order}_4_per_step"])).max() < 1e-13
    _, RE = on.run_xxz(6, K, w, x["delta"], x["t_max"], x["dt"], mode="E")
    assert np.abs(RE - np.array(x["R_exact"])).max() < 1e-12
    f = g["floquet_n4"]
    Kp = oc.paper27_K(16)[:4, :4]
    fl = on.floquet_evolve(Kp / Kp.max(), oc.OMEGA_N_16[:4], f["K_base"], f["drive_amplitude"], f["drive_frequency"], f["n_periods"],
                           f["steps_per_period"])
    assert np.abs(fl["R_values"] - np.array(f["R_values"])).max() < 1e-12
    assert fl["subharmonic_ratio"] == pytest.approx(f["subharmonic_ratio"], rel=1e-9)


class _NumpyDevice:
    """NumPy twin of tensor_jump._Device (same methods, little-endian index bits like the engine): lets the trajectory logic
    of the product module run on the CPU against the oracle's literal restatement of the reference."""

    def __init__(self, n, K, omega):
        self.n = n
        Ks = (K + K.T) / 2.0
        self.H = oc.hamiltonian_sparse(Ks, omega)
        self.psi = None

    def prepare(self, angles):
        psi = np.ones(1, dtype=np.complex128)
        for q in range(self.n):   # bit q = qubit q: later factors are less significant in np.kron -> build from the top
            psi = np.kron(np.array([np.cos(angles[q] / 2), np.sin(angles[q] / 2)], dtype=np.complex128), psi)
        self.psi = psi

    def exact_step(self, step_dt, n_steps):
        self.psi = oc.exact_step(self.psi, self.H, step_dt)

    def decay(self, rate_zero, scale):
        idx = np.arange(self.psi.shape[0])
        ones = np.array([bin(int(k)).count("1") for k in idx])
        self.psi = self.psi * (scale * np.exp(-rate_zero * (self.n - ones)))

    def jump(self, bit, kind):
        idx = np.arange(self.psi.shape[0])
        m = 1 << bit
        if kind == 0:
            out = np.zeros_like(self.psi)
            lower = idx[(idx & m) == 0]
            out[lower | m] = self.psi[lower]
            self.psi = out
        else:
            self.psi = np.where(idx & m, -self.psi, self.psi)

    def norm2(self):
        return float(np.vdot(self.psi, self.psi).real)

    def z_expectations(self):
        return oc.z_expectations(self.psi, self.n)

    def xy_sum(self):
        ex, ey = oc.xy_expectations(self.psi, self.n)
        return complex(ex.sum(), ey.sum())

    def snapshot(self):
        return self.psi.copy()

    def restore(self, snap):
        self.psi = snap.copy()

    def state(self):
        return self.psi.copy()


@pytest.mark.parametrize("n,gamma_amp,gamma_deph,seed", [(3, 0.9, 0.0, 1), (4, 0.6, 0.4, 7), (5, 0.0, 1.2, 3), (4, 0.0, 0.0, 5), (5, 1.5, 0.5, 11)])
def test_mcwf_trajectory_logic_vs_oracle(n, gamma_amp, gamma_deph, seed):
    """tensor_jump._run_trajectory (factorised step: exact unitary step, bit-count damping, jump kernels on index bit
    n-1-i) on the NumPy device twin == the oracle's literal restatement of the reference (sparse H_eff, Kronecker-order
    operators): same R history, same jumps, same final state"""
    from scpn_quantum_control_b200 import tensor_jump as tj

    K = oc.paper27_K(16)[:n, :n] * 1.5
    w = oc.OMEGA_N_16[:n]
    want = on.mcwf_trajectory(K, w, gamma_amp, gamma_deph, 1.0, 0.1, seed)
    got = tj._run_trajectory(_NumpyDevice(n, K, w), n, w, gamma_amp, gamma_deph, 1.0, 0.1, np.random.default_rng(seed))
    assert got["n_jumps"] == want["n_jumps"]
    assert np.abs(got["R"] - want["R"]).max() < 1e-9
    assert 1 - oc.fidelity(got["psi_final"], want["psi_final"]) < 1e-9
    assert np.allclose(got["times"], want["times"])
    if gamma_amp + gamma_deph >= 0.9:
        assert want["n_jumps"] > 0, "the parameters are meant to exercise the jump branch"


def test_mcwf_validation_messages():
    from scpn_quantum_control_b200.tensor_jump import mcwf_ensemble, mcwf_trajectory

    K, w = oc.paper27_K(4), oc.OMEGA_N_16[:4]
    with pytest.raises(ValueError, match="K must be a square two-dimensional coupling matrix."):
        mcwf_trajectory(K[:3], w)
    with pytest.raises(ValueError, match="omega must be a one-dimensional vector matching K."):
        mcwf_trajectory(K, w[:3])
    with pytest.raises(ValueError, match="gamma_amp must be finite and non-negative."):
        mcwf_trajectory(K, w, gamma_amp=-1.0)
    with pytest.raises(ValueError, match="dt must be finite and positive."):
        mcwf_trajectory(K, w, dt=0.0)
    with pytest.raises(ValueError, match="n_trajectories must be a positive integer."):
        mcwf_ensemble(K, w, n_trajectories=True)
    with pytest.raises(ValueError, match="n_trajectories must be at least 1."):
        mcwf_ensemble(K, w, n_trajectories=0)


def test_mcwf_oracle_and_trajectory_logic_reproduce_the_committed_vector():
    from scpn_quantum_control_b200 import tensor_jump as tj

    m = _next_rows_goldens()["mcwf_n4_seed7"]
    K, w = oc.paper27_K(16)[:4, :4] * 1.5, oc.OMEGA_N_16[:4]
    want_psi = np.array(m["psi_final_re"]) + 1j * np.array(m["psi_final_im"])
    tr = on.mcwf_trajectory(K, w, m["gamma_amp"], m["gamma_deph"], m["t_max"], m["dt"], m["seed"])
    assert tr["n_jumps"] == m["n_jumps"] and np.abs(tr["R"] - np.array(m["R"])).max() < 1e-12
    got = tj._run_trajectory(_NumpyDevice(4, K, w), 4, w, m["gamma_amp"], m["gamma_deph"], m["t_max"], m["dt"],
                             np.random.default_rng(m["seed"]))
    assert got["n_jumps"] == m["n_jumps"] and np.abs(got["R"] - np.array(m["R"])).max() < 1e-9
    assert 1 - oc.fidelity(got["psi_final"], want_psi) < 1e-9
