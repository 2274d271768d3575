"""Host-side mirror of the reference interface: validation messages, result type, time grid,
budget gate, selector, and the C-ABI surface (no GPU needed)."""

import ctypes
import os
import re

import numpy as np
import pytest

import scpn_quantum_control_b200 as pkg
from scpn_quantum_control_b200 import (DenseAllocationError, QuantumKuramotoSolver, TrajectoryResult,
                                       TrotterEvolutionConfig, recommend_backend, time_grid, xy_terms)
from scpn_quantum_control_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    header = open(os.path.join(ROOT, "include", "xyk.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(xyk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libxyk.so does not export {name}"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert _lib.load().xyk_version() >= 100


def test_constructor_validation_messages():
    K = np.zeros((3, 3))
    w = np.ones(3)
    for bad_n in (0, -1, 2.0, "3"):
        with pytest.raises(ValueError, match="n_oscillators must be a positive integer"):
            QuantumKuramotoSolver(bad_n, K, w)
    with pytest.raises(ValueError, match=r"K_coupling shape must be \(3, 3\), got \(2, 2\)"):
        QuantumKuramotoSolver(3, np.zeros((2, 2)), w)
    with pytest.raises(ValueError, match=r"omega_natural shape must be \(3,\), got \(2,\)"):
        QuantumKuramotoSolver(3, K, np.ones(2))
    Kn = K.copy()
    Kn[0, 1] = np.nan
    with pytest.raises(ValueError, match="K_coupling must contain only finite values"):
        QuantumKuramotoSolver(3, Kn, w)
    with pytest.raises(ValueError, match="omega_natural must contain only finite values"):
        QuantumKuramotoSolver(3, K, np.array([1.0, np.inf, 0.0]))
    Ka = K.copy()
    Ka[0, 1] = 1.0
    with pytest.raises(ValueError, match="K_coupling must be symmetric for the XY Kuramoto solver"):
        QuantumKuramotoSolver(3, Ka, w)
    with pytest.raises(ValueError, match="K_coupling must contain real numeric scalars"):
        QuantumKuramotoSolver(3, K.astype(bool), w)
    with pytest.raises(ValueError, match="omega_natural must contain real numeric scalars"):
        QuantumKuramotoSolver(3, K, w.astype(complex))
    with pytest.raises(ValueError, match="K_coupling must contain real numeric scalars"):
        QuantumKuramotoSolver(3, [["a"] * 3] * 3, w)
    with pytest.raises(ValueError, match="trotter_order must be 1 or 2, got 3"):
        QuantumKuramotoSolver(3, K, w, trotter_order=3)
    with pytest.raises(ValueError, match="order must be 1 or 2"):
        TrotterEvolutionConfig(order=4)


def test_diagonal_zeroed_without_mutating_input():
    K = np.full((3, 3), 0.5)
    s = QuantumKuramotoSolver(3, K, np.ones(3))
    assert np.all(np.diag(s.K) == 0.0)
    assert np.all(np.diag(K) == 0.5)


def test_run_argument_validation():
    s = QuantumKuramotoSolver(2, np.zeros((2, 2)), np.ones(2))
    with pytest.raises(ValueError, match="t_max must be finite and non-negative, got -1"):
        s.run(-1.0, 0.1)
    with pytest.raises(ValueError, match="dt must be finite and positive, got 0"):
        s.run(1.0, 0.0)
    with pytest.raises(ValueError, match="dt must be finite and positive, got nan"):
        s.run(1.0, float("nan"))
    with pytest.raises(ValueError, match="trotter_per_step must be a positive integer, got 0"):
        s.run(1.0, 0.1, trotter_per_step=0)
    with pytest.raises(ValueError, match="time must be finite and non-negative"):
        s.evolve(-0.5)
    with pytest.raises(ValueError, match="trotter_steps must be a positive integer"):
        s.evolve(0.5, 0)


def test_budget_gate_fires_before_allocation():
    """reference tests/test_xy_kuramoto.py:172-186: DenseAllocationError before any allocation."""
    s = QuantumKuramotoSolver(20, np.zeros((20, 20)), np.ones(20))
    with pytest.raises(DenseAllocationError, match="Kuramoto statevector trajectory for n=20"):
        s.run(0.1, 0.1, max_statevector_gib=0.001)
    assert s._engine is None
    assert issubclass(DenseAllocationError, MemoryError)


def test_budget_env(monkeypatch):
    monkeypatch.setenv("SCPN_MAX_DENSE_GIB", "0.0001")
    s = QuantumKuramotoSolver(18, np.zeros((18, 18)), np.ones(18))
    with pytest.raises(DenseAllocationError):
        s.run(0.1, 0.1)
    monkeypatch.setenv("SCPN_MAX_DENSE_GIB", "abc")
    with pytest.raises(ValueError, match="SCPN_MAX_DENSE_GIB must be a positive number"):
        s.run(0.1, 0.1)


def test_time_grid_matches_reference_example():
    times, steps = time_grid(0.35, 0.1)
    np.testing.assert_allclose(times, [0.0, 0.1, 0.2, 0.3, 0.35])
    np.testing.assert_allclose(steps, [0.1, 0.1, 0.1, 0.05])
    times, steps = time_grid(1.0, 0.1)
    assert times.shape == (11,) and times[-1] == 1.0
    times, steps = time_grid(0.0, 0.1)
    assert times.tolist() == [0.0]
    times, _ = time_grid(0.05, 0.1)
    assert times.tolist() == [0.0, 0.05]


def test_trajectory_result_contract():
    r = TrajectoryResult(times=[0.0, 0.1], R=[0.5, 0.4], metadata={"backend": "gpu_statevector"})
    assert list(r) == ["times", "R"] and len(r) == 2
    assert r["R"][1] == 0.4 and r.R is r["R"]
    assert set(r.to_dict()) == {"times", "R"}
    with pytest.raises(KeyError):
        r["x"]
    with pytest.raises(ValueError):
        r.R[0] = 1.0
    with pytest.raises(TypeError):
        r.metadata["k"] = 1
    with pytest.raises(ValueError, match="identical shape"):
        TrajectoryResult(times=[0.0], R=[0.1, 0.2])
    with pytest.raises(ValueError, match="finite"):
        TrajectoryResult(times=[0.0], R=[np.nan])
    with pytest.raises(ValueError, match="one-dimensional"):
        TrajectoryResult(times=[[0.0]], R=[[0.1]])


def test_recommend_backend_statevector_labels():
    assert recommend_backend(30, has_gpu=True)["backend"] == "gpu_statevector"
    assert recommend_backend(33, has_gpu=True, n_gpus=8)["backend"] == "gpu_statevector"
    assert recommend_backend(25, has_gpu=False)["backend"] == "statevector"
    assert recommend_backend(40, has_gpu=False)["backend"] == "hardware"
    with pytest.raises(ValueError, match="positive integer qubit count"):
        recommend_backend(0)
    with pytest.raises(ValueError, match="ram_gb must be finite and positive"):
        recommend_backend(4, ram_gb=-1)


def test_term_order_and_thresholds():
    """XX,YY coefficient -K, Z coefficient -omega; lexicographic pairs; 1e-15 thresholds
    (reference bridge/knm_hamiltonian.py:172-195)."""
    K = np.array([[0, 0.5, 1e-16, 0.2], [0.5, 0, 0.3, 0], [1e-16, 0.3, 0, 0.1], [0.2, 0, 0.1, 0]], dtype=float)
    w = np.array([1.0, 0.0, -2.0, 1e-16])
    z, p = xy_terms(K, w)
    assert z == [(0, 1.0), (2, -2.0)]
    assert [(i, j) for i, j, _ in p] == [(0, 1), (0, 3), (1, 2), (2, 3)]
    lib = _lib.load()
    # the native term list agrees (plan of one first-order layer lists exactly these gates)
    plan = pkg.engine.plan_describe(K, w, 0.1, 1, 1)
    gates = [(g[0], g[1]) for ps in plan["passes"] for r in ps["rounds"] for g in r["gates"]]
    assert sorted(gates) == [(0, 1), (0, 3), (1, 2), (2, 3)]
    for ps in plan["passes"]:
        for r in ps["rounds"]:
            for i, j, phi in r["gates"]:
                assert phi == pytest.approx(2.0 * 0.5 * (K[i, j] + K[j, i]) * 0.1)


def test_canonical_problem_builders():
    K = pkg.build_knm_paper27(16)
    assert K[0, 1] == 0.302 and K[3, 4] == 0.154 and K[0, 15] >= 0.05 and K[4, 6] >= 0.15
    assert np.allclose(K, K.T)
    Kr, w = pkg.build_kuramoto_ring(4, coupling=1.0, omega=pkg.OMEGA_N_16[:4])
    assert Kr[0, 1] == Kr[0, 3] == Kr[1, 2] == Kr[2, 3] == 1.0 and Kr[0, 2] == 0.0
    assert pkg.omega_for_oscillators(18)[16] == pkg.OMEGA_N_16[0]


def test_no_cpu_fallback_without_device():
    """Without a visible GPU the product path raises instead of computing on the CPU."""
    from scpn_quantum_control_b200.engine import device_count

    if device_count() > 0:
        pytest.skip("GPU present")
    s = QuantumKuramotoSolver(4, np.zeros((4, 4)), np.ones(4), evolution="trotter")
    with pytest.raises((RuntimeError, MemoryError)):
        s.run(0.1, 0.1)
