"""Monte Carlo wave function (quantum-jump) trajectories of the open Kuramoto-XY system on the GPU statevector engine
(drop-in for ``mcwf_trajectory`` / ``mcwf_ensemble`` of the reference,
src/scpn_quantum_control/phase/tensor_jump.py:120-271: same arguments, same result keys, same generator call order).

The reference propagates every step with ``expm_multiply(-i H_eff dt)``, H_eff = H - (i/2) sum L^dagger L, for the jump
operators sqrt(gamma_amp) |1><0|_q ("-" = [[0, 0], [1, 0]], :81-107) and sqrt(gamma_deph / 2) Z_q.  Their sum L^dagger L is
diagonal -- gamma_amp per zero bit plus n gamma_deph / 2 -- and commutes with H (the XY terms and Z conserve the number of
ones), so the step factorises EXACTLY: exp(-i H_eff dt) = [real factor per amplitude from its bit counts] exp(-i H dt).
The unitary factor is the engine's exact step (calibrated 6th-order composition of pair-rotation passes), the damping one
sweep of ``xyk_apply_occupation_decay``, a jump one sweep of ``xyk_apply_jump``; the state never leaves the device.

One quirk of the reference is reproduced on purpose: its state and jump operators use Kronecker (big-endian) qubit order
(:110-117, :274-291) while its Hamiltonian matrix is Qiskit's little-endian one (bridge/knm_hamiltonian.py:243-246), i.e.
oscillator i's initial angle and jump operators live on index bit n-1-i while K_ij and omega_i act on bits i, j.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from .xy_kuramoto import QuantumKuramotoSolver

_KIND_LOWER_TO_UPPER, _KIND_Z = 0, 1


def _validate_mcwf_inputs(K, omega, gamma_amp, gamma_deph, t_max, dt):
    """Same checks and messages as the reference (:50-74)."""
    K_arr = np.asarray(K, dtype=np.float64)
    omega_arr = np.asarray(omega, dtype=np.float64)
    if K_arr.ndim != 2 or K_arr.shape[0] != K_arr.shape[1]:
        raise ValueError("K must be a square two-dimensional coupling matrix.")
    if omega_arr.ndim != 1 or omega_arr.shape[0] != K_arr.shape[0]:
        raise ValueError("omega must be a one-dimensional vector matching K.")
    if not np.all(np.isfinite(K_arr)) or not np.all(np.isfinite(omega_arr)):
        raise ValueError("K and omega must contain only finite values.")
    for name, value in (("gamma_amp", gamma_amp), ("gamma_deph", gamma_deph)):
        if not np.isfinite(value) or value < 0.0:
            raise ValueError(f"{name} must be finite and non-negative.")
    if not np.isfinite(t_max) or t_max < 0.0:
        raise ValueError("t_max must be finite and non-negative.")
    if not np.isfinite(dt) or dt <= 0.0:
        raise ValueError("dt must be finite and positive.")
    return K_arr, omega_arr


def _time_step_count(t_max: float, dt: float) -> int:
    return max(1, int(np.ceil(t_max / dt)))


class _Device:
    """The handful of engine operations a trajectory needs (the NumPy twin in tests/test_next_rows_cpu.py implements the
    same methods, so the trajectory logic below is checked against the oracle without a GPU)."""

    def __init__(self, n: int, K: np.ndarray, omega: np.ndarray, device: int, max_gib: float | None, exact_tol: float):
        Ks = (K + K.T) / 2.0
        np.fill_diagonal(Ks, 0.0)
        self._solver = QuantumKuramotoSolver(n, Ks, omega, device=device, exact_tol=exact_tol)
        self._solver._budget_gate(max_gib)
        self._reps: dict[float, int] = {}
        self.n = n
        self._device = device
        self._backup = None

    def prepare(self, angles: np.ndarray) -> None:
        self._solver.prepare_product_state(angles)

    def exact_step(self, step_dt: float, n_steps: int) -> None:
        if step_dt not in self._reps:
            self._reps[step_dt] = self._solver._calibrated_exact_reps(step_dt, n_steps)
        self._solver.engine().apply_layers(step_dt, 6, self._reps[step_dt])

    def decay(self, rate_zero: float, scale: float) -> None:
        self._solver.engine().apply_occupation_decay(rate_zero, 0.0, scale)

    def jump(self, bit: int, kind: int) -> None:
        self._solver.engine().apply_jump(bit, kind)

    def norm2(self) -> float:
        return self._solver.engine().norm2()

    def z_expectations(self) -> np.ndarray:
        return self._solver.engine().z_expectations()

    def xy_sum(self) -> complex:
        ex, ey = self._solver.engine().xy_expectations()
        return complex(float(ex.sum()), float(ey.sum()))

    def snapshot(self):
        """Device-to-device copy of the live state (canonical order) into a backup buffer."""
        if self._backup is None:
            from .engine import XYKEngine

            self._backup = XYKEngine(self.n, device=self._device)
            self._backup.init_product_angles(np.zeros(self.n))
        ptr, _nbytes = self._backup.state_buffer()
        self._solver.engine().get_state_device(ptr)
        return ptr

    def restore(self, snap) -> None:
        self._solver.engine().set_state_device(snap)

    def state(self) -> np.ndarray:
        return self._solver.statevector()

    def close(self) -> None:
        if self._backup is not None:
            self._backup.close()
            self._backup = None
        self._solver.close()


def _run_trajectory(dev, n: int, omega: np.ndarray, gamma_amp: float, gamma_deph: float, t_max: float, dt: float,
                    rng: np.random.Generator) -> dict[str, Any]:
    """The reference's loop (:157-221) on a device object.  The device state is kept normalised between steps, like
    the reference's ``psi``."""
    theta = np.mod(omega, 2.0 * np.pi)
    dev.prepare(theta[::-1].copy())          # oscillator i sits on index bit n-1-i (Kronecker order of the reference)

    def order_parameter() -> float:
        return float(abs(dev.xy_sum() / n))

    if t_max == 0.0:
        return {"times": np.array([0.0]), "R": np.array([order_parameter()]), "psi_final": dev.state(), "n_jumps": 0}
    ops = []   # (bit, kind, rate) in the reference's order: per oscillator i first the amplitude, then the dephasing operator
    for i in range(n):
        if gamma_amp > 0:
            ops.append((n - 1 - i, _KIND_LOWER_TO_UPPER, gamma_amp))
        if gamma_deph > 0:
            ops.append((n - 1 - i, _KIND_Z, gamma_deph / 2.0))
    n_steps = _time_step_count(t_max, dt)
    times = np.linspace(0, t_max, n_steps + 1)
    R_history = np.zeros(n_steps + 1)
    R_history[0] = order_parameter()
    n_jumps = 0
    for step in range(1, n_steps + 1):
        step_dt = float(times[step] - times[step - 1])
        # jump statistics of the CURRENT (normalised) state -- the reference applies a jump to the state before the step
        if ops:
            ez = dev.z_expectations()
            p_zero = np.clip((1.0 + ez) / 2.0, 0.0, 1.0)          # probability of index bit b being 0
            probs = np.array([rate * (p_zero[bit] if kind == _KIND_LOWER_TO_UPPER else 1.0) for bit, kind, rate in ops])
        before = dev.snapshot() if ops else None
        dev.exact_step(step_dt, n_steps)
        dev.decay(0.5 * gamma_amp * step_dt, float(np.exp(-0.25 * gamma_deph * n * step_dt)))
        norm_sq = dev.norm2()
        dp = float(np.clip(1 - norm_sq, 0.0, 1.0))
        r = rng.uniform()
        jumped = False
        if r < dp and ops:
            total = probs.sum()
            if total > 1e-15:
                k = rng.choice(len(ops), p=probs / total)
                dev.restore(before)
                bit, kind, _rate = ops[k]
                dev.jump(bit, kind)
                after = dev.norm2()
                dev.decay(0.0, 1.0 / np.sqrt(after))
                n_jumps += 1
                jumped = True
        if not jumped and norm_sq > 1e-15:
            dev.decay(0.0, 1.0 / np.sqrt(norm_sq))
        R_history[step] = order_parameter()
    return {"times": times, "R": R_history, "psi_final": dev.state(), "n_jumps": n_jumps}


def mcwf_trajectory(K, omega, gamma_amp: float = 0.05, gamma_deph: float = 0.0, t_max: float = 1.0, dt: float = 0.05,
                    seed: int | None = None, *, max_dense_gib: float | None = None, device: int = 0,
                    exact_tol: float = 2e-10) -> dict[str, Any]:
    """Run a single MCWF trajectory; returns ``{"times", "R", "psi_final", "n_jumps"}`` (reference :120-219)."""
    K, omega = _validate_mcwf_inputs(K, omega, gamma_amp, gamma_deph, t_max, dt)
    n = K.shape[0]
    rng = np.random.default_rng(seed)
    dev = _Device(n, K, omega, device, max_dense_gib, exact_tol)
    try:
        return _run_trajectory(dev, n, omega, gamma_amp, gamma_deph, t_max, dt, rng)
    finally:
        dev.close()


def mcwf_ensemble(K, omega, gamma_amp: float = 0.05, gamma_deph: float = 0.0, t_max: float = 1.0, dt: float = 0.05,
                  n_trajectories: int = 50, seed: int | None = None, *, max_dense_gib: float | None = None,
                  device: int = 0) -> dict[str, Any]:
    """Ensemble of trajectories with the reference's child seeds (:222-271); the engine and its compiled programs are
    shared by all trajectories."""
    if isinstance(n_trajectories, bool) or not isinstance(n_trajectories, int):
        raise ValueError("n_trajectories must be a positive integer.")
    if n_trajectories < 1:
        raise ValueError("n_trajectories must be at least 1.")
    K, omega = _validate_mcwf_inputs(K, omega, gamma_amp, gamma_deph, t_max, dt)
    n = K.shape[0]
    seeds = np.random.default_rng(seed).integers(0, 2**31, size=n_trajectories)
    all_R, total_jumps = [], 0
    dev = _Device(n, K, omega, device, max_dense_gib, 2e-10)
    try:
        for i in range(n_trajectories):
            traj = _run_trajectory(dev, n, omega, gamma_amp, gamma_deph, t_max, dt, np.random.default_rng(int(seeds[i])))
            all_R.append(traj["R"])
            total_jumps += traj["n_jumps"]
    finally:
        dev.close()
    R_array = np.array(all_R)
    return {"times": np.linspace(0, t_max, _time_step_count(t_max, dt) + 1), "R_mean": np.mean(R_array, axis=0),
            "R_std": np.std(R_array, axis=0), "R_trajectories": R_array, "total_jumps": total_jumps,
            "n_trajectories": n_trajectories}
