"""Closed-loop synchronisation feedback on a LIVE device statevector.

Drop-in for the statevector half of the reference's ``RealtimeSyncFeedbackController``
(reference: src/scpn_quantum_control/control/realtime_feedback.py:83-227): every tick evolves the live state
with the coupling matrix re-scaled by the controller's current gain (:116-125 -- the reference builds a new
solver and a new circuit per tick; here ``update_problem`` re-parameterises the engine in place), measures
R exp(i psi) exactly (:127) and from finite shots (:128, :197-204), applies the reference's policy (:316-341)
and the small per-qubit Ry correction (:206-227, ``xyk_apply_ry``).  The random draws (shot sampling of <X>,
<Y> and of the computational-basis read-out) use NumPy's generator in the reference's order, so a seeded run
reproduces the reference's decisions.  The dynamic-circuit builders of the reference module (:230-314,
mid-circuit measurement templates for hardware) are Qiskit circuit objects and stay in the reference.
"""

from __future__ import annotations

from collections.abc import Mapping
from dataclasses import dataclass, field

import numpy as np

from .xy_kuramoto import QuantumKuramotoSolver

_ACTION_LABELS = {-1: "release", 0: "hold", 1: "synchronise"}
_READOUT_MAX_QUBITS = 24  # the multinomial read-out sample needs all 2^n probabilities on the host


def _require_positive(value: float, name: str, allow_zero: bool = False) -> None:
    if not np.isfinite(value) or value < 0.0 or (value == 0.0 and not allow_zero):
        qualifier = "non-negative" if allow_zero else "positive"
        raise ValueError(f"{name} must be finite and {qualifier}")


def _require_range(value: float, lower: float, upper: float, name: str) -> None:
    if not np.isfinite(value) or value < lower or value > upper:
        raise ValueError(f"{name} must be finite and in [{lower}, {upper}]")


@dataclass(frozen=True)
class RealtimeFeedbackConfig:
    """Configuration for live-shot synchronisation feedback (reference :36-64)."""

    target_r: float = 0.75
    deadband: float = 0.04
    base_dt: float = 0.08
    trotter_steps: int = 3
    measurement_shots: int = 128
    base_gain: float = 0.8
    max_gain: float = 2.0
    monitor_strength: float = 0.18
    correction_angle: float = 0.12
    feedback_correction_sign: float = -1.0

    def __post_init__(self) -> None:
        _require_range(self.target_r, 0.0, 1.0, "target_r")
        _require_range(self.deadband, 0.0, 1.0, "deadband")
        _require_positive(self.base_dt, "base_dt")
        _require_positive(self.base_gain, "base_gain", allow_zero=True)
        _require_range(self.monitor_strength, 0.0, np.pi, "monitor_strength")
        _require_range(self.correction_angle, 0.0, np.pi, "correction_angle")
        if self.feedback_correction_sign not in {-1.0, 1.0}:
            raise ValueError("feedback_correction_sign must be either -1.0 or 1.0")
        if not isinstance(self.trotter_steps, int) or self.trotter_steps < 1:
            raise ValueError("trotter_steps must be a positive integer")
        if not isinstance(self.measurement_shots, int) or self.measurement_shots < 1:
            raise ValueError("measurement_shots must be a positive integer")
        if not np.isfinite(self.max_gain) or self.max_gain < 1.0:
            raise ValueError("max_gain must be finite and at least 1.0")


@dataclass(frozen=True)
class FeedbackStep:
    """One closed-loop feedback update (reference :67-80)."""

    index: int
    action: str
    r_live: float
    r_statevector: float
    psi_statevector: float
    error: float
    applied_coupling_scale: float
    next_coupling_scale: float
    correction_angle: float
    readout_counts: Mapping[str, int] = field(default_factory=dict)


def feedback_policy_numpy(r_values, target_r: float, deadband: float, base_gain: float, max_gain: float):
    """(actions, gains, errors) of the reference policy (:316-341)."""
    r = np.asarray(r_values, dtype=np.float64)
    if not np.all(np.isfinite(r)):
        raise ValueError("r_values must contain only finite values")
    _require_range(target_r, 0.0, 1.0, "target_r")
    _require_range(deadband, 0.0, 1.0, "deadband")
    _require_positive(base_gain, "base_gain", allow_zero=True)
    if not np.isfinite(max_gain) or max_gain < 1.0:
        raise ValueError("max_gain must be finite and at least 1.0")
    errors = np.asarray(target_r - r, dtype=np.float64)
    raw = 1.0 + base_gain * errors
    below, above = errors > deadband, errors < -deadband  # too little / too much synchrony
    actions = below.astype(np.int32) - above.astype(np.int32)
    gains = np.where(below, np.minimum(raw, max_gain), np.where(above, np.clip(raw, 1.0 / max_gain, 1.0), 1.0))
    return actions, gains.astype(np.float64), errors


def _sample_axis(expectations: np.ndarray, shots: int, rng: np.random.Generator) -> np.ndarray:
    probabilities = np.clip((1.0 + expectations) / 2.0, 0.0, 1.0)
    plus_counts = rng.binomial(shots, probabilities)
    return (2.0 * plus_counts.astype(np.float64) / shots) - 1.0


class RealtimeSyncFeedbackController:
    """Stateful live-shot feedback controller; the statevector lives on the GPU between ticks."""

    def __init__(self, K_coupling, omega_natural, config: RealtimeFeedbackConfig | None = None, trotter_order: int = 1,
                 *, evolution: str = "trotter", **solver_kwargs):
        self.config = config or RealtimeFeedbackConfig()
        self.K = np.asarray(K_coupling, dtype=np.float64)
        self.omega = np.asarray(omega_natural, dtype=np.float64)
        self.n = int(self.omega.shape[0])
        self.trotter_order = trotter_order
        # evolution="trotter": the product formula the reference's API describes; "reference": exact per tick
        self._solver = QuantumKuramotoSolver(self.n, self.K, self.omega, trotter_order=trotter_order, evolution=evolution,
                                             **solver_kwargs)
        self._solver.prepare_initial_state()
        self._coupling_scale = 1.0
        self.history: list[FeedbackStep] = []

    @property
    def statevector(self) -> np.ndarray:
        """Host copy of the current controller state (canonical order)."""
        return self._solver.statevector()

    @property
    def coupling_scale(self) -> float:
        return self._coupling_scale

    def set_coupling_scale(self, scale: float) -> None:
        _require_range(scale, 1.0 / self.config.max_gain, self.config.max_gain, "scale")
        self._coupling_scale = float(scale)

    def reset(self) -> None:
        self._solver.update_problem(self.K)
        self._solver.prepare_initial_state()
        self._coupling_scale = 1.0
        self.history.clear()

    def close(self) -> None:
        self._solver.close()

    def step(self, seed: int | None = None) -> FeedbackStep:
        """One tick: evolve with K * scale on the live state, measure, decide, correct (reference :112-166)."""
        cfg = self.config
        rng = np.random.default_rng(seed)
        applied_scale = self._coupling_scale
        self._solver.update_problem(self.K * applied_scale)
        self._solver.evolve(cfg.base_dt, cfg.trotter_steps)
        eng = self._solver.engine()
        exp_x, exp_y = eng.xy_expectations()
        z = complex(float(exp_x.sum()), float(exp_y.sum())) / self.n
        r_exact, psi_exact = float(abs(z)), float(np.angle(z))
        x_samples = _sample_axis(exp_x, cfg.measurement_shots, rng)
        y_samples = _sample_axis(exp_y, cfg.measurement_shots, rng)
        r_live = float(abs(complex(float(np.sum(x_samples)), float(np.sum(y_samples))) / self.n))
        action_code, next_scale, error = feedback_policy_numpy(np.array([r_live]), cfg.target_r, cfg.deadband, cfg.base_gain,
                                                               cfg.max_gain)
        action = _ACTION_LABELS[int(action_code[0])]
        correction = self._apply_feedback_correction(float(error[0]), action)
        counts = self._sample_readout_counts(rng)
        self._coupling_scale = float(next_scale[0])
        step = FeedbackStep(index=len(self.history), action=action, r_live=r_live, r_statevector=r_exact,
                            psi_statevector=psi_exact, error=float(error[0]), applied_coupling_scale=float(applied_scale),
                            next_coupling_scale=self._coupling_scale, correction_angle=correction, readout_counts=counts)
        self.history.append(step)
        return step

    def run(self, n_steps: int, seed: int | None = None) -> list[FeedbackStep]:
        if not isinstance(n_steps, int) or n_steps < 1:
            raise ValueError("n_steps must be a positive integer")
        rng = np.random.default_rng(seed)
        return [self.step(seed=int(rng.integers(0, 2**32 - 1))) for _ in range(n_steps)]

    def _apply_feedback_correction(self, error: float, action: str) -> float:
        cfg = self.config
        if action == "hold":
            return 0.0
        correction = float(np.clip(error * cfg.correction_angle, -cfg.correction_angle, cfg.correction_angle))
        if abs(correction) < 1e-15:
            return 0.0
        correction *= -cfg.feedback_correction_sign
        direction = np.where(np.cos(self.omega) >= 0.0, 1.0, -1.0)
        self._solver.engine().apply_ry(correction * direction)
        return correction

    def _sample_readout_counts(self, rng: np.random.Generator) -> dict[str, int]:
        """Computational-basis read-out sample (reference :421-432); above 24 qubits the 2^n probabilities do not
        go to the host and the sample is skipped (empty mapping) -- the draw is the tick's LAST use of the generator,
        so the decisions stay those of the reference."""
        if self.n > _READOUT_MAX_QUBITS:
            return {}
        probabilities = np.abs(self._solver.statevector()) ** 2
        draws = rng.multinomial(self.config.measurement_shots, probabilities / probabilities.sum())
        return {format(index, f"0{self.n}b"): int(count) for index, count in enumerate(draws) if int(count) > 0}
