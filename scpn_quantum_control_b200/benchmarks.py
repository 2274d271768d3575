"""The reference's callers of the statevector Trotter path, re-pointed at the GPU engine.

Each function keeps the name, arguments and result keys of the call site it stands in for, so that a maintainer
swaps the import and the n <= 16 cap of the reference (``MAX_STATEVECTOR_OSCILLATORS``,
benchmarks/reproducible_comparison.py:51) becomes the device-memory budget:

* ``quantum_benchmark``            benchmarks/quantum_advantage.py:138-175
* ``quantum_trotter_row``          benchmarks/reproducible_comparison.py:340-342,366-373 (the ``quantum_trotter`` row)
* ``simulate_execute``             studio/executive_simulate.py:181-233 (SimulateExecutive.execute)
* ``statevector_scaling_row``      scripts/bench_s2_scaling_lite.py:254-281 (_aer_statevector_row)

The classical rows of those artefacts (SciPy ODE, dense expm) stay in the reference; nothing here computes on the CPU.
"""

from __future__ import annotations

import platform
import time
from dataclasses import dataclass
from typing import Any

import numpy as np

from .dense_budget import statevector_qubit_budget
from .hamiltonian import OMEGA_N_16, build_knm_paper27, omega_for_oscillators
from .xy_kuramoto import QuantumKuramotoSolver


def _paper27_problem(n: int, K=None, omega=None) -> tuple[np.ndarray, np.ndarray]:
    coupling = build_knm_paper27(L=n) if K is None else np.asarray(K, dtype=np.float64)
    frequencies = omega_for_oscillators(n) if omega is None else np.asarray(omega, dtype=np.float64)
    return coupling, frequencies


def quantum_benchmark(n: int, t_max: float = 1.0, dt: float = 0.1, trotter_reps: int = 5, *, K=None, omega=None,
                      device: int = 0) -> dict[str, float]:
    """Time Lie-Trotter evolution of the Paper-27 problem on the statevector engine: ``max(1, int(t_max / dt))``
    steps of ``trotter_reps`` first-order layers from the omega product state, no observable (reference :138-175).
    The timed region starts before state preparation and ends after the device has finished, as the reference's
    does around ``Statevector.from_instruction`` ... ``sv.evolve``."""
    coupling, frequencies = _paper27_problem(n, K, omega)
    n_steps = max(1, int(t_max / dt))
    solver = QuantumKuramotoSolver(n, coupling, frequencies, trotter_order=1, evolution="trotter", device=device)
    try:
        eng = solver.engine()  # allocation and schedule compilation are not part of the reference's timed region either
        eng.synchronize()
        t0 = time.perf_counter()
        solver.prepare_initial_state()
        for _ in range(n_steps):
            solver.evolve(dt, trotter_reps)
        eng.synchronize()
        t_total = time.perf_counter() - t0
    finally:
        solver.close()
    return {"t_total_ms": t_total * 1000.0, "n_trotter_steps": n_steps * trotter_reps}


@dataclass(frozen=True)
class ComparisonMethodRow:
    """One method's row of the classical-vs-quantum comparison artefact (reference :84-126)."""

    method: str
    backend: str
    available: bool
    r_final: float | None
    r_error_vs_exact: float | None
    elapsed_ms: float
    unavailable_reason: str | None = None

    def to_dict(self) -> dict[str, Any]:
        return {k: getattr(self, k) for k in ("method", "backend", "available", "r_final", "r_error_vs_exact",
                                              "elapsed_ms", "unavailable_reason")}


def quantum_trotter_row(n_oscillators: int, t_max: float = 1.0, dt: float = 0.1, trotter_per_step: int = 5, *, K=None,
                        omega=None, r_exact: float | None = None, evolution: str = "reference",
                        device: int = 0) -> ComparisonMethodRow:
    """The ``quantum_trotter`` row of ``run_reproducible_kuramoto_comparison`` (reference :340-342,366-373) from the GPU
    solver.  The reference marks the row unavailable above 16 oscillators; here the bound is the device-memory budget
    of the statevector.  ``r_exact`` (the artefact's ``classical_exact`` R) fills ``r_error_vs_exact`` when given."""
    if not isinstance(n_oscillators, int) or n_oscillators < 2:
        raise ValueError("n_oscillators must be an integer >= 2")
    if not np.isfinite(t_max) or t_max <= 0.0:
        raise ValueError("t_max must be finite and positive")
    if not np.isfinite(dt) or dt <= 0.0 or dt > t_max:
        raise ValueError("dt must be finite, positive and not exceed t_max")
    if not isinstance(trotter_per_step, int) or trotter_per_step < 1:
        raise ValueError("trotter_per_step must be an integer >= 1")
    backend = "GPU statevector Trotter (QuantumKuramotoSolver, B200)"
    limit = statevector_qubit_budget(device=device)
    if n_oscillators > limit:
        return ComparisonMethodRow("quantum_trotter", backend, False, None, None, 0.0,
                                   f"n_oscillators>{limit}: the statevector does not fit the device-memory budget")
    coupling, frequencies = _paper27_problem(n_oscillators, K, omega)
    solver = QuantumKuramotoSolver(n_oscillators, coupling, frequencies, evolution=evolution, device=device)
    try:
        start = time.perf_counter()
        quantum = solver.run(t_max=t_max, dt=dt, trotter_per_step=trotter_per_step)
        elapsed_ms = (time.perf_counter() - start) * 1000.0
    finally:
        solver.close()
    r_quantum = float(quantum["R"][-1])
    return ComparisonMethodRow("quantum_trotter", backend, True, r_quantum,
                               None if r_exact is None else abs(r_quantum - float(r_exact)), elapsed_ms)


def simulate_execute(simulate_spec: dict[str, Any], *, backend: str = "gpu_statevector", device: int = 0) -> dict[str, Any]:
    """Outputs mapping of ``SimulateExecutive.execute`` (reference :181-233) for a plan's parameters
    ``{K_nm, omega, t_max, dt, trotter_per_step, trotter_order}``."""
    k_nm = np.asarray(simulate_spec["K_nm"], dtype=np.float64)
    omega = np.asarray(simulate_spec["omega"], dtype=np.float64)
    solver = QuantumKuramotoSolver(len(simulate_spec["K_nm"]), k_nm, omega, trotter_order=simulate_spec["trotter_order"],
                                   device=device)
    try:
        trajectory = solver.run(simulate_spec["t_max"], simulate_spec["dt"], simulate_spec["trotter_per_step"])
    finally:
        solver.close()
    order_parameter = np.asarray(trajectory.R, dtype=np.float64)
    times = np.asarray(trajectory.times, dtype=np.float64)
    return {
        "backend": backend,
        "n_nodes": len(simulate_spec["K_nm"]),
        "t_max": simulate_spec["t_max"],
        "dt": simulate_spec["dt"],
        "trotter_per_step": simulate_spec["trotter_per_step"],
        "trotter_order": simulate_spec["trotter_order"],
        # the reference's summary keys (:210-224)
        "evolution_schema": "studio.quantum-evolution.v1",   # studio/verbs.py:47 QUANTUM_EVOLUTION_SCHEMA
        "n_points": int(times.shape[0]),
        "order_parameter_initial": float(order_parameter[0]),
        "order_parameter_final": float(order_parameter[-1]),
        "order_parameter_max": float(order_parameter.max()),
        "order_parameter_mean": float(order_parameter.mean()),
        "order_parameter_delta": float(order_parameter[-1] - order_parameter[0]),
        # plus the trajectory itself
        "n_time_points": int(times.size),
        "times": times.tolist(),
        "R": order_parameter.tolist(),
        "R_initial": float(order_parameter[0]),
        "R_final": float(order_parameter[-1]),
        "R_mean": float(order_parameter.mean()),
        "R_max": float(order_parameter.max()),
        "R_min": float(order_parameter.min()),
    }


def statevector_scaling_row(n_qubits: int, max_statevector_qubits: int | None = None, *, device: int = 0) -> dict[str, Any]:
    """Row of the S2 scaling campaign's statevector baseline (reference scripts/bench_s2_scaling_lite.py:254-281):
    |0...0> evolved by ``evolve(0.1, trotter_steps=1)`` of the Paper-27 problem (omega = linspace(0.1, 1, n) above 16
    qubits, :121-126), then (R, psi).  ``max_statevector_qubits`` defaults to the device-memory budget."""
    limit = statevector_qubit_budget(device=device) if max_statevector_qubits is None else max_statevector_qubits
    # the measurement's share of the reference's row schema (_base_row, :93-108); protocol id, command, dependency versions and
    # git commit are provenance fields the campaign script adds around it
    row: dict[str, Any] = {"n_qubits": n_qubits, "baseline": "gpu_statevector", "status": "ok", "wall_time_ms": None,
                           "memory_bytes": None, "metric_payload": {}, "machine": platform.platform(), "notes": []}
    if n_qubits > limit:
        row.update(status="skipped", notes=["size gate"])
        return row
    k_matrix = build_knm_paper27(n_qubits)
    omega = OMEGA_N_16[:n_qubits].copy() if n_qubits <= OMEGA_N_16.size else np.linspace(0.1, 1.0, n_qubits)
    solver = QuantumKuramotoSolver(n_qubits, k_matrix, omega, device=device)
    try:
        solver.engine().synchronize()
        start = time.perf_counter()
        solver.prepare_product_state(np.zeros(n_qubits))
        solver.evolve(0.1, trotter_steps=1)
        r_value, psi_value = solver.measure_order_parameter()
        elapsed_ms = (time.perf_counter() - start) * 1000.0
        stats = solver.engine().stats()
    finally:
        solver.close()
    row["wall_time_ms"] = elapsed_ms
    row["memory_bytes"] = int(2**n_qubits) * 16 * (2 if n_qubits > 12 else 1)
    row["metric_payload"] = {
        "trotter_steps": 1,
        "final_R": float(r_value),
        "final_psi": float(psi_value),
        "hilbert_dim": int(2**n_qubits),
        "estimated_statevector_bytes": int(2**n_qubits) * 16,
        "state_sweeps": int(stats.get("state_sweeps", 0)),
    }
    return row
