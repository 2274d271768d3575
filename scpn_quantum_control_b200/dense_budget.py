"""Fail-closed memory guard for the device statevector.

Mirrors the reference's ``require_dense_allocation`` contract
(reference: src/scpn_quantum_control/dense_budget.py:141-172; call site
phase/xy_kuramoto.py:212-217): ``DenseAllocationError(MemoryError)`` is raised *before* any
allocation when the request exceeds the active budget.  The reference budgets host RAM
(min(8 GiB, 0.3 * available)); a GPU backend budgets DEVICE memory instead: the default budget
is what the caller passes (``max_statevector_gib``), else ``SCPN_MAX_DENSE_GIB``, else 90 % of
the device's free memory (per GPU).
"""

from __future__ import annotations

import os
from dataclasses import dataclass

DEFAULT_DENSE_BUDGET_ENV = "SCPN_MAX_DENSE_GIB"
DEVICE_FRACTION = 0.90
GIB = 1024**3


class DenseAllocationError(MemoryError):
    """Raised before an unsafe dense Hilbert-space allocation is attempted."""


@dataclass(frozen=True)
class DenseAllocationEstimate:
    n_qubits: int
    dimension: int
    dtype: str
    bytes_required: int
    budget_bytes: int
    label: str
    n_devices: int = 1

    @property
    def gib_required(self) -> float:
        return self.bytes_required / GIB

    @property
    def budget_gib(self) -> float:
        return self.budget_bytes / GIB


def budget_bytes(max_gib: float | None, device_free_bytes: int | None) -> int:
    """Per-device budget in bytes (explicit kwarg > environment > device free memory)."""
    if max_gib is not None:
        if max_gib <= 0:
            raise ValueError("max_gib must be positive")
        return int(max_gib * GIB)
    env = os.environ.get(DEFAULT_DENSE_BUDGET_ENV)
    if env:
        try:
            parsed = float(env)
        except ValueError as exc:
            raise ValueError(f"{DEFAULT_DENSE_BUDGET_ENV} must be a positive number") from exc
        if parsed <= 0:
            raise ValueError(f"{DEFAULT_DENSE_BUDGET_ENV} must be positive")
        return int(parsed * GIB)
    if device_free_bytes is None:
        return int(8.0 * GIB)
    return int(device_free_bytes * DEVICE_FRACTION)


def require_statevector_allocation(
    n_qubits: int,
    *,
    itemsize: int = 16,
    buffers: int = 1,
    n_devices: int = 1,
    max_gib: float | None = None,
    device_free_bytes: int | None = None,
    label: str = "Kuramoto statevector trajectory",
) -> DenseAllocationEstimate:
    """Raise ``DenseAllocationError`` if ``buffers`` copies of the (sharded) state exceed the budget."""
    if not isinstance(n_qubits, int):
        raise TypeError("n_qubits must be an integer")
    if n_qubits < 1:
        raise ValueError("n_qubits must be >= 1")
    dim = 1 << n_qubits
    required = dim * itemsize * buffers // max(1, n_devices)
    est = DenseAllocationEstimate(
        n_qubits=n_qubits,
        dimension=dim,
        dtype="complex128" if itemsize == 16 else "complex64",
        bytes_required=required,
        budget_bytes=budget_bytes(max_gib, device_free_bytes),
        label=label,
        n_devices=n_devices,
    )
    if est.bytes_required > est.budget_bytes:
        raise DenseAllocationError(
            f"{label} for n={n_qubits} requires {est.gib_required:.2f} GiB per device for shape "
            f"({dim},) ({est.dtype}, {buffers} buffer(s), {n_devices} device(s)), above the active dense "
            f"budget {est.budget_gib:.2f} GiB. Use more GPUs, complex64, or a sparse/tensor-network "
            f"backend instead of dense allocation."
        )
    return est


def statevector_qubit_budget(*, itemsize: int = 16, buffers: int = 2, device: int = 0, max_gib: float | None = None) -> int:
    """Largest n whose ``buffers`` state copies fit the active per-device budget (the GPU analogue of the reference's
    fixed ``MAX_STATEVECTOR_OSCILLATORS = 16``, benchmarks/reproducible_comparison.py:51)."""
    from .engine import device_memory

    try:
        free, _total = device_memory(device)
    except RuntimeError:
        free = None
    budget = budget_bytes(max_gib, free)
    n = 1
    while (1 << (n + 1)) * itemsize * buffers <= budget:
        n += 1
    return n
