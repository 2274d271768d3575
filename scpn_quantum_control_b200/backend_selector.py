"""Backend selector slot for the GPU statevector route.

The reference's selector already names a ``"gpu_statevector"`` backend
(reference: src/scpn_quantum_control/phase/backend_selector.py:161-167) but never forwards
``has_gpu`` (:212-218) and has no branch for the label -- everything falls through to the
Qiskit solver (:288-293).  This module provides that missing branch: ``recommend_backend``
returns the same recommendation dictionaries for the statevector-class labels, and
``auto_solve`` routes ``gpu_statevector`` to :class:`QuantumKuramotoSolver` of this package
with the reference's call shape ``Solver(n, K, omega).run(t_max=, dt=, max_statevector_gib=)``
and result envelope ``{"backend_used", "result", "recommendation"}``.

Only the statevector-class decisions are made here; the reference's other backends (exact
diagonalisation, sector ED, MPS, Lindblad, hardware) stay in the reference -- see
INTEGRATION.md for the two-line patch that wires this function into its ``auto_solve``.
"""

from __future__ import annotations

from typing import Any

import numpy as np

from .engine import device_count, device_memory
from .xy_kuramoto import QuantumKuramotoSolver

# local qubits per GPU the fused tile-pass engine takes (32-bit amplitude offsets inside a shard; 2 x 32 GiB ping-pong
# buffers of the 180 GB).  n_local = 32 would still fit the memory but runs gate by gate, so it is not recommended.
GPU_STATEVECTOR_MAX_QUBITS_1GPU = 31


def has_gpu() -> bool:
    try:
        return device_count() > 0
    except Exception:
        return False


def recommend_backend(n: int, ram_gb: float = 32.0, has_gpu: bool | None = None, n_gpus: int = 1) -> dict[str, Any]:
    """Statevector-class recommendation (reference :52-183, branches :161-183).

    ``gpu_statevector`` when a GPU is present and the (sharded) state fits; else ``statevector``
    if it fits host RAM / 2; else ``hardware``.  The reference caps the label at n <= 30 for one
    GPU (:161); with 180 GB of HBM3e and ``n_gpus`` devices the cap is lifted to what fits.
    """
    if isinstance(n, bool) or not isinstance(n, int) or n < 1:
        raise ValueError("n must be a positive integer qubit count.")
    ram_gb = float(ram_gb)
    if not np.isfinite(ram_gb) or ram_gb <= 0.0:
        raise ValueError("ram_gb must be finite and positive.")
    if has_gpu is None:
        has_gpu = globals()["has_gpu"]()
    sv_mb = (2**n) * 16 / 1e6
    gbits = max(0, int(n_gpus).bit_length() - 1)
    if has_gpu and n - gbits <= GPU_STATEVECTOR_MAX_QUBITS_1GPU:
        return {
            "backend": "gpu_statevector",
            "reason": f"GPU statevector: {sv_mb:.0f} MB on device" + (f" over {n_gpus} GPUs" if n_gpus > 1 else ""),
            "memory_mb": sv_mb,
            "feasible": True,
        }
    if sv_mb < ram_gb * 1000 * 0.5:
        return {"backend": "statevector", "reason": f"CPU statevector: {sv_mb:.0f} MB", "memory_mb": sv_mb, "feasible": True}
    return {
        "backend": "hardware",
        "reason": f"n={n} too large for classical sim",
        "memory_mb": 0,
        "feasible": True,
        "note": "Recommendation only; submit with AsyncHardwareRunner for QPU execution.",
    }


def auto_solve(
    K,
    omega,
    ram_gb: float = 32.0,
    t_max: float = 1.0,
    dt: float = 0.1,
    *,
    max_dense_gib: float | None = None,
    evolution: str = "reference",
    **solver_kwargs: Any,
) -> dict[str, Any]:
    """Route the statevector path to the B200 engine (reference :186-293, fall-through :288-293)."""
    K = np.asarray(K)
    n = K.shape[0]
    rec = recommend_backend(n, ram_gb=ram_gb)
    if rec["backend"] != "gpu_statevector":
        raise RuntimeError(
            f"auto_solve (B200 backend) selected {rec['backend']!r}: no CUDA device is visible or the state does "
            "not fit. This backend has no CPU fallback; use the reference's CPU route instead."
        )
    solver = QuantumKuramotoSolver(n, K, omega, evolution=evolution, **solver_kwargs)
    try:
        trajectory = solver.run(t_max=t_max, dt=dt, max_statevector_gib=max_dense_gib)
    finally:
        solver.close()
    return {"backend_used": "gpu_statevector", "result": trajectory, "recommendation": rec}
