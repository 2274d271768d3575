"""B200-native exact-statevector engine for the Kuramoto-XY simulator route of
anulum/scpn-quantum-control: ``(K_nm, omega, t, dt) -> state / R(t)`` behind the reference's
``QuantumKuramotoSolver`` / ``auto_solve`` entry points.  Host code is Python; all arithmetic is
hand-written sm_100a CUDA behind the C ABI in ``include/xyk.h`` (``libxyk.so``)."""

from .dense_budget import DenseAllocationError, require_statevector_allocation
from .hamiltonian import OMEGA_N_16, build_knm_paper27, build_kuramoto_ring, omega_for_oscillators, xy_terms
from .results import TrajectoryResult
from .xy_kuramoto import QuantumKuramotoSolver, TrotterEvolutionConfig, time_grid
from .backend_selector import auto_solve, recommend_backend
from .trotter_upde import QuantumUPDESolver
from .plugin import B200StatevectorRunner, register_with

__all__ = [
    "DenseAllocationError",
    "require_statevector_allocation",
    "OMEGA_N_16",
    "build_knm_paper27",
    "build_kuramoto_ring",
    "omega_for_oscillators",
    "xy_terms",
    "TrajectoryResult",
    "QuantumKuramotoSolver",
    "TrotterEvolutionConfig",
    "time_grid",
    "auto_solve",
    "recommend_backend",
    "QuantumUPDESolver",
    "B200StatevectorRunner",
    "register_with",
]
