"""Stateful UPDE wrapper on the GPU engine -- mirror of the reference's ``QuantumUPDESolver``
(reference: src/scpn_quantum_control/phase/trotter_upde.py:44-111): keeps a live device
statevector across ``step()`` calls, ``run()`` forwards to ``QuantumKuramotoSolver.run``."""

from __future__ import annotations

import numpy as np

from .hamiltonian import OMEGA_N_16, build_knm_paper27
from .xy_kuramoto import QuantumKuramotoSolver


class QuantumUPDESolver:
    """Full 16-layer UPDE as quantum spin chain (defaults: canonical Paper-27 parameters)."""

    def __init__(self, K=None, omega=None, trotter_order: int = 1, **solver_kwargs):
        if K is None:
            K = build_knm_paper27()
        if omega is None:
            omega = OMEGA_N_16.copy()
        self.K = np.asarray(K, dtype=np.float64)
        self.omega = np.asarray(omega, dtype=np.float64)
        self.n_layers = len(self.omega)
        self._solver = QuantumKuramotoSolver(self.n_layers, self.K, self.omega, trotter_order=trotter_order,
                                             **solver_kwargs)
        self._live = False

    def step(self, dt: float = 0.1, trotter_steps: int = 5) -> dict:
        """Single interval on the live state; returns {"R_global", "psi", "dt"} (reference :75-89)."""
        if not self._live:
            self._solver.prepare_initial_state()
            self._live = True
        self._solver.evolve(dt, trotter_steps)
        R, psi = self._solver.measure_order_parameter()
        return {"R_global": R, "psi": psi, "dt": dt}

    def run(self, n_steps: int = 50, dt: float = 0.1, trotter_per_step: int = 5) -> dict:
        """Full trajectory returning R(t) over n_steps (reference :91-103)."""
        result = self._solver.run(n_steps * dt, dt, trotter_per_step)
        return {"times": result["times"], "R": result["R"], "n_layers": self.n_layers}

    def reset(self) -> None:
        """Next step() re-initialises from omega (reference :105-107)."""
        self._live = False

    def set_coupling(self, K) -> None:
        """Re-parameterise K on the live state without re-allocating (realtime feedback ticks,
        reference: control/realtime_feedback.py:98-137)."""
        self._solver.update_problem(K)
        self.K = self._solver.K.copy()

    def statevector(self) -> np.ndarray:
        return self._solver.statevector()

    def hamiltonian(self):
        return self._solver.build_hamiltonian()

    def close(self) -> None:
        self._solver.close()
