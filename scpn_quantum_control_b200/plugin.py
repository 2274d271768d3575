"""Registry adapter: a runner class with the ``(K, omega, **kwargs)`` constructor the
reference's ``PluginRegistry`` expects (reference: src/scpn_quantum_control/hardware/
plugin_registry.py:55-86,108-144; documented method ``run_trotter(t, reps)`` :23-27)."""

from __future__ import annotations

import numpy as np

from .xy_kuramoto import QuantumKuramotoSolver


class B200StatevectorRunner:
    """``registry.register_class("b200_statevector", B200StatevectorRunner)``"""

    def __init__(self, K, omega, **kwargs):
        K = np.asarray(K)
        self.solver = QuantumKuramotoSolver(int(K.shape[0]), K, omega, **kwargs)

    def run(self, t_max: float, dt: float, trotter_per_step: int | None = None, **kw):
        return self.solver.run(t_max, dt, trotter_per_step, **kw)

    def run_trotter(self, t: float, reps: int):
        """Evolve the initial product state for time ``t`` with ``reps`` Trotter repetitions;
        returns the final statevector (canonical order) and its (R, psi)."""
        self.solver.prepare_initial_state()
        self.solver.evolve(float(t), int(reps))
        R, psi = self.solver.measure_order_parameter()
        return {"statevector": self.solver.statevector(), "R": R, "psi": psi}

    def close(self) -> None:
        self.solver.close()


def register_with(registry, name: str = "b200_statevector") -> None:
    """Register the runner with a reference ``PluginRegistry`` instance."""
    registry.register_class(name, B200StatevectorRunner)
