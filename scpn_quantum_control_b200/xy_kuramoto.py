"""GPU drop-in for the reference's ``QuantumKuramotoSolver`` (statevector route).

Reference interface mirrored here: src/scpn_quantum_control/phase/xy_kuramoto.py
  * ``TrotterEvolutionConfig``                      :54-77
  * ``QuantumKuramotoSolver.__init__`` + validation :86-125  (same messages)
  * ``run(t_max, dt, trotter_per_step, *, max_statevector_gib)`` -> ``TrajectoryResult`` :187-258
  * ``measure_order_parameter`` :156-185, ``energy_expectation`` :260-264

All arithmetic runs in hand-written sm_100a kernels behind the C ABI of ``libxyk.so``; Qiskit is
not involved.  There is no CPU fallback: without the CUDA library / a GPU the solver raises.

Two numerics modes (SURVEY.md section 0 explains why both exist):

``evolution="reference"`` (default)
    reproduces what the reference's ``run()`` numerically returns under its pinned Qiskit: the
    exact propagation ``exp(-i H step_dt)`` per grid interval, independent of
    ``trotter_per_step`` / ``trotter_order``.  The engine reaches it with a 6th-order Suzuki
    composition of its pair-rotation layers whose repetition count is calibrated a posteriori
    (halving check on the observables) to ``exact_tol``.
``evolution="trotter"``
    the literal product formula the reference's circuit defines: ``trotter_per_step``
    repetitions per interval of LieTrotter (``trotter_order=1``) or SuzukiTrotter order 2
    (``trotter_order=2``), terms in the reference's order.
"""

from __future__ import annotations

from dataclasses import dataclass, replace

import numpy as np

from . import _lib
from .dense_budget import require_statevector_allocation
from .engine import XYKEngine, batch_run, device_memory
from .results import TrajectoryResult

_SMALL_N = 12  # whole state in shared memory: the run loop is a single kernel launch


def _as_real_numeric_array(name: str, values: object) -> np.ndarray:
    """float64 copy; bool / str / object / complex inputs are rejected (reference :37-51)."""
    try:
        raw = np.asarray(values)
    except ValueError as exc:
        raise ValueError(f"{name} must be a rectangular numeric array") from exc
    if raw.dtype.kind in "bOSUc":
        raise ValueError(f"{name} must contain real numeric scalars")
    try:
        return np.array(raw, dtype=np.float64, copy=True)
    except (TypeError, ValueError) as exc:
        raise ValueError(f"{name} must contain real numeric scalars") from exc


def time_grid(t_max: float, dt: float) -> tuple[np.ndarray, np.ndarray]:
    """Grid 0, dt, 2dt, ..., t_max with a possibly shorter last interval (reference :221-233)."""
    tol = max(np.finfo(float).eps * max(1.0, abs(t_max), abs(dt)) * 16.0, 1e-15)
    pts = [0.0]
    t = 0.0
    while t + dt < t_max - tol:
        t += dt
        pts.append(t)
    if t_max > pts[-1] + tol:
        pts.append(float(t_max))
    else:
        pts[-1] = float(t_max)
    times = np.asarray(pts, dtype=float)
    return times, np.diff(times)


@dataclass(frozen=True)
class TrotterEvolutionConfig:
    """Typed defaults for Kuramoto-XY Trotter evolution (reference :54-77)."""

    order: int = 1
    evolve_steps: int = 10
    run_steps_per_step: int = 5

    def __post_init__(self) -> None:
        if self.order not in (1, 2):
            raise ValueError(f"order must be 1 or 2, got {self.order}")
        if not isinstance(self.evolve_steps, int) or self.evolve_steps < 1:
            raise ValueError(f"evolve_steps must be a positive integer, got {self.evolve_steps}")
        if not isinstance(self.run_steps_per_step, int) or self.run_steps_per_step < 1:
            raise ValueError(
                f"run_steps_per_step must be a positive integer, got {self.run_steps_per_step}"
            )


_CALIBRATE_DIRECT_MAX_N = 26   # halving check on the full problem up to here (1 GiB state; an S6 repetition costs < 1 s)
_CALIBRATE_PROXY_N = 18        # size of the reduced problem used above that


def halving_reps(observe, tol_step: float, max_reps: int = 256) -> int:
    """A-posteriori choice of the repetition count: ``observe(reps)`` returns the observables after one
    interval run with ``reps`` repetitions; double ``reps`` until two consecutive results agree to
    ``tol_step`` and return the coarser of the two (a 6th-order formula: error(reps/2) ~ 64 error(reps))."""
    reps, prev = 1, None
    while reps <= max_reps:
        cur = np.asarray(observe(reps))
        if prev is not None and np.max(np.abs(cur - prev)) <= tol_step:
            return max(1, reps // 2)
        prev = cur
        reps *= 2
    return max_reps


def reduced_problem_for_calibration(K: np.ndarray, omega: np.ndarray, n_sub: int) -> tuple[np.ndarray, np.ndarray]:
    """The ``n_sub`` qubits with the largest local energy scale sum_j |K_ij| + |omega_i|, their coupling
    block scaled up so that its largest row sum equals the full problem's (the splitting error is driven
    by the local energy scale times the step)."""
    K = np.asarray(K, dtype=np.float64)
    omega = np.asarray(omega, dtype=np.float64)
    n = K.shape[0]
    if n_sub >= n:
        return K.copy(), omega.copy()
    row = np.abs(K).sum(axis=1)
    idx = np.sort(np.argsort(-(row + np.abs(omega)), kind="stable")[:n_sub])
    Ks = K[np.ix_(idx, idx)].copy()
    sub_row = np.abs(Ks).sum(axis=1).max()
    if sub_row > 0.0:
        Ks *= max(1.0, row.max() / sub_row)
    return Ks, omega[idx].copy()


class QuantumKuramotoSolver:
    """Exact-statevector simulation of Kuramoto oscillators on one B200 (or a sharded group)."""

    def __init__(
        self,
        n_oscillators: int,
        K_coupling,
        omega_natural,
        trotter_order: int | None = None,
        evolution_config: TrotterEvolutionConfig | None = None,
        *,
        evolution: str = "reference",
        dtype: str = "complex128",
        device: int = 0,
        exact_tol: float = 2e-10,
        distributed: bool = False,
        anisotropy_delta: float = 0.0,
    ):
        """K_coupling: (n,n) coupling matrix, omega_natural: (n,) frequencies.

        ``anisotropy_delta`` != 0 evolves the XXZ Hamiltonian of ``knm_to_xxz_hamiltonian(K, omega, delta)``
        (reference bridge/knm_hamiltonian.py:141-206) instead of the XY model; ``evolution="trotter"`` then runs gate by
        gate in the reference's term order (XX, YY, ZZ per pair), ``evolution="reference"`` through the tiled engine."""
        self.n = self._validate_n_oscillators(n_oscillators)
        self.K = _as_real_numeric_array("K_coupling", K_coupling)
        self.omega = _as_real_numeric_array("omega_natural", omega_natural)
        self._validate_coupling_inputs(self.n, self.K, self.omega)
        np.fill_diagonal(self.K, 0.0)
        config = evolution_config or TrotterEvolutionConfig()
        order = config.order if trotter_order is None else trotter_order
        if order not in (1, 2):
            raise ValueError(f"trotter_order must be 1 or 2, got {order}")
        self.evolution_config = replace(config, order=order)
        self.trotter_order = self.evolution_config.order
        if evolution not in ("reference", "trotter", "chebyshev"):
            raise ValueError(f"evolution must be 'reference', 'trotter' or 'chebyshev', got {evolution!r}")
        if evolution == "chebyshev" and (dtype != "complex128" or distributed):
            raise ValueError("evolution='chebyshev' needs a single-device complex128 state")
        if dtype not in ("complex128", "complex64"):
            raise ValueError(f"dtype must be 'complex128' or 'complex64', got {dtype!r}")
        self.evolution = evolution
        self.dtype = dtype
        self.device = int(device)
        self.exact_tol = float(exact_tol)
        # distributed=True: the state is sharded over the ranks of the default torch.distributed
        # (NCCL) process group, one process per GPU; every rank calls the same methods
        self.distributed = bool(distributed)
        if self.distributed and dtype != "complex128":
            raise ValueError("distributed=True supports dtype='complex128' only (sharded states are complex128)")
        if not np.isfinite(anisotropy_delta):
            raise ValueError(f"anisotropy_delta must be finite, got {anisotropy_delta}")
        self.anisotropy_delta = float(anisotropy_delta) if abs(anisotropy_delta) > 1e-15 else 0.0
        if self.anisotropy_delta and self.distributed:
            raise ValueError("anisotropy_delta != 0 is not supported with distributed=True")
        self._engine = None
        self._exact_reps_cache: dict[tuple[float, float], int] = {}
        self._calibrate_direct_max_n = _CALIBRATE_DIRECT_MAX_N
        self.exact_calibration: dict | None = None

    # -- validation (same messages as the reference) ---------------------------------------------
    @staticmethod
    def _validate_n_oscillators(n_oscillators: int) -> int:
        if not isinstance(n_oscillators, int) or n_oscillators < 1:
            raise ValueError(f"n_oscillators must be a positive integer, got {n_oscillators}")
        return n_oscillators

    @staticmethod
    def _validate_coupling_inputs(n: int, K: np.ndarray, omega: np.ndarray) -> None:
        if K.shape != (n, n):
            raise ValueError(f"K_coupling shape must be ({n}, {n}), got {K.shape}")
        if omega.shape != (n,):
            raise ValueError(f"omega_natural shape must be ({n},), got {omega.shape}")
        if not np.all(np.isfinite(K)):
            raise ValueError("K_coupling must contain only finite values")
        if not np.all(np.isfinite(omega)):
            raise ValueError("omega_natural must contain only finite values")
        if not np.allclose(K, K.T, atol=1e-12, rtol=1e-12):
            raise ValueError("K_coupling must be symmetric for the XY Kuramoto solver")

    # -- engine ----------------------------------------------------------------------------------
    def _budget_gate(self, max_statevector_gib: float | None) -> None:
        _lib.load()  # fail loudly (no CPU fallback) before anything else
        try:
            free, _total = device_memory(self.device)
        except RuntimeError:
            free = None
        buffers = 2 if self.n > _SMALL_N else 1  # tiled passes ping-pong between two buffers
        n_devices = 1
        if self.distributed:
            import torch.distributed as dist

            n_devices = dist.get_world_size() if dist.is_initialized() else 1
        require_statevector_allocation(
            self.n,
            itemsize=16 if self.dtype == "complex128" else 8,
            buffers=buffers,
            n_devices=n_devices,
            max_gib=max_statevector_gib,
            device_free_bytes=free,
            label="Kuramoto statevector trajectory",
        )

    def engine(self) -> XYKEngine:
        """The device engine holding the live statevector (created on first use)."""
        if self._engine is None:
            if self.distributed:
                from .sharded import ShardedXYKEngine

                eng = ShardedXYKEngine(self.n, device=self.device)
            else:
                eng = XYKEngine(self.n, dtype=self.dtype, device=self.device)
            self._set_problem_on(eng, self.K, self.omega)
            self._engine = eng
        return self._engine

    def _set_problem_on(self, eng, K, omega) -> None:
        if self.anisotropy_delta:
            eng.set_problem(K, omega, self.anisotropy_delta)
            if self.evolution == "trotter":
                eng.set_engine("simple")  # the literal product formula in the reference's term order: XX, YY, ZZ per pair
        else:
            eng.set_problem(K, omega)

    def update_problem(self, K_coupling, omega_natural=None) -> None:
        """Re-parameterise (K, omega) on the live engine: same validation as the constructor, the
        device statevector and its buffers stay as they are (the realtime-feedback controller of the
        reference re-scales K every tick on a live state, control/realtime_feedback.py:98-137)."""
        K = _as_real_numeric_array("K_coupling", K_coupling)
        omega = self.omega if omega_natural is None else _as_real_numeric_array("omega_natural", omega_natural)
        self._validate_coupling_inputs(self.n, K, omega)
        np.fill_diagonal(K, 0.0)
        self.K, self.omega = K, omega
        self._exact_reps_cache.clear()
        if self._engine is not None:
            self._set_problem_on(self._engine, self.K, self.omega)

    def close(self) -> None:
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def build_hamiltonian(self):
        """Ordered term list ``(z_terms, pair_terms)`` -- the gate order the engine applies."""
        from .hamiltonian import xy_terms

        return xy_terms(self.K, self.omega)

    # -- evolution -------------------------------------------------------------------------------
    def _formula(self, step_dt: float, trotter_steps: int) -> tuple[int, int]:
        """(order, reps) the engine runs for one interval."""
        if self.evolution == "trotter":
            return self.trotter_order, trotter_steps
        return 6, self._calibrated_exact_reps(step_dt)

    def _uniform_formula(self, step_sizes: np.ndarray, trotter_steps: int) -> tuple[int, int] | None:
        """(order, reps) when every interval of the grid runs the same formula, else None."""
        forms = {self._formula(float(s), trotter_steps) for s in np.unique(np.round(step_sizes, 15))}
        return forms.pop() if len(forms) == 1 else None

    def _calibrated_exact_reps(self, step_dt: float, nsteps: int = 1) -> int:
        """Smallest power-of-two repetition count of the 6th-order composition whose observables
        over one interval agree with the doubled count to ``exact_tol / nsteps`` (a-posteriori
        halving check; the error of a unitary splitting accumulates at most linearly in the step
        count).  Up to ``_CALIBRATE_DIRECT_MAX_N`` qubits the check runs on a private engine of the
        full problem; above that (and for sharded states) on a reduced problem of the most strongly
        coupled qubits with the couplings scaled up to the full problem's largest row sum, and the
        count it finds is doubled.  ``self.exact_calibration`` records which of the two was used."""
        tol_step = self.exact_tol / max(4, int(nsteps))
        key = (round(float(step_dt), 15), tol_step)
        for (k_dt, k_tol), v in self._exact_reps_cache.items():
            if k_dt == key[0] and k_tol <= tol_step:
                return v
        if step_dt == 0.0:
            return 1
        direct = self.n <= self._calibrate_direct_max_n and not self.distributed
        if direct:
            n_cal, K_cal, w_cal, safety = self.n, self.K, self.omega, 1
            self.exact_calibration = {"kind": "direct", "n": self.n}
        else:
            n_cal = min(self.n, _CALIBRATE_PROXY_N)
            K_cal, w_cal = reduced_problem_for_calibration(self.K, self.omega, n_cal)
            safety = 2
            self.exact_calibration = {"kind": "reduced_problem", "n": n_cal, "safety_factor": safety}
        eng = XYKEngine(n_cal, dtype=self.dtype, device=self.device)
        try:
            self._set_problem_on(eng, K_cal, w_cal)

            def observe(reps: int) -> np.ndarray:
                eng.init_product_state()
                eng.apply_layers(step_dt, 6, reps)
                ex, ey = eng.xy_expectations()
                return np.concatenate([ex, ey])

            chosen = halving_reps(observe, tol_step)
        finally:
            eng.close()
        chosen = max(1, safety * chosen)
        self._exact_reps_cache[key] = chosen
        return chosen

    def evolve(self, time: float, trotter_steps: int | None = None) -> None:
        """Advance the live device state by ``time`` (the reference returns a circuit that
        ``Statevector.evolve`` then applies, :132-154; here the application is the call)."""
        if not np.isfinite(time) or time < 0.0:
            raise ValueError(f"time must be finite and non-negative, got {time}")
        trotter_steps = self.evolution_config.evolve_steps if trotter_steps is None else trotter_steps
        if not isinstance(trotter_steps, int) or trotter_steps < 1:
            raise ValueError(f"trotter_steps must be a positive integer, got {trotter_steps}")
        if self.evolution == "chebyshev":
            self.engine().exact_step(float(time), self.exact_tol * 1e-3)
            return
        order, reps = self._formula(float(time), trotter_steps)
        self.engine().apply_layers(float(time), order, reps)

    def prepare_initial_state(self) -> None:
        """|psi0> = prod_i Ry(omega_i mod 2pi)|0> on the device (reference :236-242)."""
        self.engine().init_product_state()

    def prepare_product_state(self, angles) -> None:
        """|psi> = prod_i Ry(angles_i)|0> on the device: zeros give |0...0> (``Statevector.from_label("0" * n)``,
        reference scripts/bench_s2_scaling_lite.py:262), pi/2 gives |+>^n (phase/floquet_kuramoto.py:131)."""
        self.engine().init_product_angles(np.asarray(angles, dtype=np.float64))

    def measure_order_parameter(self, sv=None) -> tuple[float, float]:
        """R exp(i psi) = (1/N) sum_j (<X_j> + i <Y_j>) of the live state (reference :156-185).

        ``sv`` may be a complex array of 2^n amplitudes to measure instead of the live state."""
        if sv is None:
            return self.engine().order_parameter()
        with self._scratch_engine(sv) as eng:  # a pure function of sv, like the reference: the live state is untouched
            return eng.order_parameter()

    def energy_expectation(self, sv=None) -> float:
        """<H> of the live state (reference :260-264)."""
        if sv is None:
            return self.engine().energy()
        with self._scratch_engine(sv) as eng:
            return eng.energy()

    def _scratch_engine(self, sv) -> XYKEngine:
        """A temporary single-device engine holding the external state ``sv`` (2^n amplitudes, canonical order)."""
        sv = np.asarray(sv)
        if sv.shape != (1 << self.n,):
            raise ValueError(f"statevector must hold 2^n = {1 << self.n} amplitudes, got shape {sv.shape}")
        eng = XYKEngine(self.n, dtype=self.dtype, device=self.device)
        try:
            eng.set_problem(self.K, self.omega)
            eng.set_state(sv)
        except Exception:
            eng.close()
            raise
        return eng

    def statevector(self) -> np.ndarray:
        """Host copy of the live state in canonical (little-endian) index order."""
        return self.engine().get_state()

    def run(
        self,
        t_max: float,
        dt: float,
        trotter_per_step: int | None = None,
        *,
        max_statevector_gib: float | None = None,
    ) -> TrajectoryResult:
        """Time-stepped evolution returning R(t) (reference :187-258)."""
        if not np.isfinite(t_max) or t_max < 0.0:
            raise ValueError(f"t_max must be finite and non-negative, got {t_max}")
        if not np.isfinite(dt) or dt <= 0.0:
            raise ValueError(f"dt must be finite and positive, got {dt}")
        if trotter_per_step is None:
            trotter_per_step = self.evolution_config.run_steps_per_step
        if not isinstance(trotter_per_step, int) or trotter_per_step < 1:
            raise ValueError(f"trotter_per_step must be a positive integer, got {trotter_per_step}")
        self._budget_gate(max_statevector_gib)

        times, step_sizes = time_grid(float(t_max), float(dt))
        exact_reps: dict[float, int] = {}
        if self.evolution == "reference":
            for s in np.unique(np.round(step_sizes, 15)):
                exact_reps[float(s)] = self._calibrated_exact_reps(float(s), len(step_sizes))

        if self.evolution == "chebyshev":
            # exact step without a product formula: Chebyshev expansion on the fused H.phi sweep (xyk_exact_step)
            eng = self.engine()
            eng.init_product_state()
            R_history = np.zeros(times.shape[0])
            R_history[0], _ = eng.order_parameter()
            cheb_terms = []
            for step, step_dt in enumerate(step_sizes, start=1):
                cheb_terms.append(eng.exact_step(float(step_dt), self.exact_tol * 1e-3))
                R_history[step], _ = eng.order_parameter()
        elif self.n <= _SMALL_N and self.dtype == "complex128" and not self.distributed and not self.anisotropy_delta:
            R_history = self._run_small(step_sizes, trotter_per_step, exact_reps)
        elif not self.distributed and self._uniform_formula(step_sizes, trotter_per_step) is not None:
            # one C call enqueues the whole trajectory (state preparation, layers, R(t) on the device, one read-back)
            order, reps = self._uniform_formula(step_sizes, trotter_per_step)
            R_history = self.engine().run(step_sizes, reps, order)
        else:
            eng = self.engine()
            eng.init_product_state()
            R_history = np.zeros(times.shape[0])
            R_history[0], _ = eng.order_parameter()
            for step, step_dt in enumerate(step_sizes, start=1):
                order, reps = self._formula(float(step_dt), trotter_per_step)
                eng.apply_layers(float(step_dt), order, reps)
                R_history[step], _ = eng.order_parameter()

        meta = {
            "backend": "gpu_statevector",
            "trotter_per_step": trotter_per_step,
            "trotter_order": self.trotter_order,
            "evolution": self.evolution,
            "dtype": self.dtype,
        }
        if self.evolution == "chebyshev":
            meta["chebyshev_terms_per_step"] = cheb_terms
        if self.evolution == "reference":
            meta["exact_composition"] = {"order": 6, "reps": {str(k): v for k, v in exact_reps.items()},
                                         "calibration": self.exact_calibration}
        return TrajectoryResult(times=times, R=R_history, metadata=meta)

    def _run_small(self, step_sizes: np.ndarray, trotter_per_step: int, exact_reps: dict[float, int]) -> np.ndarray:
        """n <= 12: one kernel launch runs the whole trajectory (state resident in shared memory)."""
        if self.evolution == "trotter":
            order, reps = self.trotter_order, trotter_per_step
        else:
            order, reps = 6, max(exact_reps.values()) if exact_reps else 1
        R = batch_run(self.K[None], self.omega[None], step_sizes, reps, order, device=self.device)
        return R[0]
