"""Thin object wrapper over one ``xyk_handle`` (one GPU, one statevector or shard)."""

from __future__ import annotations

import ctypes
import json
from ctypes import POINTER, byref, c_double, c_int, c_size_t, c_void_p

import numpy as np

from . import _lib
from .dense_budget import DenseAllocationError

_DP = POINTER(c_double)


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(_DP)


class XYKEngine:
    """Owns a device statevector of ``n`` qubits (or the ``rank``-th of ``world_size`` shards)."""

    def __init__(self, n: int, *, dtype: str = "complex128", device: int = 0, world_size: int = 1, rank: int = 0):
        self._lib = _lib.load()
        if dtype not in ("complex128", "complex64"):
            raise ValueError(f"dtype must be 'complex128' or 'complex64', got {dtype!r}")
        self.n = int(n)
        self.dtype = np.dtype(dtype)
        self.device = int(device)
        self.world_size = int(world_size)
        self.rank = int(rank)
        self._h = c_void_p()
        rc = self._lib.xyk_create(self.n, 0 if dtype == "complex128" else 1, self.device, self.world_size,
                                  self.rank, byref(self._h))
        if rc != _lib.XYK_OK:
            msg = _lib.last_error(None)
            self._h = c_void_p()
            if rc == _lib.XYK_ERR_NOMEM:
                raise DenseAllocationError(f"device allocation failed: {msg}")
            if rc == _lib.XYK_ERR_INVALID:
                raise ValueError(msg)
            raise RuntimeError(f"xyk_create failed ({rc}): {msg}")
        gbits = self.world_size.bit_length() - 1
        self.n_local = self.n - gbits

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int) -> int:
        if rc in (_lib.XYK_OK, _lib.XYK_NEED_EXCHANGE, _lib.XYK_NEED_BARRIER):
            return rc
        msg = _lib.last_error(self._h)
        if rc == _lib.XYK_ERR_NOMEM:
            raise DenseAllocationError(msg)
        if rc == _lib.XYK_ERR_INVALID:
            raise ValueError(msg)
        raise RuntimeError(f"xyk error {rc}: {msg}")

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.xyk_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- problem / state ----------------------------------------------------------------------
    def set_option(self, option: int, value: int) -> None:
        self._check(self._lib.xyk_set_option(self._h, option, int(value)))

    def set_engine(self, name: str) -> None:
        table = {"auto": _lib.ENGINE_AUTO, "simple": _lib.ENGINE_SIMPLE, "tiled": _lib.ENGINE_TILED}
        self.set_option(_lib.OPT_ENGINE, table[name])

    def set_problem(self, K: np.ndarray, omega: np.ndarray, delta: float = 0.0) -> None:
        """(K, omega) of the XY model; ``delta`` != 0: XXZ anisotropy (knm_to_xxz_hamiltonian)."""
        K = np.ascontiguousarray(K, dtype=np.float64)
        omega = np.ascontiguousarray(omega, dtype=np.float64)
        if K.shape != (self.n, self.n) or omega.shape != (self.n,):
            raise ValueError("K must be (n, n) and omega (n,)")
        if delta == 0.0:
            self._check(self._lib.xyk_set_problem(self._h, _dptr(K), _dptr(omega)))
        else:
            self._check(self._lib.xyk_set_problem_xxz(self._h, _dptr(K), _dptr(omega), float(delta)))

    def term_counts(self) -> tuple[int, int]:
        a, b = c_int(), c_int()
        self._check(self._lib.xyk_term_counts(self._h, byref(a), byref(b)))
        return a.value, b.value

    def init_product_state(self) -> None:
        self._check(self._lib.xyk_init_product_state(self._h))

    def apply_occupation_decay(self, rate_zero: float, rate_one: float = 0.0, scale: float = 1.0) -> None:
        """psi_k *= scale * exp(-rate_zero * #zero bits(k) - rate_one * #one bits(k)) (MCWF no-jump damping, re-scaling)."""
        self._check(self._lib.xyk_apply_occupation_decay(self._h, float(rate_zero), float(rate_one), float(scale)))

    def apply_jump(self, qubit: int, kind: int) -> None:
        """kind 0: psi <- |1><0|_q psi (the reference's "-" matrix); kind 1: psi <- Z_q psi.  Not normalised."""
        self._check(self._lib.xyk_apply_jump(self._h, int(qubit), int(kind)))

    def exact_step(self, dt: float, tol: float = 1e-13) -> int:
        """psi <- exp(-i H dt) psi by Chebyshev expansion (no splitting error); returns the number of terms."""
        used = c_int()
        self._check(self._lib.xyk_exact_step(self._h, float(dt), float(tol), byref(used)))
        return used.value

    def init_product_angles(self, angles: np.ndarray) -> None:
        """|psi> = prod_q Ry(angles[q])|0> (0: |0...0>, pi/2: |+>^n)."""
        a = np.ascontiguousarray(angles, dtype=np.float64)
        if a.shape != (self.n,):
            raise ValueError(f"angles must have shape ({self.n},), got {a.shape}")
        self._check(self._lib.xyk_init_product_angles(self._h, _dptr(a)))

    def apply_layers(self, delta: float, order: int = 1, reps: int = 1) -> int:
        return self._check(self._lib.xyk_apply_layers(self._h, float(delta), int(order), int(reps)))

    def xy_expectations(self) -> tuple[np.ndarray, np.ndarray]:
        ex = np.zeros(self.n)
        ey = np.zeros(self.n)
        self._check(self._lib.xyk_xy_expectations(self._h, _dptr(ex), _dptr(ey)))
        return ex, ey

    def xy_expectations_peer(self) -> tuple[np.ndarray, np.ndarray]:
        """Sharded handles: this rank's partial <X_q>, <Y_q> of the sharded qubits, from peer reads (no exchange)."""
        ex, ey = np.zeros(self.n), np.zeros(self.n)
        self._check(self._lib.xyk_xy_expectations_peer(self._h, _dptr(ex), _dptr(ey)))
        return ex, ey

    def order_parameter(self) -> tuple[float, float]:
        r, p = c_double(), c_double()
        self._check(self._lib.xyk_order_parameter(self._h, byref(r), byref(p)))
        return r.value, p.value

    def z_expectations(self) -> np.ndarray:
        ez = np.zeros(self.n)
        self._check(self._lib.xyk_z_expectations(self._h, _dptr(ez)))
        return ez

    def correlation_matrix_xy(self) -> np.ndarray:
        C = np.zeros((self.n, self.n))
        self._check(self._lib.xyk_correlation_matrix_xy(self._h, _dptr(C)))
        return C

    def energy(self) -> float:
        e = c_double()
        self._check(self._lib.xyk_energy(self._h, byref(e)))
        return e.value

    def norm2(self) -> float:
        v = c_double()
        self._check(self._lib.xyk_norm2(self._h, byref(v)))
        return v.value

    def apply_ry(self, angles: np.ndarray) -> None:
        """psi <- prod_q Ry_q(angles[q]) psi on the live state."""
        a = np.ascontiguousarray(angles, dtype=np.float64)
        if a.shape != (self.n,):
            raise ValueError(f"angles must have shape ({self.n},), got {a.shape}")
        self._check(self._lib.xyk_apply_ry(self._h, _dptr(a)))

    def run(self, step_sizes: np.ndarray, reps: int, order: int) -> np.ndarray:
        steps = np.ascontiguousarray(step_sizes, dtype=np.float64)
        R = np.zeros(steps.shape[0] + 1)
        self._check(self._lib.xyk_run(self._h, _dptr(steps), steps.shape[0], int(reps), int(order), _dptr(R)))
        return R

    def get_state(self) -> np.ndarray:
        out = np.empty(1 << self.n_local, dtype=self.dtype)
        self._check(self._lib.xyk_get_state(self._h, out.ctypes.data_as(c_void_p), 0))
        return out

    def set_state(self, psi: np.ndarray) -> None:
        psi = np.ascontiguousarray(psi, dtype=self.dtype)
        if psi.shape != (1 << self.n_local,):
            raise ValueError(f"state must have {1 << self.n_local} amplitudes")
        self._check(self._lib.xyk_set_state(self._h, psi.ctypes.data_as(c_void_p), 0))

    def get_state_device(self, dst_ptr: int) -> None:
        """Canonical-order copy into a caller-owned device buffer (e.g. a torch tensor's data_ptr)."""
        self._check(self._lib.xyk_get_state(self._h, c_void_p(dst_ptr), 1))

    def set_state_device(self, src_ptr: int) -> None:
        self._check(self._lib.xyk_set_state(self._h, c_void_p(src_ptr), 1))

    def state_buffer(self) -> tuple[int, int]:
        p, b = c_void_p(), c_size_t()
        self._check(self._lib.xyk_state_buffer(self._h, byref(p), byref(b)))
        return int(p.value or 0), int(b.value)

    def layout(self) -> list[int]:
        arr = (c_int * self.n)()
        self._check(self._lib.xyk_get_layout(self._h, arr))
        return list(arr)

    def canonicalise(self) -> None:
        self._check(self._lib.xyk_canonicalise(self._h))

    def synchronize(self) -> None:
        self._check(self._lib.xyk_synchronize(self._h))

    def stream(self) -> int:
        s = c_void_p()
        self._check(self._lib.xyk_stream(self._h, byref(s)))
        return int(s.value or 0)

    def enable_timing(self, on: bool = True) -> None:
        self._check(self._lib.xyk_enable_timing(self._h, 1 if on else 0))

    def stats(self) -> dict:
        st = _lib.XYKStats()
        self._check(self._lib.xyk_get_stats(self._h, byref(st)))
        return {name: getattr(st, name) for name, _ in st._fields_}


def device_count() -> int:
    c = c_int()
    rc = _lib.load().xyk_device_count(byref(c))
    return c.value if rc == 0 else 0


def device_memory(device: int = 0) -> tuple[int, int]:
    f, t = c_size_t(), c_size_t()
    rc = _lib.load().xyk_device_memory(device, byref(f), byref(t))
    if rc != 0:
        raise RuntimeError(_lib.last_error(None))
    return int(f.value), int(t.value)


def plan_describe(K: np.ndarray, omega: np.ndarray, delta: float, order: int, reps: int, n_local: int | None = None,
                  peer_exchange: bool = False, restore_sharding: bool = False) -> dict:
    """Host-only: the pass program the schedule compiler emits (no GPU needed).  With ``n_local < n`` the
    sharded program (local programs separated by / fused with exchanges) is returned."""
    lib = _lib.load()
    K = np.ascontiguousarray(K, dtype=np.float64)
    omega = np.ascontiguousarray(omega, dtype=np.float64)
    n = omega.shape[0]
    nl = n if n_local is None else n_local
    flag = (2 if peer_exchange else 0) | (4 if restore_sharding else 0)
    need = lib.xyk_plan_describe(nl, flag, _dptr(K), _dptr(omega), n, float(delta), order, reps, None, 0)
    if need < 0:
        raise RuntimeError(_lib.last_error(None))
    buf = ctypes.create_string_buffer(int(need))
    lib.xyk_plan_describe(nl, flag, _dptr(K), _dptr(omega), n, float(delta), order, reps, buf, need)
    return json.loads(buf.value.decode())


def batch_run(K: np.ndarray, omega: np.ndarray, step_sizes: np.ndarray, reps: int, order: int, *, device: int = 0,
              return_states: bool = False):
    """B independent n<=12 problems, one CTA each (``xyk_batch_run``).  K: (B,n,n), omega: (B,n)."""
    lib = _lib.load()
    K = np.ascontiguousarray(K, dtype=np.float64)
    omega = np.ascontiguousarray(omega, dtype=np.float64)
    steps = np.ascontiguousarray(step_sizes, dtype=np.float64)
    if K.ndim != 3 or omega.ndim != 2 or K.shape[0] != omega.shape[0] or K.shape[1:] != (omega.shape[1],) * 2:
        raise ValueError("K must be (B, n, n) and omega (B, n)")
    B, n = omega.shape
    R = np.zeros((B, steps.shape[0] + 1))
    states = np.empty((B, 1 << n), dtype=np.complex128) if return_states else None
    rc = lib.xyk_batch_run(device, n, B, _dptr(K), _dptr(omega), _dptr(steps), steps.shape[0], int(reps), int(order),
                           _dptr(R), states.ctypes.data_as(c_void_p) if return_states else None)
    if rc != 0:
        msg = _lib.last_error(None)
        if rc == _lib.XYK_ERR_NOMEM:
            raise DenseAllocationError(msg)
        if rc == _lib.XYK_ERR_INVALID:
            raise ValueError(msg)
        raise RuntimeError(f"xyk_batch_run failed ({rc}): {msg}")
    return (R, states) if return_states else R


def plan_stages(K: np.ndarray, omega: np.ndarray, delta: float, order: int, reps: int, upto_stage: int = 3) -> dict:
    """Host-only: intermediate results of the schedule compiler's stages (1: tiles, 2: + rounds, 3: + layouts)."""
    lib = _lib.load()
    K = np.ascontiguousarray(K, dtype=np.float64)
    omega = np.ascontiguousarray(omega, dtype=np.float64)
    n = omega.shape[0]
    need = lib.xyk_plan_stages(_dptr(K), _dptr(omega), n, float(delta), order, reps, upto_stage, None, 0)
    if need < 0:
        raise RuntimeError(_lib.last_error(None))
    buf = ctypes.create_string_buffer(int(need))
    lib.xyk_plan_stages(_dptr(K), _dptr(omega), n, float(delta), order, reps, upto_stage, buf, need)
    return json.loads(buf.value.decode())
