"""The C restatement (oracle/xy_oracle.c) validated against the primary NumPy oracle."""

import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import xy_oracle as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libxyoracle.so")
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], env=env)
    lib = ctypes.CDLL(SO)
    lib.xyo_run.argtypes = [ctypes.c_int, DP, DP, DP, ctypes.c_int, ctypes.c_int, ctypes.c_int, DP, DP]
    return lib


def run_c(lib, n, K, w, t_max, dt, reps, order):
    K, w = oc.canonicalise(n, K, w)
    Ks = np.ascontiguousarray((K + K.T) / 2)
    times, steps = oc.time_grid(t_max, dt)
    steps = np.ascontiguousarray(steps)
    psi = np.zeros(2 << n)
    R = np.zeros(len(times))
    lib.xyo_run(n, Ks.ctypes.data_as(DP), w.ctypes.data_as(DP), steps.ctypes.data_as(DP), len(steps), reps, order,
                psi.ctypes.data_as(DP), R.ctypes.data_as(DP))
    return R, psi.view(np.complex128)


@pytest.mark.parametrize("n,order", [(1, 1), (2, 2), (3, 1), (4, 2), (7, 1), (10, 2), (11, 1)])
def test_c_oracle_matches_numpy_oracle(lib, n, order):
    K, w = oc.random_problem(n, 31 + n)
    if n == 3:
        w[1] = 0.0
        K[0, 2] = K[2, 0] = 0.0
    R, psi = run_c(lib, n, K, w, 0.35, 0.1, 3, order)
    _, R2, psi2 = oc.run(n, K, w, 0.35, 0.1, 3, order=order, return_state=True)
    assert np.abs(R - R2).max() < 1e-13
    assert np.abs(psi - psi2).max() < 1e-13


def test_c_oracle_negative_omega_uses_python_modulo(lib):
    n = 3
    K = oc.ring_K(3) * 0.3
    w = np.array([-0.4, 1.1, -7.0])
    R, psi = run_c(lib, n, K, w, 0.2, 0.1, 2, 1)
    _, R2, psi2 = oc.run(n, K, w, 0.2, 0.1, 2, order=1, return_state=True)
    assert np.abs(psi - psi2).max() < 1e-13
