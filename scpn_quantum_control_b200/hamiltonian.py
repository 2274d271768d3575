"""Canonical (K, omega) problems and the ordered term list of the XY Hamiltonian

    H = -sum_{i<j} K_ij (X_i X_j + Y_i Y_j) - sum_i omega_i Z_i

(reference: src/scpn_quantum_control/bridge/knm_hamiltonian.py:34-54 OMEGA_N_16, :92-117
build_knm_paper27, :120-138 build_kuramoto_ring, :172-195 term order).  The GPU engine builds
the same ordered, thresholded term list natively (``xyk_set_problem``); ``xy_terms`` is the
host-side view of it for inspection and tests.
"""

from __future__ import annotations

import numpy as np

KNM_SPARSITY_EPS = 1e-15

# Paper 27, Table 1: canonical natural frequencies (rad/s) -- data, reference :34-54
OMEGA_N_16 = np.array(
    [1.329, 2.610, 0.844, 1.520, 0.710, 3.780, 1.055, 0.625,
     2.210, 1.740, 0.480, 3.210, 0.915, 1.410, 2.830, 0.991],
    dtype=np.float64,
)


def omega_for_oscillators(n_oscillators: int) -> np.ndarray:
    """First ``n`` canonical frequencies, periodically extended beyond 16 (reference :57-89)."""
    if not isinstance(n_oscillators, int):
        raise TypeError("n_oscillators must be an integer")
    if n_oscillators < 1:
        raise ValueError("n_oscillators must be >= 1")
    reps = -(-n_oscillators // OMEGA_N_16.size)
    return np.tile(OMEGA_N_16, reps)[:n_oscillators].astype(np.float64, copy=True)


def build_knm_paper27(L: int = 16, K_base: float = 0.45, K_alpha: float = 0.3) -> np.ndarray:
    """K_ij = K_base exp(-K_alpha |i-j|) with the Table-2 anchors and S4.3 boosts (reference :92-117)."""
    d = np.abs(np.subtract.outer(np.arange(L), np.arange(L)))
    K = K_base * np.exp(-K_alpha * d)
    for (i, j), val in (((0, 1), 0.302), ((1, 2), 0.201), ((2, 3), 0.252), ((3, 4), 0.154)):
        if max(i, j) < L:
            K[i, j] = K[j, i] = val
    if L > 15:
        K[0, 15] = K[15, 0] = max(K[0, 15], 0.05)
    if L > 6:
        K[4, 6] = K[6, 4] = max(K[4, 6], 0.15)
    return K


def build_kuramoto_ring(n: int, coupling: float = 1.0, omega=None, rng_seed=None):
    """Nearest-neighbour ring (K, omega) (reference :120-138)."""
    K = np.zeros((n, n), dtype=np.float64)
    idx = np.arange(n)
    K[idx, (idx + 1) % n] = coupling
    K[(idx + 1) % n, idx] = coupling
    if omega is None:
        omega = np.random.default_rng(rng_seed).standard_normal(n)
    return K, np.asarray(omega, dtype=np.float64)


def xy_terms(K: np.ndarray, omega: np.ndarray):
    """Ordered terms: ``[(i, omega_i)]`` then ``[(i, j, K_ij)]`` lexicographic (reference :172-195)."""
    K = np.asarray(K, dtype=np.float64)
    omega = np.asarray(omega, dtype=np.float64)
    n = omega.shape[0]
    if K.shape[0] != n:
        raise ValueError(f"K has {K.shape[0]} rows but omega has {n} elements")
    Ks = (K + K.T) / 2.0
    z_terms = [(i, float(omega[i])) for i in range(n) if abs(omega[i]) > KNM_SPARSITY_EPS]
    iu, ju = np.triu_indices(n, 1)
    keep = np.abs(Ks[iu, ju]) >= KNM_SPARSITY_EPS
    pair_terms = [(int(i), int(j), float(Ks[i, j])) for i, j in zip(iu[keep], ju[keep])]
    return z_terms, pair_terms
