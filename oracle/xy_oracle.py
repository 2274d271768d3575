"""CPU oracle for the Kuramoto-XY exact-statevector Trotter path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``scpn_quantum_control_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or the timed CPU arm.

This is a NumPy restatement of what the reference computes on the path
``auto_solve -> QuantumKuramotoSolver.run`` (reference paths relative to /root/reference):

* canonicalisation            src/scpn_quantum_control/phase/xy_kuramoto.py:37-51,95-99,114-125
* term list (= gate order)    src/scpn_quantum_control/bridge/knm_hamiltonian.py:172-206
* time grid                   src/scpn_quantum_control/phase/xy_kuramoto.py:221-233
* initial product state       src/scpn_quantum_control/phase/xy_kuramoto.py:236-242
                              (NumPy twin: src/scpn_quantum_control/hardware/classical.py:220-231)
* observable                  scpn_quantum_engine/src/pauli.rs:195-211
                              (NumPy twin: src/scpn_quantum_control/hardware/classical.py:285-311)
* R, psi                      src/scpn_quantum_control/phase/xy_kuramoto.py:173,183-184
* H matrix elements (mode E)  scpn_quantum_engine/src/hamiltonian.rs:55-78

The Trotter arithmetic itself lives in a third-party dependency that is NOT vendored in the
reference tree: Qiskit (pinned ``qiskit==2.5.1`` in requirements.txt:24).  Its published
semantics are restated here:

* ``LieTrotter(reps)``: per repetition apply ``exp(-i c_k P_k tau)`` for every term of the
  operator in list order, ``tau = time / reps``  ("Mode T", order 1).
* ``SuzukiTrotter(order=2, reps)``: per repetition ``T_1(tau/2) .. T_{M-1}(tau/2) T_M(tau)
  T_{M-1}(tau/2) .. T_1(tau/2)``  ("Mode T", order 2).
* ``Statevector.evolve`` of the *un-decomposed* ``PauliEvolutionGate`` evaluates the exact
  ``exp(-i H step_dt)``  ("Mode E") -- this is what the reference's stored outputs equal
  (data/classical_quantum_comparison/reproducible_comparison_n8.json,
  results/qutip_comparison_2026-03-30.json).

Parity pinning: Mode E is pinned against the reference's golden vectors in
``tests/golden/reference_goldens.json`` (see ``tests/test_oracle_goldens.py``).  Mode T's term
ordering is pinned only to 3 printed digits (docs/pipeline_performance.md:376,410) because
Qiskit cannot be imported here -- "parity unpinned at 1e-10" for the product-formula ordering;
it follows from the term list above plus Qiskit's documented synthesis order.

Conventions: qubit q <-> bit q of the basis index (little endian).  All arithmetic complex128.
"""

from __future__ import annotations

import math

import numpy as np

SPARSITY_EPS = 1e-15  # src/scpn_quantum_control/_constants.py COUPLING_SPARSITY_EPS


# --------------------------------------------------------------------------------------
# problem canonicalisation and term list
# --------------------------------------------------------------------------------------
def canonicalise(n, K, omega):
    """float64 private copies, symmetric check, zero diagonal (xy_kuramoto.py:95-99,114-125)."""
    K = np.array(np.asarray(K), dtype=np.float64, copy=True)
    omega = np.array(np.asarray(omega), dtype=np.float64, copy=True)
    if K.shape != (n, n) or omega.shape != (n,):
        raise ValueError("bad shapes")
    if not (np.all(np.isfinite(K)) and np.all(np.isfinite(omega))):
        raise ValueError("non-finite input")
    if not np.allclose(K, K.T, atol=1e-12, rtol=1e-12):
        raise ValueError("K not symmetric")
    np.fill_diagonal(K, 0.0)
    return K, omega


def term_list(K, omega):
    """Ordered Hamiltonian terms (knm_hamiltonian.py:172-195).

    Returns ``(z_terms, pair_terms)``: ``z_terms = [(i, omega_i)]`` for ``|omega_i| > eps``
    (coefficient of Z_i is ``-omega_i``) and ``pair_terms = [(i, j, K_ij)]`` for ``i < j``
    lexicographic with ``|K_ij| >= eps`` (coefficient of both XX and YY is ``-K_ij``).
    """
    n = len(omega)
    Ks = (K + K.T) / 2.0
    z_terms = [(i, float(omega[i])) for i in range(n) if abs(omega[i]) > SPARSITY_EPS]
    pair_terms = []
    for i in range(n):
        for j in range(i + 1, n):
            if abs(Ks[i, j]) < SPARSITY_EPS:
                continue
            pair_terms.append((i, j, float(Ks[i, j])))
    return z_terms, pair_terms


def time_grid(t_max, dt):
    """times, step_sizes exactly as xy_kuramoto.py:221-233 builds them."""
    times = [0.0]
    cur = 0.0
    tol = max(np.finfo(float).eps * max(1.0, abs(t_max), abs(dt)) * 16.0, 1e-15)
    while cur + dt < t_max - tol:
        cur += dt
        times.append(cur)
    if t_max > times[-1] + tol:
        times.append(float(t_max))
    else:
        times[-1] = float(t_max)
    times = np.asarray(times, dtype=float)
    return times, np.diff(times)


# --------------------------------------------------------------------------------------
# state preparation, gates, observable
# --------------------------------------------------------------------------------------
def initial_state(omega):
    """prod_i Ry(omega_i mod 2pi)|0>, qubit 0 = LSB (xy_kuramoto.py:236-242)."""
    psi = np.ones(1, dtype=np.complex128)
    for w in omega:  # qubit i becomes bit i: new qubit is the new MSB
        ang = float(w) % (2.0 * math.pi)
        q = np.array([math.cos(ang / 2.0), math.sin(ang / 2.0)], dtype=np.complex128)
        psi = np.kron(q, psi)
    return psi


def apply_z_phase(psi, z_terms, tau):
    """psi[k] *= exp(+i tau sum_i omega_i (1 - 2 bit_i(k)))   (Z_i coefficient -omega_i)."""
    n = int(round(math.log2(psi.size)))
    view = psi.reshape((2,) * n)  # axis a <-> qubit n-1-a
    for i, w in z_terms:
        ax = n - 1 - i
        ph = np.array([np.exp(1j * tau * w), np.exp(-1j * tau * w)])
        shape = [1] * n
        shape[ax] = 2
        view *= ph.reshape(shape)
    return psi


def apply_pair(psi, i, j, phi):
    """exp(+i (phi/2) (X_iX_j + Y_iY_j)): rotate (|..1_i..0_j..>, |..0_i..1_j..>) by
    [[cos phi, i sin phi], [i sin phi, cos phi]]  (phi = 2 K_ij tau)."""
    n = int(round(math.log2(psi.size)))
    view = psi.reshape((2,) * n)
    ai, aj = n - 1 - i, n - 1 - j
    sl10 = [slice(None)] * n
    sl01 = [slice(None)] * n
    sl10[ai], sl10[aj] = 1, 0
    sl01[ai], sl01[aj] = 0, 1
    a = view[tuple(sl10)].copy()
    b = view[tuple(sl01)].copy()
    c, s = math.cos(phi), math.sin(phi)
    view[tuple(sl10)] = c * a + 1j * s * b
    view[tuple(sl01)] = c * b + 1j * s * a
    return psi


def layer_order1(psi, z_terms, pair_terms, tau):
    """One LieTrotter repetition: Z terms, then pairs lexicographic (XX.YY of one pair commute)."""
    apply_z_phase(psi, z_terms, tau)
    for i, j, k in pair_terms:
        apply_pair(psi, i, j, 2.0 * k * tau)
    return psi


def layer_order2(psi, z_terms, pair_terms, tau):
    """One SuzukiTrotter(order=2) repetition over the term list [Z.., XX_p1, YY_p1, ...].

    halves(all but the very last term) . full(last term) . reversed halves.  With at least
    one pair the last term is YY of the last pair; XX/YY of one pair commute, so the last
    pair gets a full ``tau`` and every other pair ``tau/2`` forward and ``tau/2`` backward.
    With no pair terms the last term is the last Z (commuting diagonal: total phase tau).
    """
    if not pair_terms:
        apply_z_phase(psi, z_terms, tau)
        return psi
    apply_z_phase(psi, z_terms, tau / 2.0)
    for i, j, k in pair_terms[:-1]:
        apply_pair(psi, i, j, 2.0 * k * tau / 2.0)
    i, j, k = pair_terms[-1]
    apply_pair(psi, i, j, 2.0 * k * tau)
    for i, j, k in reversed(pair_terms[:-1]):
        apply_pair(psi, i, j, 2.0 * k * tau / 2.0)
    apply_z_phase(psi, z_terms, tau / 2.0)
    return psi


def suzuki_substeps(order):
    """Fractions f_m such that S_order(tau) = prod_m S2(f_m tau) (Suzuki recursion, App. A.9)."""
    if order == 2:
        return [1.0]
    if order % 2 or order < 2:
        raise ValueError("even order >= 2 required")
    p = 1.0 / (4.0 - 4.0 ** (1.0 / (order - 1)))
    inner = suzuki_substeps(order - 2)
    out = []
    for w in (p, p, 1.0 - 4.0 * p, p, p):
        out.extend(w * f for f in inner)
    return out


def evolve_product_formula(psi, z_terms, pair_terms, step_dt, reps, order):
    """Mode T for one grid interval.  order 1/2 follow Qiskit; order 4/6 are Suzuki
    compositions of the order-2 layer (the engine's route to Mode E)."""
    tau = step_dt / reps
    for _ in range(reps):
        if order == 1:
            layer_order1(psi, z_terms, pair_terms, tau)
        else:
            for f in suzuki_substeps(order):
                layer_order2(psi, z_terms, pair_terms, f * tau)
    return psi


def xy_expectations(psi, n):
    """<X_q>, <Y_q> for all q:  S_q = 2 sum_{k: bit_q=0} conj(psi_k) psi_{k|2^q}  (pauli.rs:195-211)."""
    view = psi.reshape((2,) * n)
    ex = np.empty(n)
    ey = np.empty(n)
    for q in range(n):
        ax = n - 1 - q
        lo = np.take(view, 0, axis=ax)
        hi = np.take(view, 1, axis=ax)
        s = 2.0 * np.vdot(lo.ravel(), hi.ravel())  # vdot conjugates its first argument
        ex[q] = s.real
        ey[q] = s.imag
    return ex, ey


def order_parameter(psi, n):
    """(R, psi_angle) = polar form of (1/n) sum_q (<X_q> + i <Y_q>)  (xy_kuramoto.py:173,183-184)."""
    ex, ey = xy_expectations(psi, n)
    z = complex(np.sum(ex) + 1j * np.sum(ey)) / n
    return float(abs(z)), float(np.angle(z))


def z_expectations(psi, n):
    """<Z_q> = sum_k (1 - 2 bit_q(k)) |psi_k|^2  (pauli.rs:163-170)."""
    p = (psi.real**2 + psi.imag**2).reshape((2,) * n)
    out = np.empty(n)
    for q in range(n):
        ax = n - 1 - q
        out[q] = np.take(p, 0, axis=ax).sum() - np.take(p, 1, axis=ax).sum()
    return out


def correlation_matrix_xy(psi, n):
    """C_ij = <X_iX_j + Y_iY_j> (scpn_quantum_engine/src/sectors.rs:105-145)."""
    view = psi.reshape((2,) * n)
    C = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            ai, aj = n - 1 - i, n - 1 - j
            sl10 = [slice(None)] * n
            sl01 = [slice(None)] * n
            sl10[ai], sl10[aj] = 1, 0
            sl01[ai], sl01[aj] = 0, 1
            a = view[tuple(sl10)].ravel()
            b = view[tuple(sl01)].ravel()
            C[i, j] = C[j, i] = 4.0 * np.vdot(a, b).real
    return C


def energy(psi, n, K, omega):
    """<H> = -sum_{i<j} K_ij C_ij - sum_i omega_i <Z_i>  (xy_kuramoto.py:260-264)."""
    Ks = (K + K.T) / 2.0
    C = correlation_matrix_xy(psi, n)
    z = z_expectations(psi, n)
    return float(-np.sum(np.triu(Ks * C, 1)) - np.dot(omega, z))


# --------------------------------------------------------------------------------------
# Mode E: exact per-step propagation
# --------------------------------------------------------------------------------------
def hamiltonian_sparse(K, omega):
    """Sparse H with diag -sum_i omega_i (1-2 b_i), off-diag -2 K_ij on flip-flop pairs
    (hamiltonian.rs:55-78)."""
    from scipy import sparse

    n = len(omega)
    dim = 1 << n
    z_terms, pair_terms = term_list(K, omega)
    idx = np.arange(dim, dtype=np.int64)
    diag = np.zeros(dim)
    for i, w in z_terms:
        diag -= w * (1.0 - 2.0 * ((idx >> i) & 1))
    rows = [idx]
    cols = [idx]
    vals = [diag]
    for i, j, k in pair_terms:
        differ = ((idx >> i) & 1) != ((idx >> j) & 1)
        r = idx[differ]
        rows.append(r)
        cols.append(r ^ ((1 << i) | (1 << j)))
        vals.append(np.full(r.size, -2.0 * k))
    H = sparse.coo_matrix(
        (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim)
    ).tocsc()
    return H


def exact_propagator_dense(K, omega, dt):
    from scipy.linalg import expm

    H = hamiltonian_sparse(K, omega).toarray()
    return expm(-1j * H * dt)


def exact_step(psi, H_sparse, dt):
    from scipy.sparse.linalg import expm_multiply

    return expm_multiply((-1j * dt) * H_sparse, psi)


# --------------------------------------------------------------------------------------
# the path: run(K, omega, t_max, dt) -> times, R
# --------------------------------------------------------------------------------------
def run(n, K, omega, t_max, dt, trotter_per_step=5, order=1, mode="T", return_state=False):
    """Restatement of QuantumKuramotoSolver.run (xy_kuramoto.py:187-258).

    mode "T": product formula of the given ``order`` (1, 2: Qiskit; 4, 6: Suzuki compositions).
    mode "E": exact exp(-i H step_dt) per grid interval (what pinned Qiskit returns).
    """
    K, omega = canonicalise(n, K, omega)
    z_terms, pair_terms = term_list(K, omega)
    times, steps = time_grid(t_max, dt)
    psi = initial_state(omega)
    R = np.zeros(times.shape[0])
    R[0], _ = order_parameter(psi, n)
    if mode == "E":
        H = hamiltonian_sparse(K, omega)
        dense_cache = {}
    for s, step_dt in enumerate(steps, start=1):
        step_dt = float(step_dt)
        if mode == "E":
            if n <= 10:
                if step_dt not in dense_cache:
                    from scipy.linalg import expm

                    dense_cache[step_dt] = expm(-1j * H.toarray() * step_dt)
                psi = dense_cache[step_dt] @ psi
            else:
                psi = exact_step(psi, H, step_dt)
        else:
            psi = evolve_product_formula(psi, z_terms, pair_terms, step_dt, trotter_per_step, order)
        R[s], _ = order_parameter(psi, n)
    if return_state:
        return times, R, psi
    return times, R


def fidelity(a, b):
    return float(abs(np.vdot(a, b)) ** 2 / (np.vdot(a, a).real * np.vdot(b, b).real))


# --------------------------------------------------------------------------------------
# canonical problems (values are data from the reference: knm_hamiltonian.py:34-54,92-138)
# --------------------------------------------------------------------------------------
OMEGA_N_16 = np.array(
    [1.329, 2.610, 0.844, 1.520, 0.710, 3.780, 1.055, 0.625,
     2.210, 1.740, 0.480, 3.210, 0.915, 1.410, 2.830, 0.991],
    dtype=np.float64,
)


def paper27_K(L=16, K_base=0.45, K_alpha=0.3):
    """Paper-27 coupling: K_base exp(-K_alpha |i-j|) + anchors + boosts (knm_hamiltonian.py:92-117)."""
    idx = np.arange(L)
    K = K_base * np.exp(-K_alpha * np.abs(idx[:, None] - idx[None, :]))
    for (i, j), v in {(0, 1): 0.302, (1, 2): 0.201, (2, 3): 0.252, (3, 4): 0.154}.items():
        if i < L and j < L:
            K[i, j] = K[j, i] = v
    if L > 15:
        K[0, 15] = K[15, 0] = max(K[0, 15], 0.05)
    if L > 6:
        K[4, 6] = K[6, 4] = max(K[4, 6], 0.15)
    return K


def ring_K(n, coupling=1.0):
    K = np.zeros((n, n))
    for i in range(n):
        j = (i + 1) % n
        K[i, j] = K[j, i] = coupling
    return K


def random_problem(n, seed):
    """Synthetic all-to-all problem of SURVEY 8(d): K ~ sym U(0.05,1), omega ~ U(0.5,4)."""
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.05, 1.0, (n, n))
    K = (A + A.T) / 2.0
    np.fill_diagonal(K, 0.0)
    omega = rng.uniform(0.5, 4.0, n)
    return K, omega
