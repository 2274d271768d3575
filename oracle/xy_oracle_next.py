"""CPU oracle for the SURVEY 8(f) "next" rows around the Kuramoto-XY path.

TEST INFRASTRUCTURE ONLY (same rule as ``oracle/xy_oracle.py``): nothing under
``scpn_quantum_control_b200/`` may import this module.

NumPy restatements (reference paths relative to /root/reference):

* realtime feedback tick       src/scpn_quantum_control/control/realtime_feedback.py:112-166 (step),
                               :197-204 (shot-sampled R), :206-227 (Ry correction), :316-341 (policy),
                               :402-423 (binomial axis sample, multinomial read-out)
* XXZ term list                src/scpn_quantum_control/bridge/knm_hamiltonian.py:172-206 (delta != 0: per pair
                               XX, YY, ZZ with coefficients -K, -K, -K*delta)
* MCWF trajectories            src/scpn_quantum_control/phase/tensor_jump.py:81-291 (sparse H_eff, Kronecker-order jump
                               operators, jump lottery)
* Floquet drive                src/scpn_quantum_control/phase/floquet_kuramoto.py:67-221 (K(t) = K_base (1 + A cos(Omega t))
                               K_topology sampled at the midpoint of every step, exact exp(-i H dt) per step from |+>^n)

Parity pinning: the feedback tick is pinned by construction to the reference's NumPy policy and generator call order
(no stored golden history exists in the reference tree: "parity unpinned" for the sampled histories, the exact
quantities r_statevector / psi_statevector follow from the pinned path oracle).  XXZ Mode E is pinned against
tests/golden/reference_goldens.json where the reference holds a value; see tests/test_oracle_goldens.py.
"""

from __future__ import annotations

import math

import numpy as np

from . import xy_oracle as oc


# --------------------------------------------------------------------------------------
# single-qubit Ry layer (Qiskit RYGate: [[cos t/2, -sin t/2], [sin t/2, cos t/2]])
# --------------------------------------------------------------------------------------
def apply_ry(psi, n, q, theta):
    c, s = math.cos(theta / 2.0), math.sin(theta / 2.0)
    v = psi.reshape(1 << (n - 1 - q), 2, 1 << q)
    a = v[:, 0, :].copy()
    b = v[:, 1, :].copy()
    v[:, 0, :] = c * a - s * b
    v[:, 1, :] = s * a + c * b
    return psi


# --------------------------------------------------------------------------------------
# realtime feedback controller (control/realtime_feedback.py)
# --------------------------------------------------------------------------------------
def feedback_policy(r, target_r, deadband, base_gain, max_gain):
    """Scalar form of the reference policy (:316-341): returns (action code, next gain, error)."""
    error = target_r - r
    if error > deadband:
        return 1, min(1.0 + base_gain * error, max_gain), error
    if error < -deadband:
        return -1, min(max(1.0 + base_gain * error, 1.0 / max_gain), 1.0), error
    return 0, 1.0, error


def feedback_run(K, omega, n_steps, seed, *, target_r=0.75, deadband=0.04, base_dt=0.08, trotter_steps=3,
                 measurement_shots=128, base_gain=0.8, max_gain=2.0, correction_angle=0.12,
                 feedback_correction_sign=-1.0, order=1, mode="T"):
    """History of RealtimeSyncFeedbackController(K, omega, cfg, order).run(n_steps, seed) (:83-177).

    mode "T": each tick applies the product formula (what the evolve() circuit describes);
    mode "E": each tick applies exp(-i H(K*scale) base_dt) (what pinned Qiskit evaluates).
    Returns (list of dict rows, final state)."""
    omega = np.asarray(omega, dtype=np.float64)
    n = omega.shape[0]
    K = np.asarray(K, dtype=np.float64)
    labels = {-1: "release", 0: "hold", 1: "synchronise"}
    psi = oc.initial_state(omega)
    scale = 1.0
    rows = []
    parent = np.random.default_rng(seed)
    for index in range(n_steps):
        rng = np.random.default_rng(int(parent.integers(0, 2**32 - 1)))
        applied = scale
        Kc, wc = oc.canonicalise(n, K * applied, omega)
        if mode == "E":
            psi = oc.exact_step(psi, oc.hamiltonian_sparse(Kc, wc), base_dt)
        else:
            z, p = oc.term_list(Kc, wc)
            psi = oc.evolve_product_formula(psi, z, p, base_dt, trotter_steps, order)
        ex, ey = oc.xy_expectations(psi, n)
        zc = complex(float(ex.sum()), float(ey.sum())) / n
        r_exact, psi_exact = abs(zc), float(np.angle(zc))
        xs = 2.0 * rng.binomial(measurement_shots, np.clip((1.0 + ex) / 2.0, 0.0, 1.0)).astype(np.float64) / measurement_shots - 1.0
        ys = 2.0 * rng.binomial(measurement_shots, np.clip((1.0 + ey) / 2.0, 0.0, 1.0)).astype(np.float64) / measurement_shots - 1.0
        r_live = abs(complex(float(xs.sum()), float(ys.sum())) / n)
        code, gain, error = feedback_policy(r_live, target_r, deadband, base_gain, max_gain)
        correction = 0.0
        if code != 0:
            correction = float(np.clip(error * correction_angle, -correction_angle, correction_angle))
            if abs(correction) < 1e-15:
                correction = 0.0
            else:
                correction *= -feedback_correction_sign
                for q in range(n):
                    apply_ry(psi, n, q, correction * (1.0 if math.cos(float(omega[q])) >= 0.0 else -1.0))
        prob = np.abs(psi) ** 2
        draws = rng.multinomial(measurement_shots, prob / prob.sum())
        counts = {format(i, f"0{n}b"): int(c) for i, c in enumerate(draws) if int(c) > 0}
        scale = float(gain)
        rows.append(dict(index=index, action=labels[code], r_live=float(r_live), r_statevector=float(r_exact),
                         psi_statevector=psi_exact, error=float(error), applied_coupling_scale=float(applied),
                         next_coupling_scale=scale, correction_angle=correction, readout_counts=counts))
    return rows, psi


# --------------------------------------------------------------------------------------
# XXZ (delta != 0) and Floquet-driven coupling
# --------------------------------------------------------------------------------------
def term_list_xxz(K, omega, delta):
    """(z_terms, pair_terms) with pair_terms = [(i, j, K_ij, K_ij * delta)] in the reference's order
    (knm_hamiltonian.py:172-206: Z_i ascending, then per pair i<j: XX, YY and -- for delta != 0 -- ZZ)."""
    z, p = oc.term_list(K, omega)
    return z, [(i, j, k, k * delta) for (i, j, k) in p]


def apply_zz(psi, n, i, j, theta):
    """psi <- exp(+i theta Z_i Z_j) psi (the term is -K delta Z_i Z_j, evolved for tau: theta = K delta tau)."""
    idx = np.arange(psi.shape[0])
    par = ((idx >> i) ^ (idx >> j)) & 1
    psi *= np.where(par == 0, np.exp(1j * theta), np.exp(-1j * theta))
    return psi


def layer_order1_xxz(psi, n, z_terms, pair_terms, tau):
    """One Lie-Trotter repetition of the XXZ term list: Z phases, then per pair XX.YY rotation and ZZ phase."""
    oc.apply_z_phase(psi, z_terms, tau)
    for (i, j, k, kz) in pair_terms:
        oc.apply_pair(psi, i, j, 2.0 * k * tau)
        if kz != 0.0:
            apply_zz(psi, n, i, j, kz * tau)
    return psi


def layer_order2_xxz(psi, n, z_terms, pair_terms, tau):
    """SuzukiTrotter(order=2) repetition over the term list [Z..., (XX,YY,ZZ)_pairs...]: every term but the last for
    tau/2 forwards, the last term for tau, then backwards.  XX_ij, YY_ij and ZZ_ij of one pair commute, so the pair
    is treated as one block (same operator)."""
    half = tau / 2.0
    oc.apply_z_phase(psi, z_terms, half)
    if pair_terms:
        for (i, j, k, kz) in pair_terms[:-1]:
            oc.apply_pair(psi, i, j, 2.0 * k * half)
            if kz != 0.0:
                apply_zz(psi, n, i, j, kz * half)
        i, j, k, kz = pair_terms[-1]
        oc.apply_pair(psi, i, j, 2.0 * k * tau)
        if kz != 0.0:
            apply_zz(psi, n, i, j, kz * tau)
        for (i, j, k, kz) in reversed(pair_terms[:-1]):
            if kz != 0.0:
                apply_zz(psi, n, i, j, kz * half)
            oc.apply_pair(psi, i, j, 2.0 * k * half)
    oc.apply_z_phase(psi, z_terms, half)
    return psi


def layer_order2_xxz_grouped(psi, n, z_terms, pair_terms, tau):
    """The tiled engine's splitting for delta != 0 (include/xyk.h, xyk_set_problem_xxz): all diagonal terms (Z and ZZ)
    grouped at both ends of the symmetric pair sweep -- D(tau/2) P_fwd P_bwd D(tau/2).  Not the reference's term order
    (that is layer_order2_xxz); the same exact limit."""
    half = tau / 2.0

    def diag(t):
        oc.apply_z_phase(psi, z_terms, t)
        for (i, j, k, kz) in pair_terms:
            if kz != 0.0:
                apply_zz(psi, n, i, j, kz * t)

    diag(half)
    if pair_terms:
        for (i, j, k, kz) in pair_terms[:-1]:
            oc.apply_pair(psi, i, j, 2.0 * k * half)
        i, j, k, kz = pair_terms[-1]
        oc.apply_pair(psi, i, j, 2.0 * k * tau)
        for (i, j, k, kz) in reversed(pair_terms[:-1]):
            oc.apply_pair(psi, i, j, 2.0 * k * half)
    diag(half)
    return psi


def hamiltonian_sparse_xxz(K, omega, delta):
    """H = -sum K_ij (XX + YY + delta ZZ) - sum omega_i Z_i as scipy.sparse (knm_hamiltonian.py:209-217)."""
    import scipy.sparse as sp

    n = omega.shape[0]
    H = oc.hamiltonian_sparse(K, omega).tocsr()
    if delta == 0.0:
        return H
    idx = np.arange(1 << n)
    diag = np.zeros(1 << n)
    for i in range(n):
        for j in range(i + 1, n):
            if abs(K[i, j]) >= oc.SPARSITY_EPS:
                par = ((idx >> i) ^ (idx >> j)) & 1
                diag += -K[i, j] * delta * (1.0 - 2.0 * par)
    return (H + sp.diags(diag)).tocsr()


def run_xxz(n, K, omega, delta, t_max, dt, trotter_per_step=5, order=1, mode="T", return_state=False):
    """run() of a solver whose Hamiltonian is knm_to_xxz_hamiltonian(K, omega, delta)."""
    K, omega = oc.canonicalise(n, K, omega)
    z, p = term_list_xxz(K, omega, delta)
    times, steps = oc.time_grid(t_max, dt)
    psi = oc.initial_state(omega)
    R = np.zeros(times.shape[0])
    R[0], _ = oc.order_parameter(psi, n)
    H = hamiltonian_sparse_xxz(K, omega, delta) if mode == "E" else None
    for s, step_dt in enumerate(steps, start=1):
        if mode == "E":
            psi = oc.exact_step(psi, H, float(step_dt))
        else:
            tau = float(step_dt) / trotter_per_step
            for _ in range(trotter_per_step):
                psi = (layer_order1_xxz if order == 1 else layer_order2_xxz)(psi, n, z, p, tau)
        R[s], _ = oc.order_parameter(psi, n)
    return (times, R, psi) if return_state else (times, R)


def phase_coherence_R(psi, n):
    """floquet_kuramoto._order_parameter (:67-91): |mean_q exp(i atan2(<Y_q>, <X_q>))| (unit phasors, not the
    amplitude-weighted R of the solver)."""
    ex, ey = oc.xy_expectations(psi, n)
    return float(abs(np.mean(np.exp(1j * np.arctan2(ey, ex)))))


def subharmonic_ratio(signal, drive_freq, dt):
    """Power around Omega/2 over power around Omega of the AC part of R(t) (floquet_kuramoto.py:184-221)."""
    signal = np.asarray(signal, dtype=np.float64)
    m = signal.shape[0]
    if m < 4:
        return 0.0
    power = np.abs(np.fft.rfft(signal - signal.mean())) ** 2
    freqs = np.fft.rfftfreq(m, d=dt)
    f0 = drive_freq / (2.0 * np.pi)

    def window(f):
        k = int(np.argmin(np.abs(freqs - f)))
        return float(power[max(0, k - 1):min(power.shape[0], k + 2)].sum())

    p_full, p_half = window(f0), window(f0 / 2.0)
    return 0.0 if p_full < 1e-30 else p_half / p_full


def floquet_evolve(K_topology, omega, K_base, drive_amplitude, drive_frequency, n_periods=10, steps_per_period=20,
                   return_state=False):
    """floquet_evolve (floquet_kuramoto.py:94-182): psi(0) = |+>^n; dt = (2 pi / Omega) / steps_per_period; step s
    applies exp(-i H(K_t K_topology, omega) dt) EXACTLY with K_t = K_base (1 + A cos(Omega (s + 1/2) dt)) (midpoint
    rule); R by phase_coherence_R at every grid point; drive_signal[0] = K_base (1 + A), drive_signal[s + 1] = K_t."""
    K_topology = np.asarray(K_topology, dtype=np.float64)
    omega = np.asarray(omega, dtype=np.float64)
    n = omega.shape[0]
    dt = (2.0 * np.pi / drive_frequency) / steps_per_period
    n_steps = n_periods * steps_per_period
    psi = np.full(1 << n, 1.0 / math.sqrt(1 << n), dtype=np.complex128)
    times = np.arange(n_steps + 1, dtype=np.float64) * dt
    R = np.zeros(n_steps + 1)
    drive = np.zeros(n_steps + 1)
    R[0] = phase_coherence_R(psi, n)
    drive[0] = K_base * (1.0 + drive_amplitude)
    for s in range(n_steps):
        K_t = K_base * (1.0 + drive_amplitude * np.cos(drive_frequency * (s + 0.5) * dt))
        Ks = K_t * K_topology
        Ks = (Ks + Ks.T) / 2.0
        psi = oc.exact_step(psi, oc.hamiltonian_sparse(Ks, omega), dt)
        R[s + 1] = phase_coherence_R(psi, n)
        drive[s + 1] = K_t
    out = dict(times=times, R_values=R, drive_signal=drive, subharmonic_ratio=subharmonic_ratio(R[1:], drive_frequency, dt))
    return (out, psi) if return_state else out


# --------------------------------------------------------------------------------------
# Monte Carlo wave function trajectories (phase/tensor_jump.py:81-291)
# --------------------------------------------------------------------------------------
def _kron_single(op2, qubit, n):
    """Sparse single-qubit operator in the reference's Kronecker order: oscillator 0 is the FIRST factor (:81-93)."""
    import scipy.sparse as sp

    out = sp.identity(1, dtype=np.complex128, format="csr")
    ident = sp.identity(2, dtype=np.complex128, format="csr")
    for i in range(n):
        out = sp.kron(out, sp.csr_matrix(op2) if i == qubit else ident, format="csr")
    return out


def mcwf_order_parameter(psi, n):
    """_order_param_vec (:274-291): |sum_i (<X_i> + i <Y_i>)| / n with oscillator i on index bit n-1-i."""
    ex, ey = oc.xy_expectations(psi, n)   # a sum over all qubits does not care about the labelling
    return float(abs(complex(ex.sum(), ey.sum()) / n))


def mcwf_trajectory(K, omega, gamma_amp=0.05, gamma_deph=0.0, t_max=1.0, dt=0.05, seed=None):
    """Literal restatement of mcwf_trajectory (:120-219): H = the little-endian XY Hamiltonian matrix
    (knm_to_sparse_matrix = SparsePauliOp.to_matrix), state / jump operators in Kronecker (big-endian) order, every step
    expm_multiply(-i H_eff dt) with H_eff = H - (i/2) sum L^dagger L, then the jump lottery in the reference's generator
    call order."""
    from scipy import sparse
    from scipy.sparse.linalg import expm_multiply

    K = np.asarray(K, dtype=np.float64)
    omega = np.asarray(omega, dtype=np.float64)
    n = K.shape[0]
    dim = 1 << n
    rng = np.random.default_rng(seed)
    psi = np.array([1.0 + 0.0j])
    for w in omega:
        a = float(w) % (2 * np.pi)
        psi = np.kron(psi, np.array([np.cos(a / 2), np.sin(a / 2)], dtype=np.complex128))
    if t_max == 0.0:
        return {"times": np.array([0.0]), "R": np.array([mcwf_order_parameter(psi, n)]), "psi_final": psi, "n_jumps": 0}
    Ks = (K + K.T) / 2.0
    H = oc.hamiltonian_sparse(Ks, omega).astype(np.complex128)
    L_ops = []
    for i in range(n):
        if gamma_amp > 0:
            L_ops.append(np.sqrt(gamma_amp) * _kron_single(np.array([[0, 0], [1, 0]], dtype=np.complex128), i, n))
        if gamma_deph > 0:
            L_ops.append(np.sqrt(gamma_deph / 2) * _kron_single(np.array([[1, 0], [0, -1]], dtype=np.complex128), i, n))
    anti = sparse.csr_matrix((dim, dim), dtype=np.complex128)
    for L in L_ops:
        anti = anti + (L.getH() @ L)
    H_eff = (H - 0.5j * anti).tocsc()
    n_steps = max(1, int(np.ceil(t_max / dt)))
    times = np.linspace(0, t_max, n_steps + 1)
    R = np.zeros(n_steps + 1)
    R[0] = mcwf_order_parameter(psi, n)
    n_jumps = 0
    for step in range(1, n_steps + 1):
        step_dt = float(times[step] - times[step - 1])
        psi_new = expm_multiply(-1j * H_eff * step_dt, psi)
        norm_sq = float(np.real(np.dot(psi_new.conj(), psi_new)))
        dp = float(np.clip(1 - norm_sq, 0.0, 1.0))
        r = rng.uniform()
        if r < dp and len(L_ops) > 0:
            probs = np.array([float(np.real(np.dot((L @ psi).conj(), L @ psi))) for L in L_ops])
            total = probs.sum()
            if total > 1e-15:
                probs /= total
                k = rng.choice(len(L_ops), p=probs)
                psi = L_ops[k] @ psi
                psi = psi / np.linalg.norm(psi)
                n_jumps += 1
            else:
                psi = psi_new / np.sqrt(norm_sq) if norm_sq > 1e-15 else psi_new
        else:
            psi = psi_new / np.sqrt(norm_sq) if norm_sq > 1e-15 else psi_new
        R[step] = mcwf_order_parameter(psi, n)
    return {"times": times, "R": R, "psi_final": psi, "n_jumps": n_jumps}
