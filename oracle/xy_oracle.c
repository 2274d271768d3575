/* CPU oracle (C restatement) for the Kuramoto-XY statevector Trotter path.
 *
 * TEST INFRASTRUCTURE ONLY -- see oracle/xy_oracle.py for the full header.  Used by tests/
 * (parity at n up to ~24), __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  The product path (scpn_quantum_control_b200/) never links it.
 *
 * Reference (relative to /root/reference):
 *   gate order / coefficients : src/scpn_quantum_control/bridge/knm_hamiltonian.py:172-195
 *   initial state             : src/scpn_quantum_control/phase/xy_kuramoto.py:236-242
 *   observable                : scpn_quantum_engine/src/pauli.rs:195-211
 *   Trotter definition        : Qiskit 2.5.1 LieTrotter / SuzukiTrotter(order=2) (third party,
 *                               not vendored; semantics restated in xy_oracle.py)
 *
 * The Python oracle is the primary restatement (pinned against the reference goldens);
 * this file is validated against it in tests/test_oracle_c.py.
 *
 * State layout: interleaved complex128 (re, im), qubit q = bit q of the index.
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC xy_oracle.c -o _build/libxyoracle.so -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS_SPARSE 1e-15

int xyo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void xyo_set_num_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* psi0[k] = prod_i (bit_i(k) ? sin : cos)((omega_i mod 2pi)/2) */
void xyo_init_state(int n, const double *omega, double *psi) {
    const int64_t dim = (int64_t)1 << n;
    double c[64], s[64];
    for (int i = 0; i < n; ++i) {
        double a = fmod(omega[i], 2.0 * M_PI);
        if (a < 0.0) a += 2.0 * M_PI; /* Python's % is non-negative for positive modulus */
        c[i] = cos(a / 2.0);
        s[i] = sin(a / 2.0);
    }
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < dim; ++k) {
        double v = 1.0;
        for (int i = 0; i < n; ++i) v *= ((k >> i) & 1) ? s[i] : c[i];
        psi[2 * k] = v;
        psi[2 * k + 1] = 0.0;
    }
}

/* psi[k] *= exp(+i tau sum_i omega_i (1 - 2 bit_i(k))), terms with |omega_i| <= eps skipped */
void xyo_apply_z_phase(int n, const double *omega, double tau, double *psi) {
    const int64_t dim = (int64_t)1 << n;
    /* factorised tables over the low and high halves of the index */
    const int nl = n / 2, nh = n - nl;
    const int64_t dl = (int64_t)1 << nl, dh = (int64_t)1 << nh;
    if (dl > 65536 || dh > 65536) { /* n > 32: direct evaluation */
#pragma omp parallel for schedule(static)
        for (int64_t k = 0; k < dim; ++k) {
            double ph = 0.0;
            for (int i = 0; i < n; ++i)
                if (fabs(omega[i]) > EPS_SPARSE) ph += omega[i] * (1.0 - 2.0 * (double)((k >> i) & 1));
            ph *= tau;
            const double cr = cos(ph), ci = sin(ph);
            const double re = psi[2 * k], im = psi[2 * k + 1];
            psi[2 * k] = re * cr - im * ci;
            psi[2 * k + 1] = re * ci + im * cr;
        }
        return;
    }
    double *tl = (double *)malloc(sizeof(double) * 2 * (size_t)dl);
    double *th = (double *)malloc(sizeof(double) * 2 * (size_t)dh);
    for (int64_t a = 0; a < dl; ++a) {
        double ph = 0.0;
        for (int i = 0; i < nl; ++i)
            if (fabs(omega[i]) > EPS_SPARSE) ph += omega[i] * (1.0 - 2.0 * (double)((a >> i) & 1));
        tl[2 * a] = cos(tau * ph);
        tl[2 * a + 1] = sin(tau * ph);
    }
    for (int64_t b = 0; b < dh; ++b) {
        double ph = 0.0;
        for (int i = 0; i < nh; ++i)
            if (fabs(omega[nl + i]) > EPS_SPARSE)
                ph += omega[nl + i] * (1.0 - 2.0 * (double)((b >> i) & 1));
        th[2 * b] = cos(tau * ph);
        th[2 * b + 1] = sin(tau * ph);
    }
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < dh; ++b) {
        const double hr = th[2 * b], hi = th[2 * b + 1];
        for (int64_t a = 0; a < dl; ++a) {
            const double cr = hr * tl[2 * a] - hi * tl[2 * a + 1];
            const double ci = hr * tl[2 * a + 1] + hi * tl[2 * a];
            const int64_t k = (b << nl) | a;
            const double re = psi[2 * k], im = psi[2 * k + 1];
            psi[2 * k] = re * cr - im * ci;
            psi[2 * k + 1] = re * ci + im * cr;
        }
    }
    free(tl);
    free(th);
}

/* (a, b) = (psi[..1_i..0_j..], psi[..0_i..1_j..]) <- (c a + i s b, c b + i s a), phi = 2 K tau */
void xyo_apply_pair(int n, int i, int j, double phi, double *psi) {
    const int64_t quarter = (int64_t)1 << (n - 2);
    const double c = cos(phi), s = sin(phi);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int64_t mi = (int64_t)1 << i, mj = (int64_t)1 << j;
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < quarter; ++t) {
        /* insert zero bits at positions lo and hi */
        int64_t k = ((t >> lo) << (lo + 1)) | (t & (((int64_t)1 << lo) - 1));
        k = ((k >> hi) << (hi + 1)) | (k & (((int64_t)1 << hi) - 1));
        const int64_t ka = k | mi, kb = k | mj;
        const double ar = psi[2 * ka], ai = psi[2 * ka + 1];
        const double br = psi[2 * kb], bi = psi[2 * kb + 1];
        psi[2 * ka] = c * ar - s * bi;
        psi[2 * ka + 1] = c * ai + s * br;
        psi[2 * kb] = c * br - s * ai;
        psi[2 * kb + 1] = c * bi + s * ar;
    }
}

static void sweep_pairs(int n, const double *K, double tau, double *psi, int reverse, int skip_last,
                        int only_last) {
    /* pairs i<j lexicographic with |K_ij| >= eps; K row-major, already symmetrised */
    int li = -1, lj = -1;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (fabs(K[i * n + j]) >= EPS_SPARSE) { li = i; lj = j; }
    if (li < 0) return;
    if (only_last) {
        xyo_apply_pair(n, li, lj, 2.0 * K[li * n + lj] * tau, psi);
        return;
    }
    if (!reverse) {
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) {
                if (fabs(K[i * n + j]) < EPS_SPARSE) continue;
                if (skip_last && i == li && j == lj) continue;
                xyo_apply_pair(n, i, j, 2.0 * K[i * n + j] * tau, psi);
            }
    } else {
        for (int i = n - 1; i >= 0; --i)
            for (int j = n - 1; j > i; --j) {
                if (fabs(K[i * n + j]) < EPS_SPARSE) continue;
                if (skip_last && i == li && j == lj) continue;
                xyo_apply_pair(n, i, j, 2.0 * K[i * n + j] * tau, psi);
            }
    }
}

static int has_pairs(int n, const double *K) {
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (fabs(K[i * n + j]) >= EPS_SPARSE) return 1;
    return 0;
}

/* one LieTrotter repetition */
void xyo_layer_order1(int n, const double *K, const double *omega, double tau, double *psi) {
    xyo_apply_z_phase(n, omega, tau, psi);
    sweep_pairs(n, K, tau, psi, 0, 0, 0);
}

/* one SuzukiTrotter(order=2) repetition */
void xyo_layer_order2(int n, const double *K, const double *omega, double tau, double *psi) {
    if (!has_pairs(n, K)) {
        xyo_apply_z_phase(n, omega, tau, psi);
        return;
    }
    xyo_apply_z_phase(n, omega, tau / 2.0, psi);
    sweep_pairs(n, K, tau / 2.0, psi, 0, 1, 0);
    sweep_pairs(n, K, tau, psi, 0, 0, 1);
    sweep_pairs(n, K, tau / 2.0, psi, 1, 1, 0);
    xyo_apply_z_phase(n, omega, tau / 2.0, psi);
}

/* reps repetitions of the order-1 / order-2 formula over one grid interval step_dt */
void xyo_evolve(int n, const double *K, const double *omega, double step_dt, int reps, int order,
                double *psi) {
    const double tau = step_dt / (double)reps;
    for (int r = 0; r < reps; ++r) {
        if (order == 1) xyo_layer_order1(n, K, omega, tau, psi);
        else xyo_layer_order2(n, K, omega, tau, psi);
    }
}

/* S_q = 2 sum_{k: bit_q=0} conj(psi_k) psi_{k|2^q};  exp_x[q] = Re S_q, exp_y[q] = Im S_q */
void xyo_xy_expectations(int n, const double *psi, double *exp_x, double *exp_y) {
    const int64_t half = (int64_t)1 << (n - 1);
    for (int q = 0; q < n; ++q) {
        const int64_t m = (int64_t)1 << q;
        double ex = 0.0, ey = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : ex, ey)
        for (int64_t t = 0; t < half; ++t) {
            const int64_t k = ((t >> q) << (q + 1)) | (t & (m - 1));
            const int64_t f = k | m;
            const double ar = psi[2 * k], ai = psi[2 * k + 1];
            const double br = psi[2 * f], bi = psi[2 * f + 1];
            ex += ar * br + ai * bi;
            ey += ar * bi - ai * br;
        }
        exp_x[q] = 2.0 * ex;
        exp_y[q] = 2.0 * ey;
    }
}

double xyo_order_parameter(int n, const double *psi, double *psi_angle) {
    double ex[64], ey[64], zr = 0.0, zi = 0.0;
    xyo_xy_expectations(n, psi, ex, ey);
    for (int q = 0; q < n; ++q) { zr += ex[q]; zi += ey[q]; }
    zr /= n; zi /= n;
    if (psi_angle) *psi_angle = atan2(zi, zr);
    return sqrt(zr * zr + zi * zi);
}

/* run(): R at every grid point; step_sizes[m]; R_out[m+1]; psi must hold 2*2^n doubles */
void xyo_run(int n, const double *K, const double *omega, const double *step_sizes, int m, int reps,
             int order, double *psi, double *R_out) {
    xyo_init_state(n, omega, psi);
    R_out[0] = xyo_order_parameter(n, psi, 0);
    for (int s = 0; s < m; ++s) {
        xyo_evolve(n, K, omega, step_sizes[s], reps, order, psi);
        R_out[s + 1] = xyo_order_parameter(n, psi, 0);
    }
}
