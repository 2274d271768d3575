// Small-n regime (n <= 12): the whole statevector lives in shared memory, one CTA per problem,
// the complete run() loop (init, layers, R(t)) executes inside one kernel launch.  Used for the
// batched parameter sweep (BASELINE.json configs[2]) and by the solver for n <= 12.
//
// Reference loop being replaced: QuantumKuramotoSolver.run (xy_kuramoto.py:236-248) called once
// per problem; observable scpn_quantum_engine/src/pauli.rs:195-211.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <string>
#include <vector>

#include "xyk_internal.h"
#include "xyk_kernels.cuh"

namespace xyk {

constexpr int kSmallMaxN = 12;
constexpr int kSmallThreads = 256;
constexpr int kSmallMaxPairs = kSmallMaxN * (kSmallMaxN - 1) / 2;
constexpr int kSmallMaxFrac = 32;

struct SmallParams {
    int32_t n, B, m, reps, order, nfrac;
    double frac[kSmallMaxFrac];  // S2 sub-step fractions (order >= 2)
    const double* K;             // [B][n][n]
    const double* omega;         // [B][n]
    const double* steps;         // [m]
    double* R;                   // [B][m+1]
    double2* final_states;       // nullptr or [B][2^n]
};

struct SmallShared {
    double2 ph_lo[64], ph_hi[64];
    double kk[kSmallMaxPairs];        // K_ij in gate order (0 where dropped)
    double cf[kSmallMaxPairs][2];     // (cos, sin) full angle 2 K sigma
    double ch[kSmallMaxPairs][2];     // (cos, sin) half angle K sigma
    double w[kSmallMaxN];
    double red[2 * kSmallMaxN * 8];
    unsigned char pi[kSmallMaxPairs], pj[kSmallMaxPairs];
    int npairs, last_pair;
};

__device__ __forceinline__ void small_phase(double2* psi, SmallShared& S, int n, double tau) {
    // tables over the low 6 and the remaining high bits
    const int nlo = n < 6 ? n : 6, nhi = n - nlo;
    for (int v = threadIdx.x; v < 128; v += blockDim.x) {
        const bool hi = v >= 64;
        const int idx = v & 63;
        const int cnt = hi ? nhi : nlo, off = hi ? nlo : 0;
        if (idx < (1 << cnt)) {
            double ph = 0.0;
            for (int b = 0; b < cnt; ++b) ph += S.w[off + b] * (1.0 - 2.0 * (double)((idx >> b) & 1));
            double sn, cs;
            sincos(tau * ph, &sn, &cs);
            (hi ? S.ph_hi : S.ph_lo)[idx] = make_double2(cs, sn);
        }
    }
    __syncthreads();
    const int dim = 1 << n;
    for (int k = threadIdx.x; k < dim; k += blockDim.x) {
        const double2 a = S.ph_lo[k & 63], b = S.ph_hi[k >> 6];
        const double fr = a.x * b.x - a.y * b.y, fi = a.x * b.y + a.y * b.x;
        const double2 v = psi[k];
        psi[k] = make_double2(v.x * fr - v.y * fi, v.x * fi + v.y * fr);
    }
    __syncthreads();
}

__device__ __forceinline__ void small_coeffs(SmallShared& S, double sigma) {
    for (int g = threadIdx.x; g < S.npairs; g += blockDim.x) {
        double sn, cs;
        sincos(2.0 * S.kk[g] * sigma, &sn, &cs);
        S.cf[g][0] = cs;
        S.cf[g][1] = sn;
        sincos(S.kk[g] * sigma, &sn, &cs);
        S.ch[g][0] = cs;
        S.ch[g][1] = sn;
    }
    __syncthreads();
}

__device__ __forceinline__ void small_gate(double2* psi, int n, int pi, int pj, double c, double s) {
    const int quarter = 1 << (n - 2);
    const int lo = pi < pj ? pi : pj, hi = pi < pj ? pj : pi;
    for (int t = threadIdx.x; t < quarter; t += blockDim.x) {
        const int k = insert_zero(insert_zero(t, lo), hi);
        const int ka = k | (1 << pi), kb = k | (1 << pj);
        const double2 a = psi[ka], b = psi[kb];
        psi[ka] = make_double2(c * a.x - s * b.y, c * a.y + s * b.x);
        psi[kb] = make_double2(c * b.x - s * a.y, c * b.y + s * a.x);
    }
    __syncthreads();
}

__device__ __forceinline__ double small_order_parameter(const double2* psi, SmallShared& S, int n) {
    double ax[kSmallMaxN], ay[kSmallMaxN];
#pragma unroll
    for (int q = 0; q < kSmallMaxN; ++q) ax[q] = ay[q] = 0.0;
    const int dim = 1 << n;
    for (int k = threadIdx.x; k < dim; k += blockDim.x) {
        const double2 a = psi[k];
#pragma unroll
        for (int q = 0; q < kSmallMaxN; ++q) {
            if (q < n && !((k >> q) & 1)) {
                const double2 b = psi[k | (1 << q)];
                ax[q] += a.x * b.x + a.y * b.y;
                ay[q] += a.x * b.y - a.y * b.x;
            }
        }
    }
    double zr = 0.0, zi = 0.0;
#pragma unroll
    for (int q = 0; q < kSmallMaxN; ++q) {
        zr += ax[q];
        zi += ay[q];
    }
    // (sum over q commutes with the sum over k: R only needs the total)
    zr = warp_sum(zr);
    zi = warp_sum(zi);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) {
        S.red[warp] = zr;
        S.red[32 + warp] = zi;
    }
    __syncthreads();
    double tr = 0.0, ti = 0.0;
    for (int wv = 0; wv < nw; ++wv) {
        tr += S.red[wv];
        ti += S.red[32 + wv];
    }
    __syncthreads();
    tr = 2.0 * tr / n;
    ti = 2.0 * ti / n;
    return sqrt(tr * tr + ti * ti);
}

__global__ void __launch_bounds__(kSmallThreads) small_run_kernel(const __grid_constant__ SmallParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmallShared& S = *reinterpret_cast<SmallShared*>(smem_raw);
    double2* psi = reinterpret_cast<double2*>(smem_raw + ((sizeof(SmallShared) + 127) / 128) * 128);
    const int n = P.n, dim = 1 << n;
    const int b = blockIdx.x;
    const double* K = P.K + (size_t)b * n * n;
    const double* om = P.omega + (size_t)b * n;

    if (threadIdx.x == 0) {
        int np = 0;
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) {
                const double k = 0.5 * (K[i * n + j] + K[j * n + i]);
                if (fabs(k) < kSparsityEps) continue;  // knm_hamiltonian.py:184
                S.kk[np] = k;
                S.pi[np] = (unsigned char)i;
                S.pj[np] = (unsigned char)j;
                ++np;
            }
        S.npairs = np;
        for (int i = 0; i < kSmallMaxN; ++i) S.w[i] = (i < n && fabs(om[i]) > kSparsityEps) ? om[i] : 0.0;
    }
    __syncthreads();
    // product state (xy_kuramoto.py:236-242)
    for (int k = threadIdx.x; k < dim; k += blockDim.x) {
        double v = 1.0;
        for (int q = 0; q < n; ++q) {
            double a = fmod(om[q], 2.0 * M_PI);
            if (a < 0.0) a += 2.0 * M_PI;
            double sn, cs;
            sincos(0.5 * a, &sn, &cs);
            v *= ((k >> q) & 1) ? sn : cs;
        }
        psi[k] = make_double2(v, 0.0);
    }
    __syncthreads();
    double r = small_order_parameter(psi, S, n);
    if (threadIdx.x == 0) P.R[(size_t)b * (P.m + 1)] = r;

    const int np = S.npairs;
    for (int s = 0; s < P.m; ++s) {
        const double tau = P.steps[s] / P.reps;
        for (int rep = 0; rep < P.reps; ++rep) {
            if (P.order == 1) {
                if (rep == 0) small_coeffs(S, tau);
                small_phase(psi, S, n, tau);
                for (int g = 0; g < np; ++g) small_gate(psi, n, S.pi[g], S.pj[g], S.cf[g][0], S.cf[g][1]);
            } else {
                for (int f = 0; f < P.nfrac; ++f) {
                    const double t = P.frac[f] * tau;
                    if (np == 0) {
                        small_phase(psi, S, n, t);
                        continue;
                    }
                    small_coeffs(S, t);
                    small_phase(psi, S, n, 0.5 * t);
                    for (int g = 0; g < np - 1; ++g) small_gate(psi, n, S.pi[g], S.pj[g], S.ch[g][0], S.ch[g][1]);
                    small_gate(psi, n, S.pi[np - 1], S.pj[np - 1], S.cf[np - 1][0], S.cf[np - 1][1]);
                    for (int g = np - 2; g >= 0; --g) small_gate(psi, n, S.pi[g], S.pj[g], S.ch[g][0], S.ch[g][1]);
                    small_phase(psi, S, n, 0.5 * t);
                }
            }
        }
        r = small_order_parameter(psi, S, n);
        if (threadIdx.x == 0) P.R[(size_t)b * (P.m + 1) + s + 1] = r;
    }
    if (P.final_states)
        for (int k = threadIdx.x; k < dim; k += blockDim.x) P.final_states[(size_t)b * dim + k] = psi[k];
}

inline void small_suzuki_fracs(int order, std::vector<double>& out) {
    if (order == 2) {
        out.push_back(1.0);
        return;
    }
    const double p = 1.0 / (4.0 - std::pow(4.0, 1.0 / (order - 1)));
    std::vector<double> inner;
    small_suzuki_fracs(order - 2, inner);
    const double w[5] = {p, p, 1.0 - 4.0 * p, p, p};
    for (double ww : w)
        for (double f : inner) out.push_back(ww * f);
}

// host launcher (host pointers in, host pointers out); returns an xyk status code
inline int small_batch_run(int device, int n, int B, const double* K, const double* omega, const double* step_sizes,
                           int m, int reps, int order, double* R_out, void* final_states, std::string& err) {
    if (n < 1 || n > kSmallMaxN) { err = "xyk_batch_run supports 1 <= n <= 12"; return 1; }
    if (B < 1 || m < 0 || reps < 1 || !K || !omega || !R_out || (m > 0 && !step_sizes)) { err = "bad batch arguments"; return 1; }
    if (!(order == 1 || order == 2 || order == 4 || order == 6)) { err = "order must be 1, 2, 4 or 6"; return 1; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return 2; }
    SmallParams P{};
    P.n = n; P.B = B; P.m = m; P.reps = reps; P.order = order;
    std::vector<double> fr;
    if (order >= 2) small_suzuki_fracs(order, fr);
    if ((int)fr.size() > kSmallMaxFrac) { err = "too many Suzuki sub-steps"; return 5; }
    P.nfrac = (int)fr.size();
    for (size_t i = 0; i < fr.size(); ++i) P.frac[i] = fr[i];
    const size_t dim = (size_t)1 << n;
    double *dK = nullptr, *dW = nullptr, *dS = nullptr, *dR = nullptr;
    double2* dF = nullptr;
    auto cleanup = [&]() {
        cudaFree(dK); cudaFree(dW); cudaFree(dS); cudaFree(dR); cudaFree(dF);
    };
    bool ok = cudaMalloc(&dK, sizeof(double) * B * n * n) == cudaSuccess &&
              cudaMalloc(&dW, sizeof(double) * B * n) == cudaSuccess &&
              cudaMalloc(&dS, sizeof(double) * std::max(m, 1)) == cudaSuccess &&
              cudaMalloc(&dR, sizeof(double) * B * (m + 1)) == cudaSuccess;
    if (ok && final_states) ok = cudaMalloc(&dF, sizeof(double2) * B * dim) == cudaSuccess;
    if (!ok) { err = std::string("batch allocation failed: ") + cudaGetErrorString(cudaGetLastError()); cleanup(); return 3; }
    cudaMemcpy(dK, K, sizeof(double) * B * n * n, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, omega, sizeof(double) * B * n, cudaMemcpyHostToDevice);
    if (m > 0) cudaMemcpy(dS, step_sizes, sizeof(double) * m, cudaMemcpyHostToDevice);
    P.K = dK; P.omega = dW; P.steps = dS; P.R = dR; P.final_states = dF;
    const size_t smem = ((sizeof(SmallShared) + 127) / 128) * 128 + dim * sizeof(double2);
    e = cudaFuncSetAttribute(small_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        small_run_kernel<<<B, kSmallThreads, smem>>>(P);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(R_out, dR, sizeof(double) * B * (m + 1), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && final_states)
        e = cudaMemcpy(final_states, dF, sizeof(double2) * B * dim, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) { err = std::string("batch run failed: ") + cudaGetErrorString(e); return 2; }
    return 0;
}

}  // namespace xyk
