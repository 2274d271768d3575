// CUDA kernels of the Kuramoto-XY statevector engine (sm_100a).
//
//  * gate-by-gate kernels (any layout): product-state init, diagonal phase, one pair rotation,
//    <X_q>/<Y_q>, <Z_q>, <XX+YY>, norm, bit-permuting copy;
//  * the fused tile pass: one read+write sweep over the state applying every pair rotation
//    the schedule compiler packed into it, tiles staged in shared memory, 2^RB amplitudes per
//    thread in registers, output written through a bit permutation.
//
// Reference arithmetic being replaced: Qiskit Statevector.evolve of the Trotter circuit
// (call site src/scpn_quantum_control/phase/xy_kuramoto.py:247) and the Rust observable loops
// (scpn_quantum_engine/src/pauli.rs:122-214, sectors.rs:105-145).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "xyk_internal.h"

namespace xyk {

template <typename real> struct cplx;
template <> struct cplx<double> { using type = double2; };
template <> struct cplx<float> { using type = float2; };

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
// tile-local index -> shared-memory slot (16-byte units) of the padded tile layout:
//     slot(x) = x + (x >> 3) + (x >> 6) + (x >> 9)
// The map is ADDITIVE over the index bits (slot(x) = sum of slot(1 << p) over the set bits p), so an
// address is "per-thread base + warp-uniform offset" and needs no per-access vector arithmetic
// (LDS [R + UR]), and the strides slot(1 << p) mod 8 run 1, 2, 4, 1, 2, 4, ...: any three lane bits on
// positions with three different p mod 3 spread eight lanes over the eight 16-byte columns of a
// 128-byte bank window (conflict free).  Cost: 14 % more shared memory than a dense tile.
__host__ __device__ constexpr uint32_t pad_slot(uint32_t x) { return x + (x >> 3) + (x >> 6) + (x >> 9); }
template <int TB>
__host__ __device__ constexpr uint32_t tile_bytes() {
    return (pad_slot((1u << TB) - 1u) + 1u) * 16u;
}

__host__ __device__ constexpr int insert_zero(int v, int pos) {
    return ((v >> pos) << (pos + 1)) | (v & ((1 << pos) - 1));
}

__device__ __forceinline__ uint64_t insert_zero64(uint64_t v, int pos) {
    return ((v >> pos) << (pos + 1)) | (v & ((1ull << pos) - 1ull));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of NV values per thread; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* >= NV*32 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; ++k) smem[k * 32 + warp] = v[k];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = lane < nw ? smem[k * 32 + lane] : 0.0;
            v[k] = warp_sum(x);
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// gate-by-gate kernels
// ---------------------------------------------------------------------------------------------
// psi0[k] = prod_q (bit_q ? sin : cos)(theta_q/2)     (xy_kuramoto.py:236-242)
// The product over the local physical bits comes from three host-built chunk tables (bits [0,b1), [b1,b2), [b2,nlocal));
// the factor of the sharded (rank) bits is folded into the last table.
template <typename real>
__global__ void init_product_kernel(typename cplx<real>::type* __restrict__ psi, const double* __restrict__ tab, int stride, int nlocal,
                                    int b1, int b2) {
    using C = typename cplx<real>::type;
    const uint64_t dim = 1ull << nlocal;
    const uint32_t m0 = (1u << b1) - 1u, m1 = (1u << (b2 - b1)) - 1u;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < dim; p += (uint64_t)gridDim.x * blockDim.x) {
        const double v = __ldg(&tab[(uint32_t)p & m0]) * __ldg(&tab[stride + ((uint32_t)(p >> b1) & m1)]) *
                         __ldg(&tab[2 * stride + (uint32_t)(p >> b2)]);
        C o;
        o.x = (real)v;
        o.y = (real)0;
        __stcs(psi + p, o);
    }
}

struct PhaseSpec {
    double w[kMaxQ];     // omega of the logical qubit on physical bit b (0 where dropped), b < nbits
    int32_t nbits;       // physical bits (local + global)
    int32_t nlocal;
    int32_t rank;
    int32_t b1, b2;      // chunk boundaries: [0,b1) [b1,b2) [b2,nlocal)
    double tau;
};

// tables[c][v] = exp(+i tau sum_{b in chunk c} w_b (1 - 2 bit_b(v)))   (A.3 step 1)
__global__ void phase_table_kernel(double2* __restrict__ tab, const __grid_constant__ PhaseSpec S, int stride) {
    const int c = blockIdx.y;
    const int lo = c == 0 ? 0 : (c == 1 ? S.b1 : S.b2);
    const int hi = c == 0 ? S.b1 : (c == 1 ? S.b2 : S.nlocal);
    const int cnt = 1 << (hi - lo);
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < cnt; v += gridDim.x * blockDim.x) {
        double ph = 0.0;
        for (int b = lo; b < hi; ++b) ph += S.w[b] * (1.0 - 2.0 * (double)((v >> (b - lo)) & 1));
        if (c == 2)  // global bits (fixed by the rank) ride on the last chunk
            for (int b = S.nlocal; b < S.nbits; ++b)
                ph += S.w[b] * (1.0 - 2.0 * (double)((S.rank >> (b - S.nlocal)) & 1));
        double sn, cs;
        sincos(S.tau * ph, &sn, &cs);
        tab[(size_t)c * stride + v] = make_double2(cs, sn);
    }
}

template <typename real>
__global__ void phase_apply_kernel(typename cplx<real>::type* __restrict__ psi, const double2* __restrict__ tab,
                                   int stride, int nlocal, int b1, int b2) {
    const uint64_t dim = 1ull << nlocal;
    const uint32_t m0 = (1u << b1) - 1u, m1 = (1u << (b2 - b1)) - 1u;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < dim; p += (uint64_t)gridDim.x * blockDim.x) {
        const double2 t0 = __ldg(&tab[(uint32_t)p & m0]);
        const double2 t1 = __ldg(&tab[stride + ((uint32_t)(p >> b1) & m1)]);
        const double2 t2 = __ldg(&tab[2 * stride + (uint32_t)(p >> b2)]);
        const double ar = t0.x * t1.x - t0.y * t1.y, ai = t0.x * t1.y + t0.y * t1.x;
        const double fr = ar * t2.x - ai * t2.y, fi = ar * t2.y + ai * t2.x;
        typename cplx<real>::type v = psi[p];
        const double xr = (double)v.x, xi = (double)v.y;
        v.x = (real)(xr * fr - xi * fi);
        v.y = (real)(xr * fi + xi * fr);
        psi[p] = v;
    }
}

// (a, b) = (psi[..1_pi..0_pj..], psi[..0_pi..1_pj..]) <- (c a + i s b, c b + i s a)   (A.3 step 2)
template <typename real>
__global__ void pair_gate_kernel(typename cplx<real>::type* __restrict__ psi, int nlocal, int pi, int pj, double c,
                                 double s) {
    const uint64_t quarter = 1ull << (nlocal - 2);
    const int lo = pi < pj ? pi : pj, hi = pi < pj ? pj : pi;
    const uint64_t mi = 1ull << pi, mj = 1ull << pj;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < quarter; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = insert_zero64(insert_zero64(t, lo), hi);
        typename cplx<real>::type a = psi[k | mi], b = psi[k | mj];
        const double ar = a.x, ai = a.y, br = b.x, bi = b.y;
        a.x = (real)(c * ar - s * bi);
        a.y = (real)(c * ai + s * br);
        b.x = (real)(c * br - s * ai);
        b.y = (real)(c * bi + s * ar);
        psi[k | mi] = a;
        psi[k | mj] = b;
    }
}

// Sharded qubit on rank bit b: <X_q> + i <Y_q> = 2 sum_{k} conj(psi^{(r)}_k) psi^{(r | 2^b)}_k over the ranks r with bit b clear --
// the partner's amplitudes are READ through the peer mapping (NVLink), no exchange.  Both ranks of a pair take half of the
// index range [k0, k1) each (the one with the bit set conjugates the PEER's amplitudes: mine_is_upper), so every rank reads
// half a shard per sharded qubit.  partial[blockIdx.x] = sum conj(lower_k) upper_k.
__global__ void __launch_bounds__(256) peer_dot_kernel(const double2* __restrict__ mine, const double2* __restrict__ peer, uint64_t k0, uint64_t k1,
                                                       int mine_is_upper, double2* __restrict__ partial) {
    __shared__ double red[2 * 32];
    double v[2] = {0.0, 0.0};
    for (uint64_t k = k0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < k1; k += (uint64_t)gridDim.x * blockDim.x) {
        const double2 m = mine[k];
        const double2 p = __ldcs(peer + k);   // streamed once over NVLink
        const double2 a = mine_is_upper ? p : m, b = mine_is_upper ? m : p;
        v[0] += a.x * b.x + a.y * b.y;
        v[1] += a.x * b.y - a.y * b.x;
    }
    block_sum<2>(v, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(v[0], v[1]);
}

// ---- quantum-jump (MCWF) helpers: phase/tensor_jump.py:157-206 ----
// kind 0: psi <- sigma_q psi with sigma = |1><0| on physical bit `pos` (the reference's "-" matrix [[0, 0], [1, 0]]: the
//         amplitude with the bit clear moves to its partner with the bit set, the clear half becomes 0);
// kind 1: psi <- Z_q psi.
template <typename real>
__global__ void jump_kernel(typename cplx<real>::type* __restrict__ psi, int nlocal, int pos, int kind) {
    const uint64_t half = 1ull << (nlocal - 1);
    const uint64_t m = 1ull << pos;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = insert_zero64(t, pos);
        if (kind == 0) {
            psi[k | m] = psi[k];
            psi[k].x = (real)0;
            psi[k].y = (real)0;
        } else {
            typename cplx<real>::type b = psi[k | m];
            b.x = -b.x;
            b.y = -b.y;
            psi[k | m] = b;
        }
    }
}

// ---- XXZ anisotropy: diagonal ZZ phases (bridge/knm_hamiltonian.py:196-201, term -K_ij delta Z_i Z_j) ----
// psi_k *= exp(+i theta z_i z_j), z = 1 - 2 bit: one pair, gate-by-gate engine (reference term order XX, YY, ZZ per pair)
template <typename real>
__global__ void zz_gate_kernel(typename cplx<real>::type* __restrict__ psi, int nlocal, int pi, int pj, double c, double s) {
    const uint64_t dim = 1ull << nlocal;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (uint64_t)gridDim.x * blockDim.x) {
        const double sg = (((k >> pi) ^ (k >> pj)) & 1) ? -s : s;
        typename cplx<real>::type v = psi[k];
        const double xr = (double)v.x, xi = (double)v.y;
        v.x = (real)(xr * c - xi * sg);
        v.y = (real)(xr * sg + xi * c);
        psi[k] = v;
    }
}

// All ZZ terms at once: psi_k *= exp(+i t sum_{a<b} J[la][lb] z_a z_b) over PHYSICAL bits a, b (la = log_of_bit[a]: the
// logical qubit on bit a; bits >= nlocal are fixed by the rank).  The bits split into Lo = [0, LB) and Hi = the rest:
//   theta(k) = Q_lo(k_lo) + Q_hi(k_hi) + sum_{a in Lo} z_a f_a(k_hi),   f_a = sum_{b in Hi} J_ab z_b
// Q_lo comes from a 2^LB-entry table (zz_low_table_kernel); a CTA takes one k_hi at a time, computes f_a and Q_hi, turns
// the linear part into two 64-entry factor tables in shared memory and multiplies its 2^LB amplitudes: one read and one
// write of the state, no trigonometry per amplitude.
struct ZZSpec {
    int nbits, nlocal, rank, LB;
    int log_of_bit[kMaxQ];
    double t;
};

__global__ void zz_low_table_kernel(double2* __restrict__ tab, const double* __restrict__ J, int ldJ, const __grid_constant__ ZZSpec S) {
    const int cnt = 1 << S.LB;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < cnt; v += gridDim.x * blockDim.x) {
        double q = 0.0;
        for (int a = 0; a < S.LB; ++a)
            for (int b = a + 1; b < S.LB; ++b) {
                const double jab = J[S.log_of_bit[a] * ldJ + S.log_of_bit[b]];
                q += (((v >> a) ^ (v >> b)) & 1) ? -jab : jab;
            }
        double sn, cs;
        sincos(S.t * q, &sn, &cs);
        tab[v] = make_double2(cs, sn);
    }
}

template <typename real>
__global__ void __launch_bounds__(256) zz_diag_kernel(typename cplx<real>::type* __restrict__ psi, const double2* __restrict__ tab,
                                                      const double* __restrict__ J, int ldJ, const __grid_constant__ ZZSpec S) {
    __shared__ double f[16];
    __shared__ double qhi;
    __shared__ double2 fa[64], fb[64];
    const int LB = S.LB, nhi = S.nbits - LB;
    const uint64_t nblk = 1ull << (S.nlocal - LB);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t kh = blockIdx.x; kh < nblk; kh += gridDim.x) {
        const uint64_t hibits = kh | ((uint64_t)S.rank << (S.nlocal - LB));  // bit (b - LB) of hibits = physical bit b
        if (warp == 0) {
            if (lane < LB) {
                double acc = 0.0;
                for (int b = 0; b < nhi; ++b) {
                    const double jab = J[S.log_of_bit[lane] * ldJ + S.log_of_bit[LB + b]];
                    acc += ((hibits >> b) & 1) ? -jab : jab;
                }
                f[lane] = acc;
            }
        } else if (warp == 1) {
            double acc = 0.0;
            for (int idx = lane; idx < nhi * nhi; idx += 32) {
                const int a = idx / nhi, b = idx - a * nhi;
                if (a < b) {
                    const double jab = J[S.log_of_bit[LB + a] * ldJ + S.log_of_bit[LB + b]];
                    acc += (((hibits >> a) ^ (hibits >> b)) & 1) ? -jab : jab;
                }
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) qhi = acc;
        }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int v = threadIdx.x & 63, half = threadIdx.x >> 6;  // half 0: bits [0, 6), half 1: bits [6, 12) (+ Q_hi)
            double ph = half ? qhi : 0.0;
            for (int b = 0; b < 6; ++b) {
                const int a = half * 6 + b;
                if (a < LB) ph += ((v >> b) & 1) ? -f[a] : f[a];
            }
            double sn, cs;
            sincos(S.t * ph, &sn, &cs);
            (half ? fb : fa)[v] = make_double2(cs, sn);
        }
        __syncthreads();
        typename cplx<real>::type* base = psi + (kh << LB);
        for (int x = threadIdx.x; x < (1 << LB); x += 256) {
            const double2 t0 = __ldg(&tab[x]), t1 = fa[x & 63], t2 = fb[x >> 6];
            const double ar = t0.x * t1.x - t0.y * t1.y, ai = t0.x * t1.y + t0.y * t1.x;
            const double fr = ar * t2.x - ai * t2.y, fi = ar * t2.y + ai * t2.x;
            typename cplx<real>::type v = base[x];
            const double xr = (double)v.x, xi = (double)v.y;
            v.x = (real)(xr * fr - xi * fi);
            v.y = (real)(xr * fi + xi * fr);
            base[x] = v;
        }
        __syncthreads();
    }
}

// <sum_{a<b} J_ab Z_a Z_b> = sum_k |psi_k|^2 q(k) with the quadratic form q of zz_diag_kernel (same Lo / Hi split: the low-bit
// part from a table of q values, the linear part from two 64-entry tables per k_hi): the ZZ part of the XXZ energy in ONE
// sweep instead of one sweep per pair.  partial[blockIdx.x] = this block's sum.
__global__ void zz_low_qtable_kernel(double* __restrict__ qtab, const double* __restrict__ J, int ldJ, const __grid_constant__ ZZSpec S) {
    const int cnt = 1 << S.LB;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < cnt; v += gridDim.x * blockDim.x) {
        double q = 0.0;
        for (int a = 0; a < S.LB; ++a)
            for (int b = a + 1; b < S.LB; ++b) {
                const double jab = J[S.log_of_bit[a] * ldJ + S.log_of_bit[b]];
                q += (((v >> a) ^ (v >> b)) & 1) ? -jab : jab;
            }
        qtab[v] = q;
    }
}

template <typename real>
__global__ void __launch_bounds__(256) zz_energy_kernel(const typename cplx<real>::type* __restrict__ psi, const double* __restrict__ qtab,
                                                        const double* __restrict__ J, int ldJ, const __grid_constant__ ZZSpec S,
                                                        double* __restrict__ partial) {
    __shared__ double f[16];
    __shared__ double qhi;
    __shared__ double la[64], lb[64];
    __shared__ double red[32];
    const int LB = S.LB, nhi = S.nbits - LB;
    const uint64_t nblk = 1ull << (S.nlocal - LB);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[1] = {0.0};
    for (uint64_t kh = blockIdx.x; kh < nblk; kh += gridDim.x) {
        const uint64_t hibits = kh | ((uint64_t)S.rank << (S.nlocal - LB));
        if (warp == 0) {
            if (lane < LB) {
                double a = 0.0;
                for (int b = 0; b < nhi; ++b) {
                    const double jab = J[S.log_of_bit[lane] * ldJ + S.log_of_bit[LB + b]];
                    a += ((hibits >> b) & 1) ? -jab : jab;
                }
                f[lane] = a;
            }
        } else if (warp == 1) {
            double a = 0.0;
            for (int idx = lane; idx < nhi * nhi; idx += 32) {
                const int x = idx / nhi, y = idx - x * nhi;
                if (x < y) {
                    const double jab = J[S.log_of_bit[LB + x] * ldJ + S.log_of_bit[LB + y]];
                    a += (((hibits >> x) ^ (hibits >> y)) & 1) ? -jab : jab;
                }
            }
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) qhi = a;
        }
        __syncthreads();
        if (threadIdx.x < 128) {
            const int v = threadIdx.x & 63, half = threadIdx.x >> 6;
            double ph = half ? qhi : 0.0;
            for (int b = 0; b < 6; ++b) {
                const int a = half * 6 + b;
                if (a < LB) ph += ((v >> b) & 1) ? -f[a] : f[a];
            }
            (half ? lb : la)[v] = ph;
        }
        __syncthreads();
        const typename cplx<real>::type* base = psi + (kh << LB);
        for (int x = threadIdx.x; x < (1 << LB); x += 256) {
            const typename cplx<real>::type v = base[x];
            const double pr = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
            acc[0] += pr * (__ldg(&qtab[x]) + la[x & 63] + lb[x >> 6]);
        }
        __syncthreads();
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc[0];
}

// ---- exact step by Chebyshev expansion on a fused H.phi sweep (reference twin: hardware/classical.py:201-208, expm_multiply) ----
// H = -sum_{i<j} K_ij (XX + YY + delta ZZ) - sum_i w_i Z_i acts on amplitude k as
//   (H phi)_k = d(k) phi_k - 2 sum_{(a,b): bits differ} K_ab phi_{k ^ (m_a | m_b)},   d(k) = -sum_a w_a z_a - delta sum K_ab z_a z_b
// (scpn_quantum_engine/src/hamiltonian.rs:55-78).  The pair terms are SUMMED, so there is no ordering constraint: one
// kernel applies all of them.  A warp's 32 consecutive k gather 32 consecutive (or lane-permuted) amplitudes per pair:
// every gather is one coalesced 512-byte read that the L2 serves while the state fits it.
struct HPair {
    unsigned long long ma, mb;  // bit masks of the two qubits in the current layout
    double k, kz;               // K_ab and K_ab * delta
};

// next = alpha (H cur) / E - prev (prev == nullptr: no subtraction);  acc = (first ? 0 : acc) ... see flags:
//   acc_mode 0: acc += coef * next;  1: acc = coef0 * cur + coef * next (first term of the series)
__global__ void __launch_bounds__(256) cheb_step_kernel(const double2* __restrict__ cur, const double2* prev, double2* next,
                                                        double2* __restrict__ acc, const HPair* __restrict__ pairs, int npairs,
                                                        const double* __restrict__ wbit, int nbits, uint64_t dim, double alpha_over_E,
                                                        double2 coef, double2 coef0, int acc_mode) {
    extern __shared__ unsigned char cheb_smem[];
    HPair* sp = reinterpret_cast<HPair*>(cheb_smem);
    double* sw = reinterpret_cast<double*>(sp + npairs);
    for (int t = threadIdx.x; t < npairs; t += blockDim.x) sp[t] = pairs[t];
    for (int t = threadIdx.x; t < nbits; t += blockDim.x) sw[t] = wbit[t];
    __syncthreads();
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < dim; k += (uint64_t)gridDim.x * blockDim.x) {
        double diag = 0.0;
        for (int b = 0; b < nbits; ++b) diag += ((k >> b) & 1) ? sw[b] : -sw[b];
        double hr = 0.0, hi = 0.0;
        for (int t = 0; t < npairs; ++t) {
            const HPair pr = sp[t];
            const bool differ = ((k & pr.ma) != 0) != ((k & pr.mb) != 0);
            if (differ) {
                const double2 v = __ldg(&cur[k ^ (pr.ma | pr.mb)]);
                hr -= 2.0 * pr.k * v.x;
                hi -= 2.0 * pr.k * v.y;
                diag += pr.kz;
            } else {
                diag -= pr.kz;
            }
        }
        const double2 c = cur[k];
        hr += diag * c.x;
        hi += diag * c.y;
        double nr = alpha_over_E * hr, ni = alpha_over_E * hi;
        if (prev) {
            const double2 p = prev[k];
            nr -= p.x;
            ni -= p.y;
        }
        next[k] = make_double2(nr, ni);
        double ar = coef.x * nr - coef.y * ni, ai = coef.x * ni + coef.y * nr;
        if (acc_mode == 1) {
            ar += coef0.x * c.x - coef0.y * c.y;
            ai += coef0.x * c.y + coef0.y * c.x;
        } else {
            const double2 a = acc[k];
            ar += a.x;
            ai += a.y;
        }
        acc[k] = make_double2(ar, ai);
    }
}

// (a, b) = (psi[..0_pos..], psi[..1_pos..]) <- (c a - s b, s a + c b): Ry(theta) with c = cos(theta/2), s = sin(theta/2)
template <typename real>
__global__ void ry_gate_kernel(typename cplx<real>::type* __restrict__ psi, int nlocal, int pos, double c, double s) {
    const uint64_t half = 1ull << (nlocal - 1);
    const uint64_t m = 1ull << pos;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = insert_zero64(t, pos);
        typename cplx<real>::type a = psi[k], b = psi[k | m];
        const double ar = a.x, ai = a.y, br = b.x, bi = b.y;
        a.x = (real)(c * ar - s * br);
        a.y = (real)(c * ai - s * bi);
        b.x = (real)(s * ar + c * br);
        b.y = (real)(s * ai + c * bi);
        psi[k] = a;
        psi[k | m] = b;
    }
}

// partial[blockIdx.x] = (sum_k Re conj(psi_k) psi_{k|m}, sum_k Im ...) over k with bit `pos` clear
// (scpn_quantum_engine/src/pauli.rs:195-211, without the factor 2)
template <typename real>
__global__ void xy_expect_kernel(const typename cplx<real>::type* __restrict__ psi, int nlocal, int pos,
                                 double2* __restrict__ partial) {
    __shared__ double red[2 * 32];
    const uint64_t half = 1ull << (nlocal - 1);
    const uint64_t m = 1ull << pos;
    double acc[2] = {0.0, 0.0};
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < half; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = insert_zero64(t, pos);
        const typename cplx<real>::type a = psi[k], b = psi[k | m];
        acc[0] += (double)a.x * (double)b.x + (double)a.y * (double)b.y;
        acc[1] += (double)a.x * (double)b.y - (double)a.y * (double)b.x;
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = make_double2(acc[0], acc[1]);
}

// <X>, <Y> of M qubits in ONE sweep: each thread holds the 2^M amplitudes spanned by the M physical
// bits pos[0..M) (ascending) of one base index and accumulates, for every m, the pair products over
// the 2^(M-1) pairs that differ in bit m.  partial[(qidx[m]) * stride + blockIdx.x] = (Re, Im) sums.
struct ObsSpec {
    int32_t pos[8];    // physical bits, ascending
    int32_t qidx[8];   // row of `partial` each one reports to
    int32_t nlocal;
    int32_t stride;
};

template <typename real, int M>
__global__ void __launch_bounds__(256) xy_expect_multi_kernel(const typename cplx<real>::type* __restrict__ psi,
                                                              const __grid_constant__ ObsSpec S,
                                                              double2* __restrict__ partial) {
    constexpr int NA = 1 << M;
    __shared__ double red[2 * M * 32];
    const uint64_t nbase = 1ull << (S.nlocal - M);
    double acc[2 * M];
#pragma unroll
    for (int k = 0; k < 2 * M; ++k) acc[k] = 0.0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbase; b += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k0 = b;
#pragma unroll
        for (int m = 0; m < M; ++m) k0 = insert_zero64(k0, S.pos[m]);
        double ar[NA], ai[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            uint64_t off = 0;
#pragma unroll
            for (int m = 0; m < M; ++m)
                if ((j >> m) & 1) off |= 1ull << S.pos[m];
            const typename cplx<real>::type v = psi[k0 | off];
            ar[j] = (double)v.x;
            ai[j] = (double)v.y;
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                if (!((j >> m) & 1)) {
                    const int f = j | (1 << m);
                    acc[2 * m] += ar[j] * ar[f] + ai[j] * ai[f];
                    acc[2 * m + 1] += ar[j] * ai[f] - ai[j] * ar[f];
                }
            }
        }
    }
    block_sum<2 * M>(acc, red);
    if (threadIdx.x == 0)
#pragma unroll
        for (int m = 0; m < M; ++m)
            partial[(size_t)S.qidx[m] * S.stride + blockIdx.x] = make_double2(acc[2 * m], acc[2 * m + 1]);
}

// partial[blockIdx.x][b] = sum_k (1 - 2 bit_b(k)) |psi_k|^2 for every physical bit b < nlocal,
// partial[blockIdx.x][nlocal] = sum_k |psi_k|^2     (pauli.rs:163-170)
template <typename real>
__global__ void z_expect_kernel(const typename cplx<real>::type* __restrict__ psi, int nlocal,
                                double* __restrict__ partial /* [grid][kMaxQ+1] */) {
    // The grid is a power of two (2^T threads in all): the low T index bits of every amplitude a thread reads are its global
    // thread id, so only the nlocal - T high bits (the iteration counter) need per-amplitude bookkeeping -- at most kZHi
    // conditional adds per amplitude instead of one per qubit.
    constexpr int kZHi = 16;
    __shared__ double red[32];
    const uint64_t dim = 1ull << nlocal;
    const uint32_t nthreads = gridDim.x * blockDim.x;   // power of two (host)
    const int T = 31 - __clz(nthreads);
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t iters = nthreads >= dim ? 1 : (dim >> T);
    double hi1[kZHi];   // weight with high bit T + b set
#pragma unroll
    for (int b = 0; b < kZHi; ++b) hi1[b] = 0.0;
    double tot = 0.0;
    if (gtid < dim) {
        for (uint64_t it = 0; it < iters; ++it) {
            const typename cplx<real>::type a = psi[(it << T) + gtid];
            const double pr = (double)a.x * (double)a.x + (double)a.y * (double)a.y;
            tot += pr;
#pragma unroll
            for (int b = 0; b < kZHi; ++b) hi1[b] += ((it >> b) & 1ull) ? pr : 0.0;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int b = 0; b <= kMaxQ; ++b) {
        double v;
        if (b == kMaxQ) v = tot;
        else if (b < T) v = ((gtid >> b) & 1u) ? tot : 0.0;
        else {
            v = 0.0;
#pragma unroll
            for (int k = 0; k < kZHi; ++k)
                if (k == b - T) v = hi1[k];
        }
        v = warp_sum(v);
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double x = lane < nw ? red[lane] : 0.0;
            x = warp_sum(x);
            if (lane == 0) partial[(size_t)blockIdx.x * (kMaxQ + 1) + b] = x;
        }
        __syncthreads();
    }
}

// partial[blockIdx.x] = sum over k with bit pi = 1, bit pj = 0 of Re conj(psi_k) psi_{k ^ (mi|mj)}
// (sectors.rs:105-145: C_ij = 4 * that sum)
template <typename real>
__global__ void xxyy_expect_kernel(const typename cplx<real>::type* __restrict__ psi, int nlocal, int pi, int pj,
                                   double* __restrict__ partial) {
    __shared__ double red[32];
    const uint64_t quarter = 1ull << (nlocal - 2);
    const int lo = pi < pj ? pi : pj, hi = pi < pj ? pj : pi;
    const uint64_t mi = 1ull << pi, mj = 1ull << pj;
    double acc[1] = {0.0};
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < quarter; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = insert_zero64(insert_zero64(t, lo), hi);
        const typename cplx<real>::type a = psi[k | mi], b = psi[k | mj];
        acc[0] += (double)a.x * (double)b.x + (double)a.y * (double)b.y;
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc[0];
}

// out[j] = sum_b partial[b * stride + j]   (deterministic second reduction stage; one block per j)
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nblocks, int stride,
                                       double* __restrict__ out) {
    __shared__ double red[32];
    const int j = blockIdx.x;
    double acc[1] = {0.0};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc[0] += partial[(size_t)b * stride + j];
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) out[j] = acc[0];
}

// out[t] = sum_b partial[t * stride + b]   (row layout; one block per t)
__global__ void reduce_rows_kernel(const double* __restrict__ partial, int nblocks, int stride,
                                   double* __restrict__ out) {
    __shared__ double red[32];
    const int t = blockIdx.x;
    double acc[1] = {0.0};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc[0] += partial[(size_t)t * stride + b];
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) out[t] = acc[0];
}

// out[2q] = <X_q> = 2 sum_b partial[q*stride+b].x, out[2q+1] = <Y_q>   (one block per qubit)
__global__ void xy_finalize_kernel(const double2* __restrict__ partial, int nblocks, int stride,
                                   double* __restrict__ out) {
    __shared__ double red[2 * 32];
    const int q = blockIdx.x;
    double acc[2] = {0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        const double2 v = partial[(size_t)q * stride + b];
        acc[0] += v.x;
        acc[1] += v.y;
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
        out[2 * q] = 2.0 * acc[0];
        out[2 * q + 1] = 2.0 * acc[1];
    }
}

// R = |sum_q (<X_q> + i <Y_q>)| / n from the finalized expectations (xy[2q], xy[2q+1]); summed in qubit order by one thread,
// exactly like the host loop of xyk_order_parameter (reference: phase/xy_kuramoto.py:173,183-184)
__global__ void order_param_kernel(const double* __restrict__ xy, int n, double* __restrict__ R) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double zr = 0.0, zi = 0.0;
        for (int q = 0; q < n; ++q) {
            zr += xy[2 * q];
            zi += xy[2 * q + 1];
        }
        zr /= n;
        zi /= n;
        *R = sqrt(zr * zr + zi * zi);
    }
}

struct PermSpec {
    uint8_t dst_of_src[kMaxQ];  // source physical bit b -> destination physical bit
    int32_t nlocal;
};

// dst[perm(p)] = src[p].  Gather form: the thread index runs over the DESTINATION so that
// writes are coalesced.
template <typename real>
__global__ void permute_kernel(const typename cplx<real>::type* __restrict__ src,
                               typename cplx<real>::type* __restrict__ dst, const __grid_constant__ PermSpec S) {
    const uint64_t dim = 1ull << S.nlocal;
    for (uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; d < dim; d += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t s = 0;
        for (int b = 0; b < S.nlocal; ++b) s |= ((d >> S.dst_of_src[b]) & 1ull) << b;
        dst[d] = src[s];
    }
}

template <typename from, typename to>
__global__ void convert_kernel(const typename cplx<from>::type* __restrict__ src,
                               typename cplx<to>::type* __restrict__ dst, uint64_t count) {
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < count; p += (uint64_t)gridDim.x * blockDim.x) {
        typename cplx<to>::type o;
        o.x = (to)src[p].x;
        o.y = (to)src[p].y;
        dst[p] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// fused tile pass
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRounds = 14;
constexpr int kPhTabLo = 128, kPhTabBits = 10, kPhTabHi = kPhTabLo + (1 << kPhTabBits), kPhTabSize = kPhTabHi + (1 << kPhTabBits);

constexpr uint32_t kPassFuseLoad = 1u;   // round 0 reads its amplitudes straight from global memory
constexpr uint32_t kPassFuseStore = 2u;  // the last round writes its amplitudes straight to global memory
constexpr uint32_t kPassHasPhase = 4u;   // diagonal phase folded into round 0
constexpr uint32_t kPassNoPrefetch = 8u; // do not pull the next tile into L2 ahead of time
constexpr uint32_t kPassPfLate = 16u;     // experiment: prefetch the next tile at the start of the LAST round
constexpr uint32_t kPassPfEvictLast = 32u; // experiment: prefetch with an L2 evict_last policy
constexpr uint32_t kPassPlainLd = 64u;    // experiment: plain ld.global instead of ld.global.cs

template <int TB, int RB>
struct PassDesc {
    static constexpr int NT = TB - RB;
    static constexpr int NSLOT = RB * (RB - 1) / 2;
    struct Round {
        uint8_t thrpos[8];   // tile-local position of thread bit b (NT used); split rounds: of thread bit b + 1 (NT - 1 used)
        uint8_t regpos[8];   // tile-local position of register bit k (RB used; split rounds: RB + 1)
        uint16_t regsw[8];   // shared-memory byte stride of register bit k: 16 * pad_slot(1 << regpos[k])
        uint32_t mask;       // enabled pair slots
        uint16_t cta_sync;   // 1: block-wide barrier after this round's exchange; 0: warp-level barrier suffices
        uint16_t split;      // 0: complex round; 1, 2, 3: split round (RB + 1 register qubits, one real component per amplitude;
                             // see PlanRound) whose left group has that many qubits (the right group sits on the high register bits)
        double ca[NSLOT];    // shear form: -tan(phi/2); standard form: cos(phi)
        double cb[NSLOT];    // sin(phi)
    };
    int32_t nq;              // local qubits
    int32_t nrounds;
    uint32_t flags;
    int32_t pad0;
    uint8_t hi_out[kMaxQ];   // output bit of input bit TB + b (the tile-id bits)
    uint8_t st_lsrc[16];     // unfused store: tile-local position of the tile bit with the m-th smallest output bit
    uint8_t st_odst[16];     //                that output bit
    uint8_t last_thr_out[8]; // fused store: output bit of the qubit on thread bit b (split: b + 1) of the LAST round
    uint8_t last_reg_out[8]; //              output bit of the qubit on register bit k of the LAST round
    uint32_t st_reg_idx[8];  // fused store: output INDEX stride (1 << output bit) of the LAST round's register bit k
    uint32_t arr_reg[8];     // fused load: byte stride of round 0's register bit k in the arrival layout (arrival_stride_bytes)
    uint64_t out_base;       // peer passes: output-index contribution of this device's rank bits
    // fused phase exp(+i tau sum_b w_b (1 - 2 bit_b)) over all input bits = product of four table factors (built on the host):
    //   ph_tab[thread id]                      round 0's thread qubits and the sharded (rank) bits
    //   ph_tab[kPhTabLo + (tile & mask)]       the low ph_b1 tile-id bits
    //   ph_tab[kPhTabHi + (tile >> ph_b1)]     the other tile-id bits
    //   ph_reg[register index]                 round 0's register qubits
    const double2* ph_tab;
    int32_t ph_b1;
    int32_t pad1;
    double2 ph_reg[1 << RB];
    Round rounds[kMaxRounds];
};

template <bool LIFT>
__device__ __forceinline__ void rotate2(double& x, double& y, double ca, double cb) {
    if constexpr (LIFT) {  // three shears: R(phi) = Sx(-tan(phi/2)) Sy(sin phi) Sx(-tan(phi/2))
        x = fma(ca, y, x);
        y = fma(cb, x, y);
        x = fma(ca, y, x);
    } else {
        const double nx = fma(ca, x, -cb * y);
        const double ny = fma(cb, x, ca * y);
        x = nx;
        y = ny;
    }
}

// One pass = one read + one write of the state.  Per tile (2^TB consecutive amplitudes of the
// input buffer): round 0 pulls 2^RB amplitudes per thread (straight from global memory when the
// schedule allows a coalesced mapping, else through a cp.async staged tile), every round applies
// the enabled rotations among its RB register qubits and exchanges through swizzled shared
// memory, and the last round scatters the result through the output bit permutation (straight
// from registers when the schedule allows, else through shared memory).
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk_evict_last(const void* gptr, uint32_t bytes) {
    asm volatile(
        "{ .reg .b64 pol; createpolicy.fractional.L2::evict_last.b64 pol, 1.0;\n"
        "cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, pol; }\n" ::"l"(gptr), "r"(bytes)
        : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const double2& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// Output buffers of every rank of a sharded state (mapped through CUDA IPC): a pass whose store is
// fused with the qubit exchange writes each amplitude into the rank selected by the output-index
// bits >= n_local (st.global on peer memory over NVLink) -- compute and collective in one kernel.
struct PeerPtrs {
    double2* p[8];
};

template <bool PEER, bool C64 = false>
__device__ __forceinline__ void store_out(double2* __restrict__ out, const PeerPtrs& peers, int nl, uint64_t addr,
                                          const double2& v) {
    if constexpr (C64) {  // complex64 state: 8-byte amplitudes in global memory, FP64 everywhere on chip
        static_assert(!PEER, "complex64 states are not sharded");
        __stcs(reinterpret_cast<float2*>(out) + addr, make_float2((float)v.x, (float)v.y));
    } else if constexpr (PEER) {
        double2* base = peers.p[addr >> nl];
        __stcs(base + (addr & ((1ull << nl) - 1ull)), v);
    } else {
        __stcs(out + addr, v);
    }
}

// Shared-memory addresses / thread id through volatile asm: ptxas otherwise rematerialises them
// (S2R SR_TID, S2UR SR_CgaCtaId) inside every round to save registers, which puts their latency on
// the critical path of a kernel that runs only three warps per scheduler.
__device__ __forceinline__ uint32_t smem_addr_once(const void* p) {
    uint32_t a;
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }\n" : "=r"(a) : "l"(p));
    return a;
}
__device__ __forceinline__ uint32_t tid_once() {
    uint32_t t;
    asm volatile("mov.u32 %0, %%tid.x;\n" : "=r"(t));
    return t;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("{ .reg .u16 h; ld.shared.u16 h, [%1]; cvt.u32.u16 %0, h; }\n" : "=r"(v) : "r"(addr));
    return v;
}

// ---- bulk asynchronous tile load (TMA, cp.async.bulk) signalled through an mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(mbar),
        "r"(parity)
        : "memory");
}
// Arrival layout of a bulk-loaded tile: the 2^TB amplitudes arrive as 2^(TB-9) segments of 8 KiB, segment s at
// byte offset s * (8 KiB + 16): slot0(x) = x + (x >> 9).  Additive like pad_slot, and the 16-byte-column
// strides of positions 0, 1, 2 AND of positions 9, 10, 11 are 1, 2, 4 (mod 8): round 0 reads the tile
// bank-conflict free when its three low lane bits sit on one position of each pair {0, 9}, {1, 10}, {2, 11}
// -- the schedule compiler parks three of the round's thread qubits on positions 9..11, so the first round
// is free to hold ANY tile qubits in registers (a plain linear tile would force positions 0..2 onto lanes).
__host__ __device__ constexpr uint32_t arrival_stride_bytes(int p) { return p < 9 ? (16u << p) : (16u << p) + (16u << (p - 9)); }
template <int TB>
constexpr uint32_t arrival_bytes() {
    return (16u << TB) + (TB > 9 ? ((1u << (TB - 9)) - 1u) * 16u : 0u);
}
// one thread: the generic-proxy reads of the buffer are ordered before the async-proxy writes, the
// mbarrier expects the whole tile, every segment completes on it
template <int TB>
__device__ __forceinline__ void bulk_load_tile(uint32_t smem_dst, const double2* gsrc, uint32_t mbar) {
    constexpr uint32_t kSeg = TB > 9 ? (16u << 9) : (16u << TB);
    constexpr int kNumSeg = TB > 9 ? (1 << (TB - 9)) : 1;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(kSeg * kNumSeg) : "memory");
#pragma unroll
    for (int sgm = 0; sgm < kNumSeg; ++sgm)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst + sgm * (kSeg + 16u)),
                     "l"(reinterpret_cast<const char*>(gsrc) + (size_t)sgm * kSeg), "r"(kSeg), "r"(mbar)
                     : "memory");
}

// complex64 state: the tile's 2^TB float2 amplitudes (8 << TB bytes, contiguous) arrive at the FRONT of the tile buffer; the
// group then widens them in place into the FP64 arrival layout (widen_tile_c64) before round 0 reads the tile.
template <int TB>
__device__ __forceinline__ void bulk_load_tile_c64(uint32_t smem_dst, const char* gsrc, uint32_t mbar) {
    constexpr uint32_t kBytes = 8u << TB, kSeg = kBytes < 8192u ? kBytes : 8192u;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(kBytes) : "memory");
#pragma unroll
    for (uint32_t off = 0; off < kBytes; off += kSeg)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst + off),
                     "l"(gsrc + off), "r"(kSeg), "r"(mbar)
                     : "memory");
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// The per-round descriptors are copied from the kernel parameter (constant bank) into shared memory
// once per CTA: a dynamically indexed constant load that misses costs several hundred cycles, and the
// compiler issues it right at its first use; a shared-memory copy is fetched with explicit loads one
// step ahead of its use (next round's header while this round stores, next gate's coefficients while
// this gate is applied).
// one real component (comp = 0: Re, 1: Im) of the amplitude at output index `addr`
template <bool PEER>
__device__ __forceinline__ void store_out_comp(double2* __restrict__ out, const PeerPtrs& peers, int nl, uint64_t addr, uint32_t comp,
                                               double v) {
    if constexpr (PEER) {
        double* base = reinterpret_cast<double*>(peers.p[addr >> nl]);
        __stcs(base + 2ull * (addr & ((1ull << nl) - 1ull)) + comp, v);
    } else {
        __stcs(reinterpret_cast<double*>(out) + 2ull * addr + comp, v);
    }
}
// v with its sign bit XORed with `signmask` (0 or 0x80000000): a negation on the integer pipe
__device__ __forceinline__ double flip_sign(double v, uint32_t signmask) {
    return __hiloint2double(__double2hiint(v) ^ (int)signmask, __double2loint(v));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}

// Split rounds, NL = size of the left group (register bits [0, NL)), right group on bits [NL, RS).
// role(j) = parity of the right-group bits of j: the thread of half h holds component role ^ h of
// amplitude j, negated when role = 1 and h = 1.  Addresses: eight per-thread bases over the three high
// register bits (component chosen by the parity of those bits) + eight warp-uniform offsets over the
// three low bits; elements whose LOW bits carry an odd share of the role (NL < 3 only) are served in a
// second sweep after flipping the component bit (^ 8) of the eight bases.
template <int NL, int RS, int NV, bool STORE>
__device__ __forceinline__ void split_exchange(double (&v)[NV], uint32_t base, uint32_t half, const uint32_t (&st)[RS], uint32_t smask) {
    static_assert(RS == 6 && NV == 64, "address split: three high + three low register bits");
    const uint32_t base1 = base + 8u - (half << 4);  // the other real component of the same amplitude slot
    uint32_t bh[8], lo[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t hi = 0, l = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if ((c >> k) & 1) hi += st[3 + k];
            if ((c >> k) & 1) l += st[k];
        }
        bh[c] = ((__builtin_popcount(c) & 1) ? base1 : base) + hi;
        lo[c] = l;
    }
#pragma unroll
    for (int sweep = 0; sweep < (NL < 3 ? 2 : 1); ++sweep) {
        if (sweep == 1) {  // elements whose low bits flip the role: the other component
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                uint32_t hi = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if ((c >> k) & 1) hi += st[3 + k];
                bh[c] = ((__builtin_popcount(c) & 1) ? base : base1) + hi;
            }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if ((__builtin_popcount((j & 7) >> NL) & 1) != sweep) continue;
            const bool fj = __builtin_popcount(j >> NL) & 1;
            if constexpr (STORE) {
                sts64(bh[j >> 3] + lo[j & 7], fj ? flip_sign(v[j], smask) : v[j]);
            } else {
                v[j] = lds64(bh[j >> 3] + lo[j & 7]);
                if (fj) v[j] = flip_sign(v[j], smask);
            }
        }
    }
}

template <int NL, int RS, int NV, bool LIFT, typename Round>
__device__ __forceinline__ void split_gates(double (&v)[NV], uint32_t mask, const Round& R) {
#pragma unroll
    for (int x = 0; x < NL; ++x) {
#pragma unroll
        for (int y = NL; y < RS; ++y) {
            const int slot = (RS - NL) * x + (y - NL);
            if (mask & (1u << slot)) {
                const double ca = R.ca[slot], cb = R.cb[slot];
#pragma unroll
                for (int m = 0; m < NV / 4; ++m) {
                    const int j0 = insert_zero(insert_zero(m, x), y);
                    const int jA = j0 | (1 << x), jB = j0 | (1 << y);
                    // half 0 holds Re(A), Im(B) when role(jA) = 0 and Im(A), Re(B) otherwise (roles swap)
                    if (__builtin_popcount(jA >> NL) & 1) rotate2<LIFT>(v[jB], v[jA], ca, cb);
                    else rotate2<LIFT>(v[jA], v[jB], ca, cb);
                }
            }
        }
    }
}

template <int NL, int RS, int NV, bool PEER, bool C64 = false>
__device__ __forceinline__ void split_global_store(const double (&v)[NV], double2* __restrict__ out, const PeerPtrs& peers, int nq,
                                                   uint64_t dbase, const uint8_t* last_reg_out, uint32_t half, uint32_t smask) {
    if constexpr (C64) {
        static_assert(!PEER, "complex64 states are not sharded");
        char* pb = reinterpret_cast<char*>(out) + (dbase << 3);
        char* pb0 = pb + (half << 2);      // elements with role 0: component `half`
        char* pb1 = pb + 4 - (half << 2);  // elements with role 1: the other component
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            uint64_t off = 0;
#pragma unroll
            for (int k = 0; k < RS; ++k)
                if ((j >> k) & 1) off += 8ull << last_reg_out[k];
            const bool fj = __builtin_popcount(j >> NL) & 1;
            __stcs(reinterpret_cast<float*>((fj ? pb1 : pb0) + off), (float)(fj ? flip_sign(v[j], smask) : v[j]));
        }
    } else if constexpr (PEER) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            uint64_t off = 0;
#pragma unroll
            for (int k = 0; k < RS; ++k)
                if ((j >> k) & 1) off += 1ull << last_reg_out[k];
            const uint32_t fj = __builtin_popcount(j >> NL) & 1;
            store_out_comp<PEER>(out, peers, nq, dbase + off, fj ^ half, fj ? flip_sign(v[j], smask) : v[j]);
        }
    } else {
        // per-thread byte pointer once per tile + warp-uniform byte offsets of the register bits
        char* pb = reinterpret_cast<char*>(out) + (dbase << 4);
        char* pb0 = pb + (half << 3);      // elements with role 0: component `half`
        char* pb1 = pb + 8 - (half << 3);  // elements with role 1: the other component
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            uint64_t off = 0;
#pragma unroll
            for (int k = 0; k < RS; ++k)
                if ((j >> k) & 1) off += 16ull << last_reg_out[k];
            const bool fj = __builtin_popcount(j >> NL) & 1;
            __stcs(reinterpret_cast<double*>((fj ? pb1 : pb0) + off), fj ? flip_sign(v[j], smask) : v[j]);
        }
    }
}

// ---- development only (-DXYK_TRACE, tools/trace build): per-warp phase timeline of the pass kernel ----
#ifdef XYK_TRACE
__device__ unsigned long long* xyk_trace_buf = nullptr;   // [warp slot][kTraceMaxE]
__device__ int xyk_trace_k0 = 100, xyk_trace_nk = 4;     // traced window: the k0-th .. (k0+nk-1)-th tile of every CTA
constexpr int kTraceMaxE = 256;
__device__ __forceinline__ unsigned long long trace_clock() {
    unsigned long long c;
    asm volatile("mov.u64 %0, %%clock64;\n" : "=l"(c)::"memory");
    return c;
}
// consume `x` (a shared-memory store of its high word) so that the following clock read issues after x is available
__device__ __forceinline__ void trace_consume(uint32_t scratch, double x) {
    asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(scratch), "r"(__double2hiint(x)) : "memory");
}
#define XYK_TR(code)                                                                  \
    do {                                                                              \
        if (tr_on && tr_i < kTraceMaxE) tr_p[tr_i++] = (trace_clock() << 8) | (unsigned long long)(code); \
    } while (0)
#define XYK_TR_DEP(code, x)                      \
    do {                                         \
        if (tr_on) {                             \
            trace_consume(tr_scratch, (x));      \
            XYK_TR(code);                        \
        }                                        \
    } while (0)
#else
#define XYK_TR(code) do { } while (0)
#define XYK_TR_DEP(code, x) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// <X_q>, <Y_q> of up to twelve qubits per sweep (complex128): tiles of 2^12 amplitudes -- a contiguous chunk of 2^nb amplitudes
// times 12 - nb further physical bits -- arrive in shared memory by bulk asynchronous copies (arrival layout); per tile up to
// three register rounds read 32 amplitudes per thread and accumulate conj(psi_k) psi_{k|m} for their register bits in FP64.
// Reference semantics: scpn_quantum_engine/src/pauli.rs:195-211 (one sweep PER QUBIT there).
// ---------------------------------------------------------------------------------------------
struct ObsTileSpec {
    int32_t nlocal, nb, nrounds, stride;
    uint8_t tile_pos[12];        // physical bit of tile-local position p (the first nb are 0 .. nb-1)
    struct Round {
        uint8_t regpos[4];       // tile-local position of register bit k
        uint8_t thrpos[8];       // tile-local position of thread bit b (the three lowest: one of each pair {0,9}, {1,10}, {2,11});
                                 // thrpos[7]: the bit the two batches of a round differ in
        int16_t row[4];          // row of `partial` register bit k reports to, -1: not measured in this round
    } rounds[3];
};

__global__ void __launch_bounds__(128, 3)
xy_expect_tile_kernel(const double2* __restrict__ psi, const __grid_constant__ ObsTileSpec S, double2* __restrict__ partial) {
    constexpr int TB = 12, RB = 4, NREG = 1 << RB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar_mem;
    __shared__ double red[2 * 32];
    const uint32_t tile_u32 = smem_addr_once(smem_raw), mbar = smem_addr_once(&mbar_mem);
    const uint32_t tid = threadIdx.x;
    const uint32_t ntiles = 1u << (S.nlocal - TB);
    const int nhi = TB - S.nb;                       // strided tile bits
    const uint32_t chunk_bytes = 16u << S.nb;
    double acc[3][2 * RB];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 2 * RB; ++k) acc[r][k] = 0.0;
    uint32_t tbase[3];   // per round: byte offset of this thread's base amplitude in the arrival layout
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        tbase[r] = 0;
#pragma unroll
        for (int b = 0; b < 7; ++b)
            if ((tid >> b) & 1u) tbase[r] += arrival_stride_bytes(S.rounds[r].thrpos[b]);
    }
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    auto issue = [&](uint32_t t) {
        uint64_t base = (uint64_t)t << S.nb;
        for (int i = 0; i < nhi; ++i) base = insert_zero64(base, S.tile_pos[S.nb + i]);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(16u << TB) : "memory");
        for (uint32_t ch = tid; ch < (1u << nhi); ch += 128u) {   // chunk index = tile-local index >> nb
            uint64_t off = 0;
            for (int i = 0; i < nhi; ++i)
                if ((ch >> i) & 1u) off |= 1ull << S.tile_pos[S.nb + i];
            const uint32_t x = ch << S.nb;                      // tile-local index of the chunk's first amplitude
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                             tile_u32 + (x + (x >> 9)) * 16u),
                         "l"(psi + base + off), "r"(chunk_bytes), "r"(mbar)
                         : "memory");
        }
    };
    uint32_t parity = 0;
    if (blockIdx.x < ntiles) issue(blockIdx.x);
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(mbar, parity);
        parity ^= 1u;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            if (r < S.nrounds) {
                const uint32_t s0 = arrival_stride_bytes(S.rounds[r].regpos[0]), s1 = arrival_stride_bytes(S.rounds[r].regpos[1]);
                const uint32_t s2 = arrival_stride_bytes(S.rounds[r].regpos[2]), s3 = arrival_stride_bytes(S.rounds[r].regpos[3]);
                const uint32_t sb = arrival_stride_bytes(S.rounds[r].thrpos[7]);
#pragma unroll
                for (int batch = 0; batch < 2; ++batch) {
                    double ar[NREG], ai[NREG];
                    const uint32_t b0 = tile_u32 + tbase[r] + (batch ? sb : 0u);
                    const uint32_t bh[4] = {b0, b0 + s2, b0 + s3, b0 + s2 + s3};
                    const uint32_t lo[4] = {0u, s0, s1, s0 + s1};
#pragma unroll
                    for (int j = 0; j < NREG; ++j) {
                        const double2 v = lds128(bh[j >> 2] + lo[j & 3]);
                        ar[j] = v.x;
                        ai[j] = v.y;
                    }
                    if (r == S.nrounds - 1 && batch == 1) {  // the buffer is free: the next tile's copies overlap the arithmetic
                        __syncthreads();
                        if (t + gridDim.x < ntiles) issue(t + gridDim.x);
                    }
#pragma unroll
                    for (int k = 0; k < RB; ++k) {
#pragma unroll
                        for (int j = 0; j < NREG; ++j) {
                            if (!((j >> k) & 1)) {
                                const int f = j | (1 << k);
                                acc[r][2 * k] += ar[j] * ar[f] + ai[j] * ai[f];
                                acc[r][2 * k + 1] += ar[j] * ai[f] - ai[j] * ar[f];
                            }
                        }
                    }
                }
            }
        }
    }
    // block reduction, one row per measured qubit
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < RB; ++k) {
            double v[2] = {acc[r][2 * k], acc[r][2 * k + 1]};
            block_sum<2>(v, red);
            const int row = r < S.nrounds ? S.rounds[r].row[k] : -1;
            if (tid == 0 && row >= 0) partial[(size_t)row * S.stride + blockIdx.x] = make_double2(v[0], v[1]);
        }
}

// NG = 1: one tile per CTA (MINB CTAs per SM run free).  NG = 3: one CTA per SM holds three tile groups of 2^(TB-RB) threads,
// each with its own tile buffer; the shared-memory phases of the groups (round-0 load | store, barrier, load | global store) take
// turns through a token handed round the ring 0 -> 1 -> 2 (named barriers): while one group moves data the other two do
// arithmetic, which free-running CTAs do not keep up (measured: their FP64 and shared-memory phases drift into phase).
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// C64: the state in global memory is complex64 (8 B per amplitude; `in` / `out` point to float2 data), everything on chip is FP64.
template <int TB, int RB, bool LIFT, int MINB, bool PEER, int UNB, int NG, bool C64 = false>
__global__ void __launch_bounds__(NG << (TB - RB), MINB)
tile_pass_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                 const __grid_constant__ PassDesc<TB, RB> P, uint64_t ntiles, const __grid_constant__ PeerPtrs peers) {
    static_assert(!(C64 && PEER), "complex64 states are not sharded");
    constexpr int kAmpShift = C64 ? 3 : 4;  // log2 bytes per amplitude in global memory
    const char* in_b = reinterpret_cast<const char*>(in);
    constexpr int NT = TB - RB, NTHR = 1 << NT, NREG = 1 << RB, NSLOT = RB * (RB - 1) / 2;
    static_assert(NG == 1 || NG == 3, "one free-running tile per CTA, or three tile groups in a token ring");
    constexpr int RS = RB + 1, NV = 1 << RS, HS = RS / 2;  // split rounds: register qubits, reals per thread, group size
    static_assert(RB == 5, "round header packs six register-bit offsets; split rounds need an even RB + 1");
    static_assert(HS * HS <= NSLOT, "split-round slots share the coefficient table");
    using Round = typename PassDesc<TB, RB>::Round;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // per round: {rs0|rs1<<16, rs2|rs3<<16, rs4|rs5<<16, mask|sync<<16|nleft<<17 (nleft = 0: complex round)}, {ts0|ts1<<16, ts2|ts3<<16, ts4|ts5<<16, ts6}
    // rs = byte stride of a register bit, ts = byte stride of a thread bit
    __shared__ __align__(16) uint4 hdr_tab[2 * (kMaxRounds + 1)];
    __shared__ __align__(8) unsigned long long tile_mbar[NG];
    const uint32_t tid_cta = tid_once();
    // tile group of this thread; the lane-0 broadcast tells the compiler that it is warp-uniform (uniform registers for the
    // tile index and everything derived from it)
    const uint32_t grp = NG == 1 ? 0u : __shfl_sync(0xffffffffu, tid_cta >> NT, 0);
    const uint32_t mbar_u32 = smem_addr_once(&tile_mbar[grp]);
    const uint32_t tile_u32 = smem_addr_once(smem_raw + (size_t)grp * tile_bytes<TB>());
    const uint32_t hdr_u32 = smem_addr_once(hdr_tab);
    double2* tile = reinterpret_cast<double2*>(smem_raw + (size_t)grp * tile_bytes<TB>());
    const uint32_t tid = NG == 1 ? tid_cta : (tid_cta & (uint32_t)(NTHR - 1));  // thread index inside the group
    // barrier over the threads of one tile group; token ring: barrier 4 + g is "the token arrives at group g"
    auto group_sync = [&]() {
        if constexpr (NG == 1) __syncthreads();
        else named_bar_sync(1u + grp, NTHR);
    };
    // two tokens: kind 0 = the store phase of a round (shared-memory stores or the global store), kind 1 = the load phase
    auto token_acquire = [&](uint32_t kind) {
        if constexpr (NG > 1) named_bar_sync(4u + 3u * kind + grp, 2 * NTHR);
    };
    auto token_release = [&](uint32_t kind) {
        if constexpr (NG > 1) named_bar_arrive(4u + 3u * kind + (grp + 1u == (uint32_t)NG ? 0u : grp + 1u), 2 * NTHR);
    };
    const uint32_t flags = P.flags;
    const int nr = P.nrounds;
    const uint32_t half = tid & 1u;   // split rounds: which real half-state this thread holds
    const uint32_t tq = tid >> 1;     // split rounds: thread-qubit bits
    const bool first_split = P.rounds[0].split != 0, last_split = P.rounds[nr - 1].split != 0;

    for (int r = tid; r < nr; r += NTHR) {
        const Round& R = P.rounds[r];
        uint4 h, g;
        h.x = (uint32_t)R.regsw[0] | ((uint32_t)R.regsw[1] << 16);
        h.y = (uint32_t)R.regsw[2] | ((uint32_t)R.regsw[3] << 16);
        h.z = (uint32_t)R.regsw[4] | ((uint32_t)R.regsw[5] << 16);
        h.w = (R.mask & 0xffffu) | ((uint32_t)(R.cta_sync ? 1u : 0u) << 16) | ((uint32_t)(R.split & 3u) << 17);
        uint32_t ts[8];
        for (int b = 0; b < 8; ++b) ts[b] = b < (R.split ? NT - 1 : NT) ? 16u * pad_slot(1u << R.thrpos[b]) : 0u;
        g.x = ts[0] | (ts[1] << 16);
        g.y = ts[2] | (ts[3] << 16);
        g.z = ts[4] | (ts[5] << 16);
        g.w = ts[6];
        hdr_tab[2 * r] = h;
        hdr_tab[2 * r + 1] = g;
    }
    // byte offset of this thread's base slot in a round: sum of the strides of its set thread bits
    auto thread_base = [&](const uint4& g, bool split) -> uint32_t {
        const uint32_t bits = split ? tq : tid;
        uint32_t b = split ? (half << 3) : 0u;
        if (bits & 1u) b += g.x & 0xffffu;
        if (bits & 2u) b += g.x >> 16;
        if (bits & 4u) b += g.y & 0xffffu;
        if (bits & 8u) b += g.y >> 16;
        if (bits & 16u) b += g.z & 0xffffu;
        if (bits & 32u) b += g.z >> 16;
        if (bits & 64u) b += g.w;
        return b;
    };
    // fused load: byte offset of this thread's round-0 base amplitude in the arrival layout
    uint32_t lin0 = 0;
    if (first_split) {
#pragma unroll
        for (int b = 0; b < NT - 1; ++b)
            if ((tq >> b) & 1u) lin0 += arrival_stride_bytes(P.rounds[0].thrpos[b]);
    } else {
#pragma unroll
        for (int b = 0; b < NT; ++b)
            if ((tid >> b) & 1u) lin0 += arrival_stride_bytes(P.rounds[0].thrpos[b]);
    }
    // per-thread constants of the (unfused) load and store phases
    const uint32_t ld_slot = pad_slot(tid);  // index bits [0, NT) <- thread id
    uint32_t st_l = 0;
    uint64_t st_d = 0;
    if (flags & kPassFuseStore) {
        if (last_split) {
#pragma unroll
            for (int b = 0; b < NT - 1; ++b)
                if ((tq >> b) & 1u) st_d |= 1ull << P.last_thr_out[b];
        } else {
#pragma unroll
            for (int b = 0; b < NT; ++b)
                if ((tid >> b) & 1u) st_d |= 1ull << P.last_thr_out[b];
        }
    } else {
#pragma unroll
        for (int b = 0; b < NT; ++b)
            if ((tid >> b) & 1u) {
                st_l |= 1u << P.st_lsrc[b];
                st_d |= 1ull << P.st_odst[b];
            }
        st_l = pad_slot(st_l);
    }
    // split rounds: the odd half sees every rotation with the opposite angle; equivalently it keeps the
    // Im-role elements (f = 1) negated and uses the same coefficients (which must stay warp-uniform)
    const uint32_t smask = half << 31;
#ifdef XYK_TRACE
    __shared__ uint32_t tr_scratch_mem[4];
    const uint32_t tr_scratch = smem_addr_once(&tr_scratch_mem[tid >> 5]);
    unsigned long long* tr_p = xyk_trace_buf ? xyk_trace_buf + (((size_t)blockIdx.x * NG + grp) * (NTHR / 32) + (tid >> 5)) * kTraceMaxE : nullptr;
    int tr_i = 2, tr_k = 0;
    bool tr_on = false;
    const unsigned long long tr_c0 = trace_clock();
    if (tr_p && (tid & 31u) == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        tr_p[0] = smid;
    }
#endif
    if (tid == 0) mbar_init(mbar_u32, 1);
    __syncthreads();  // tables and mbarriers visible
    if (NG > 1 && grp + 1u == (uint32_t)NG) {  // both tokens start at group 0
        token_release(0);
        token_release(1);
    }

    // Fused load = the whole 64 KiB tile arrives by bulk asynchronous copies (eight 8 KiB segments) into the front of the
    // tile buffer; it is issued as soon as the previous tile's last round has read its data out of the
    // buffer, so the global load overlaps that round's arithmetic and the global store.
    uint32_t tile_parity = 0;
    // tile indices fit 32 bits (at most 2^31 local amplitudes)
    const uint32_t tstride = gridDim.x * NG;   // tiles between two consecutive tiles of a group
    const uint32_t tfirst = grp * gridDim.x + blockIdx.x;   // group 0 of every CTA first: few tiles spread over all SMs
    const uint32_t ntile32 = (uint32_t)ntiles;
    auto issue_tile_load = [&](uint32_t tt) {
        if constexpr (C64) bulk_load_tile_c64<TB>(tile_u32, in_b + ((uint64_t)tt << (TB + kAmpShift)), mbar_u32);
        else bulk_load_tile<TB>(tile_u32, in + ((uint64_t)tt << TB), mbar_u32);
    };
    if ((flags & kPassFuseLoad) && tid == 0 && tfirst < ntile32) issue_tile_load(tfirst);

    // every group of a CTA runs the same number of iterations: the token ring needs all of them (a group without a tile left
    // only passes the token on)
    for (uint32_t t = tfirst; t - grp * gridDim.x < ntile32; t += tstride) {  // as long as group 0 of this CTA has a tile
        if (NG > 1 && t >= ntile32) {
            for (int k = 0; k < nr; ++k) {
                token_acquire(1);
                token_release(1);
                token_acquire(0);
                token_release(0);
            }
            continue;
        }
        const double2* src = in + ((uint64_t)t << TB);
#ifdef XYK_TRACE
        // traced window: the first nk tiles this CTA starts after k0 * 1024 cycles of kernel time (co-resident CTAs overlap in time)
        if (tr_p != nullptr && (tid & 31u) == 0) {
            if (tr_k == 0 && (long long)(trace_clock() - tr_c0) > (long long)xyk_trace_k0 * 1024) tr_k = 1;
            tr_on = tr_k >= 1 && tr_k <= xyk_trace_nk;
            if (tr_k >= 1) ++tr_k;
        }
#endif
        auto prefetch_next = [&]() {
            if (tid == 0 && !(flags & kPassNoPrefetch) && t + tstride < ntile32) {
                if (flags & kPassPfEvictLast) prefetch_l2_bulk_evict_last(in_b + ((uint64_t)(t + tstride) << (TB + kAmpShift)), 1u << (TB + kAmpShift));
                else prefetch_l2_bulk(in_b + ((uint64_t)(t + tstride) << (TB + kAmpShift)), 1u << (TB + kAmpShift));
            }
        };
        if (!(flags & (kPassPfLate | kPassFuseLoad))) prefetch_next();
        uint64_t obase = P.out_base;
        for (int b = 0; b < P.nq - TB; ++b)
            if ((t >> b) & 1ull) obase |= 1ull << P.hi_out[b];
        uint4 hdr = lds_u4(hdr_u32);
        uint32_t boff = tile_u32 + thread_base(lds_u4(hdr_u32 + 16u), first_split);

        for (int r = 0; r < nr; ++r) {
            uint32_t rb[RS];
            rb[0] = hdr.x & 0xffffu;
            rb[1] = hdr.x >> 16;
            rb[2] = hdr.y & 0xffffu;
            rb[3] = hdr.y >> 16;
            rb[4] = hdr.z & 0xffffu;
            rb[5] = hdr.z >> 16;
            const uint32_t nl = (hdr.w >> 17) & 3u;  // 0: complex round; else the left-group size of a split round
            const bool split = nl != 0u;
            // unbalanced split round (UNB = 2: 2 | 4, UNB = 1: 1 | 5): separate code, present only in the kernel instantiation for it
            const bool unb = UNB != 0 && nl == (uint32_t)UNB;
            const Round& R = P.rounds[r];  // gate coefficients stay in the constant bank: a uniform operand of DFMA (measured: a
                                           // shared-memory coefficient stream into vector registers costs 4-5 % of the layer)
            const uint32_t boff1 = boff + 8u - (half << 4);  // split rounds: base of the elements with f(j) = 1
            if ((flags & kPassPfLate) && !(flags & kPassFuseLoad) && r == nr - 1) prefetch_next();
            // complex round: v[2j], v[2j+1] = Re, Im of register index j (RB bits);
            // split round:   v[j] = one real component of register index j (RB + 1 bits)
            double v[NV];

            // ---------------- load ----------------
#ifdef XYK_TRACE
            if (r > 0 && tr_on) {  // the barrier before this round has released once a shared-memory load returns
                trace_consume(tr_scratch, lds64(hdr_u32));
                XYK_TR(3);
            }
#endif
            if (r == 0 && (flags & kPassFuseLoad)) {
                XYK_TR(1);
                mbar_wait(mbar_u32, tile_parity);
                tile_parity ^= 1u;
                token_acquire(1);
                XYK_TR(2);
                if constexpr (C64) {
                    // widen in place: every thread pulls its 2^RB float2 amplitudes out of the staging area (front of the buffer,
                    // consecutive lanes on consecutive amplitudes), the group synchronises, then the FP64 arrival layout is written
                    float2 f[NREG];
#pragma unroll
                    for (int k = 0; k < NREG; ++k) f[k] = lds_f2(tile_u32 + ((((uint32_t)k << NT) + tid) << 3));
                    group_sync();
#pragma unroll
                    for (int k = 0; k < NREG; ++k) {
                        const uint32_t x = ((uint32_t)k << NT) + tid;
                        sts128(tile_u32 + (x + (x >> 9)) * 16u, make_double2((double)f[k].x, (double)f[k].y));
                    }
                    group_sync();
                }
                if (unb) {
                    if constexpr (UNB) {
                        uint32_t st[RS];
                        for (int k = 0; k < RS; ++k) st[k] = P.arr_reg[k];
                        split_exchange<(UNB ? UNB : 2), RS, NV, false>(v, tile_u32 + lin0 + (half << 3), half, st, smask);
                    }
                } else if (split) {
                    const uint32_t p0 = tile_u32 + lin0 + (half << 3);
                    const uint32_t p1 = p0 + 8u - (half << 4);  // elements with f = 1: the other component
                    uint32_t bh[1 << HS], lo[1 << HS];
#pragma unroll
                    for (int c = 0; c < (1 << HS); ++c) {
                        uint32_t hi = 0, l = 0;
#pragma unroll
                        for (int k = 0; k < HS; ++k) {
                            if ((c >> k) & 1) hi += P.arr_reg[HS + k];
                            if ((c >> k) & 1) l += P.arr_reg[k];
                        }
                        bh[c] = ((__builtin_popcount(c) & 1) ? p1 : p0) + hi;
                        lo[c] = l;
                    }
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        v[j] = lds64(bh[j >> HS] + lo[j & ((1 << HS) - 1)]);
                        if (__builtin_popcount(j >> HS) & 1) v[j] = flip_sign(v[j], smask);
                    }
                } else {
                    const uint32_t p0 = tile_u32 + lin0;
                    uint32_t bh[8], lo[4];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint32_t hi = 0;
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if ((c >> k) & 1) hi += P.arr_reg[2 + k];
                        bh[c] = p0 + hi;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) lo[c] = ((c & 1) ? P.arr_reg[0] : 0u) + ((c & 2) ? P.arr_reg[1] : 0u);
#pragma unroll
                    for (int j = 0; j < NREG; ++j) {
                        const double2 a = lds128(bh[j >> 2] + lo[j & 3]);
                        v[2 * j] = a.x;
                        v[2 * j + 1] = a.y;
                    }
                }
            } else {
                if (r == 0) {
                    token_acquire(1);
                    group_sync();  // every warp is done reading the previous tile out of shared memory
#pragma unroll
                    for (int k = 0; k < NREG; ++k) {
                        const uint32_t hi = pad_slot((uint32_t)k << NT);  // constant after unrolling
                        if constexpr (C64) {
                            const float2 f = __ldg(reinterpret_cast<const float2*>(in_b) + ((uint64_t)t << TB) + ((uint32_t)k << NT) + tid);
                            tile[ld_slot + hi] = make_double2((double)f.x, (double)f.y);
                        } else {
                            cp_async16(&tile[ld_slot + hi], src + ((uint32_t)k << NT) + tid);
                        }
                    }
                    cp_async_wait_all();
                    group_sync();
                }
                // address = per-thread base of the high register bits (8 vector adds per burst) + warp-uniform
                // offset of the low register bits: LDS [R + UR], no per-access vector arithmetic
                if (unb) {
                    if constexpr (UNB) {
                        split_exchange<(UNB ? UNB : 2), RS, NV, false>(v, boff, half, rb, smask);
                    }
                } else if (split) {
                    uint32_t bh[1 << HS], lo[1 << HS];
#pragma unroll
                    for (int c = 0; c < (1 << HS); ++c) {
                        uint32_t hi = 0, l = 0;
#pragma unroll
                        for (int k = 0; k < HS; ++k) {
                            if ((c >> k) & 1) hi += rb[HS + k];
                            if ((c >> k) & 1) l += rb[k];
                        }
                        // component f ^ half, f = parity of the right-group (high) bits: two bases 8 bytes apart
                        bh[c] = ((__builtin_popcount(c) & 1) ? boff1 : boff) + hi;
                        lo[c] = l;
                    }
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        v[j] = lds64(bh[j >> HS] + lo[j & ((1 << HS) - 1)]);
                        if (__builtin_popcount(j >> HS) & 1) v[j] = flip_sign(v[j], smask);
                    }
                } else {
                    uint32_t bh[8], lo[4];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint32_t hi = 0;
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if ((c >> k) & 1) hi += rb[2 + k];
                        bh[c] = boff + hi;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) lo[c] = ((c & 1) ? rb[0] : 0u) + ((c & 2) ? rb[1] : 0u);
#pragma unroll
                    for (int j = 0; j < NREG; ++j) {
                        const double2 a = lds128(bh[j >> 2] + lo[j & 3]);
                        v[2 * j] = a.x;
                        v[2 * j + 1] = a.y;
                    }
                }
            }

            token_release(1);  // loads issued: the next group may start its load phase
            XYK_TR_DEP(4, v[NV - 1]);
            // the last round holds the tile in registers: the buffer is free for the next tile
            if ((flags & kPassFuseLoad) && (flags & kPassFuseStore) && r == nr - 1) {
                group_sync();
                if (tid == 0 && t + tstride < ntile32) issue_tile_load(t + tstride);
            }

            // ---------------- diagonal phase (complex round 0 only) ----------------
            if (r == 0 && (flags & kPassHasPhase)) {
                // tile / thread part of the phase: three table factors (no sincos, no loop over the tile-id bits)
                const double2 f0 = __ldg(P.ph_tab + tid);
                const double2 f1 = __ldg(P.ph_tab + kPhTabLo + (t & ((1u << P.ph_b1) - 1u)));
                const double2 f2 = __ldg(P.ph_tab + kPhTabHi + (t >> P.ph_b1));
                const double gr = f1.x * f2.x - f1.y * f2.y, gi = f1.x * f2.y + f1.y * f2.x;
                const double cs = f0.x * gr - f0.y * gi, sn = f0.x * gi + f0.y * gr;
#pragma unroll
                for (int j = 0; j < NREG; ++j) {
                    const double2 f = P.ph_reg[j];
                    const double fr = cs * f.x - sn * f.y, fi = cs * f.y + sn * f.x;
                    const double xr = v[2 * j], xi = v[2 * j + 1];
                    v[2 * j] = xr * fr - xi * fi;
                    v[2 * j + 1] = xr * fi + xi * fr;
                }
            }

            // ---------------- gates ----------------
            const uint32_t mask = hdr.w & 0xffffu;
            if (unb) {
                if constexpr (UNB) {
                    split_gates<(UNB ? UNB : 2), RS, NV, LIFT>(v, mask, R);
                }
            } else if (split) {
                // a full round (all nine gates: the common case) runs as straight-line code: no mask test between the gates, so
                // the next gate's first rotations issue while the previous gate's last ones are still in the FP64 pipe
                auto split_round = [&](auto all) {
#pragma unroll
                    for (int x = 0; x < HS; ++x) {
#pragma unroll
                        for (int y = HS; y < RS; ++y) {
                            const int slot = HS * x + (y - HS);
                            if (decltype(all)::value || (mask & (1u << slot))) {
                                const double ca = R.ca[slot], cb = R.cb[slot];
#pragma unroll
                                for (int m = 0; m < NV / 4; ++m) {
                                    const int j0 = insert_zero(insert_zero(m, x), y);
                                    const int jA = j0 | (1 << x), jB = j0 | (1 << y);
                                    // half 0 holds Re(A), Im(B) when f(jA) = 0 and Im(A), Re(B) otherwise (roles swap)
                                    if (__builtin_popcount(jA >> HS) & 1) rotate2<LIFT>(v[jB], v[jA], ca, cb);
                                    else rotate2<LIFT>(v[jA], v[jB], ca, cb);
                                }
                            }
                        }
                    }
                };
                if (mask == (1u << (HS * HS)) - 1u) split_round(std::true_type{});
                else split_round(std::false_type{});
            } else {
#pragma unroll
                for (int x = 0; x < RB; ++x) {
#pragma unroll
                    for (int y = x + 1; y < RB; ++y) {
                        const int slot = x * RB - x * (x + 1) / 2 + (y - x - 1);
                        if (mask & (1u << slot)) {
                            const double ca = R.ca[slot], cb = R.cb[slot];
#pragma unroll
                            for (int m = 0; m < NREG / 4; ++m) {
                                const int j0 = insert_zero(insert_zero(m, x), y);
                                const int jA = j0 | (1 << x), jB = j0 | (1 << y);
                                rotate2<LIFT>(v[2 * jA], v[2 * jB + 1], ca, cb);
                                rotate2<LIFT>(v[2 * jB], v[2 * jA + 1], ca, cb);
                            }
                        }
                    }
                }
            }

            // ---------------- store / exchange ----------------
#ifdef XYK_TRACE
            if (tr_on) {  // every value the gates may have touched last is available
                trace_consume(tr_scratch, v[NV - 1]);
                trace_consume(tr_scratch, v[NV - 2]);
                trace_consume(tr_scratch, v[NV / 2 - 1]);
                trace_consume(tr_scratch, v[NV / 2 + 1]);
                trace_consume(tr_scratch, v[NV / 4]);
                trace_consume(tr_scratch, v[NV / 4 + 3]);
                trace_consume(tr_scratch, v[5]);
                trace_consume(tr_scratch, v[2]);
                XYK_TR(5);
            }
#endif
            token_acquire(0);
            if (r == nr - 1 && (flags & kPassFuseStore)) {
                const uint64_t dbase = obase + st_d;
                if constexpr (PEER) {
                    if (unb) {
                        if constexpr (UNB) {
                            split_global_store<(UNB ? UNB : 2), RS, NV, PEER>(v, out, peers, P.nq, dbase, P.last_reg_out, half, smask);
                        }
                    } else if (split) {
#pragma unroll
                        for (int j = 0; j < NV; ++j) {
                            uint64_t off = 0;
#pragma unroll
                            for (int k = 0; k < RS; ++k)
                                if ((j >> k) & 1) off += 1ull << P.last_reg_out[k];
                            const uint32_t fj = __builtin_popcount(j >> HS) & 1;
                            store_out_comp<PEER>(out, peers, P.nq, dbase + off, fj ^ half, fj ? flip_sign(v[j], smask) : v[j]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < NREG; ++j) {
                            uint64_t off = 0;
#pragma unroll
                            for (int k = 0; k < RB; ++k)
                                if ((j >> k) & 1) off += 1ull << P.last_reg_out[k];
                            store_out<PEER>(out, peers, P.nq, dbase + off, make_double2(v[2 * j], v[2 * j + 1]));
                        }
                    }
                } else {
                    // per-thread byte pointer once per tile + warp-uniform byte offsets of the register bits
                    char* pb = reinterpret_cast<char*>(out) + (dbase << kAmpShift);
                    using sreal = typename std::conditional<C64, float, double>::type;    // storage scalar
                    using scplx = typename std::conditional<C64, float2, double2>::type;  // storage amplitude
                    if (unb) {
                        if constexpr (UNB) {
                            split_global_store<(UNB ? UNB : 2), RS, NV, PEER, C64>(v, out, peers, P.nq, dbase, P.last_reg_out, half, smask);
                        }
                    } else if (split) {
                        // eight per-thread pointers over the three high register bits (component by their parity) and
                        // eight warp-uniform 32-bit index offsets over the three low bits: one IMAD.WIDE per store
                        sreal* q0 = reinterpret_cast<sreal*>(pb) + half;        // elements with f = 0: component `half`
                        sreal* q1 = reinterpret_cast<sreal*>(pb) + (1u - half); // elements with f = 1: the other component
                        sreal* bq[8];
                        uint32_t lo2[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint32_t hi = 0, l = 0;
#pragma unroll
                            for (int k = 0; k < HS; ++k) {
                                if ((c >> k) & 1) hi += P.st_reg_idx[HS + k];
                                if ((c >> k) & 1) l += P.st_reg_idx[k];
                            }
                            bq[c] = ((__builtin_popcount(c) & 1) ? q1 : q0) + 2ull * hi;
                            lo2[c] = 2u * l;
                        }
#pragma unroll
                        for (int j = 0; j < NV; ++j) {
                            const bool fj = __builtin_popcount(j >> HS) & 1;
                            __stcs(bq[j >> HS] + lo2[j & ((1 << HS) - 1)], (sreal)(fj ? flip_sign(v[j], smask) : v[j]));
                        }
                    } else {
                        scplx* bq[8];
                        uint32_t lo[4];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint32_t hi = 0;
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                if ((c >> k) & 1) hi += P.st_reg_idx[2 + k];
                            bq[c] = reinterpret_cast<scplx*>(pb) + hi;
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) lo[c] = ((c & 1) ? P.st_reg_idx[0] : 0u) + ((c & 2) ? P.st_reg_idx[1] : 0u);
#pragma unroll
                        for (int j = 0; j < NREG; ++j) {
                            scplx a;
                            a.x = (sreal)v[2 * j];
                            a.y = (sreal)v[2 * j + 1];
                            __stcs(bq[j >> 2] + lo[j & 3], a);
                        }
                    }
                }
                token_release(0);
                XYK_TR(6);
            } else {
                // next round's header and base slot: issued ahead of the stores, back before the barrier
                const uint4 hdr_next = lds_u4(hdr_u32 + (uint32_t)(r + 1) * 32u);
                const uint4 thr_next = lds_u4(hdr_u32 + (uint32_t)(r + 1) * 32u + 16u);
                // tile boundary: the previous tile's last round (other warps) must be done reading
                if (r == 0 && (flags & kPassFuseLoad)) group_sync();
                if (unb) {
                    if constexpr (UNB) {
                        split_exchange<(UNB ? UNB : 2), RS, NV, true>(v, boff, half, rb, smask);
                    }
                } else if (split) {
                    uint32_t bh[1 << HS], lo[1 << HS];
#pragma unroll
                    for (int c = 0; c < (1 << HS); ++c) {
                        uint32_t hi = 0, l = 0;
#pragma unroll
                        for (int k = 0; k < HS; ++k) {
                            if ((c >> k) & 1) hi += rb[HS + k];
                            if ((c >> k) & 1) l += rb[k];
                        }
                        bh[c] = ((__builtin_popcount(c) & 1) ? boff1 : boff) + hi;
                        lo[c] = l;
                    }
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const bool fj = __builtin_popcount(j >> HS) & 1;
                        sts64(bh[j >> HS] + lo[j & ((1 << HS) - 1)], fj ? flip_sign(v[j], smask) : v[j]);
                    }
                } else {
                    uint32_t bh[8], lo[4];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint32_t hi = 0;
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if ((c >> k) & 1) hi += rb[2 + k];
                        bh[c] = boff + hi;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) lo[c] = ((c & 1) ? rb[0] : 0u) + ((c & 2) ? rb[1] : 0u);
#pragma unroll
                    for (int j = 0; j < NREG; ++j) sts128(bh[j >> 2] + lo[j & 3], make_double2(v[2 * j], v[2 * j + 1]));
                }
                XYK_TR(6);
                if (r + 1 < nr) token_release(0);
                if ((hdr.w >> 16) & 1u) group_sync();
                else __syncwarp();
                if (r + 1 < nr) token_acquire(1);
                hdr = hdr_next;
                boff = tile_u32 + thread_base(thr_next, ((hdr_next.w >> 17) & 3u) != 0u);
            }
        }

        if (!(flags & kPassFuseStore)) {
            // ---- store through the bit permutation, staged in shared memory ----
            const uint64_t dbase = obase + st_d;
#pragma unroll
            for (int k = 0; k < NREG; ++k) {
                uint32_t lk = 0;
                uint64_t dk = 0;
#pragma unroll
                for (int b = 0; b < RB; ++b)
                    if ((k >> b) & 1) {
                        lk |= 1u << P.st_lsrc[NT + b];
                        dk |= 1ull << P.st_odst[NT + b];
                    }
                store_out<PEER, C64>(out, peers, P.nq, dbase + dk, tile[st_l + pad_slot(lk)]);
            }
            token_release(0);
            if (flags & kPassFuseLoad) {
                group_sync();
                if (tid == 0 && t + tstride < ntile32) issue_tile_load(t + tstride);
            }
        }
    }
    if (NG > 1 && grp == 0) {  // absorb the last hand-overs (barrier state balanced at exit)
        token_acquire(0);
        token_acquire(1);
    }
}

}  // namespace xyk
