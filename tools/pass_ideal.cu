// Upper bound for the fused tile pass design on a B200: the same resource usage as
// tile_pass_kernel (3 CTAs x 128 threads per SM, 64 KiB tile per CTA, 32 amplitudes per thread,
// R rounds of G pair rotations with a shared-memory exchange between rounds, contiguous tile
// reads, scattered 2^c-amplitude chunk writes) but with fixed, immediate addressing and no
// descriptors.  What this kernel cannot reach, the real one cannot either.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pass_ideal.cu -o pass_ideal
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const double2& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gptr), "r"(bytes) : "memory");
}

template <int G>
__global__ void __launch_bounds__(128, 3) pass_ideal(const double2* __restrict__ in, double2* __restrict__ out, int nb, int c, int R, int prefetch,
                                                     double ca, double cb) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = base + warp * 16384u + lane * 16u;
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t a_ = t & ((1ull << abits) - 1), rest = t >> abits;
        const uint64_t obase = (a_ << c) | (rest << (24 - c));
        const double2* src = in + (t << 12);
        if (prefetch && tid == 0 && t + gridDim.x < ntiles) prefetch_l2_bulk(in + ((t + gridDim.x) << 12), 65536u);
        double2 a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = __ldcs(src + k * 128 + tid);
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int gte = 0; gte < G; ++gte) {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    // two real 2x2 rotations per amplitude pair, three shears each (as rotate2<LIFT>)
                    double& x0 = a[4 * m + 1].x;
                    double& y0 = a[4 * m + 2].y;
                    double& x1 = a[4 * m + 2].x;
                    double& y1 = a[4 * m + 1].y;
                    x0 = fma(ca, y0, x0); y0 = fma(cb, x0, y0); x0 = fma(ca, y0, x0);
                    x1 = fma(ca, y1, x1); y1 = fma(cb, x1, y1); x1 = fma(ca, y1, x1);
                }
            }
            if (r + 1 < R) {
                if (r == 0) __syncthreads();
#pragma unroll
                for (int j = 0; j < 32; ++j) sts128(w0 + j * 512, a[j]);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = lds128(w0 + ((j * 512) ^ ((r & 1) << 9)));
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const uint32_t o = k * 128 + tid;
            const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << 12);
            __stcs(out + dst, a[k]);
        }
    }
}

// Variant: between full shared-memory exchanges, S cheap "shuffle swaps" (one register qubit <-> one lane
// qubit: half of the thread's amplitudes trade places with the partner lane, 64 SHFL.32 per thread)
// each followed by 4 gates (the incoming qubit against the four resident ones).
template <int G>
__global__ void __launch_bounds__(128, 3) pass_shfl(const double2* __restrict__ in, double2* __restrict__ out, int nb, int c, int R, int S, int prefetch,
                                                    double ca, double cb) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = base + warp * 16384u + lane * 16u;
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t a_ = t & ((1ull << abits) - 1), rest = t >> abits;
        const uint64_t obase = (a_ << c) | (rest << (24 - c));
        const double2* src = in + (t << 12);
        if (prefetch && tid == 0 && t + gridDim.x < ntiles) prefetch_l2_bulk(in + ((t + gridDim.x) << 12), 65536u);
        double2 a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = __ldcs(src + k * 128 + tid);
        for (int r = 0; r < R; ++r) {
            for (int sidx = 0; sidx <= S; ++sidx) {
                const int ng = sidx == 0 ? G : 4;
                for (int gte = 0; gte < ng; ++gte) {
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        double& x0 = a[4 * m + 1].x;
                        double& y0 = a[4 * m + 2].y;
                        double& x1 = a[4 * m + 2].x;
                        double& y1 = a[4 * m + 1].y;
                        x0 = fma(ca, y0, x0); y0 = fma(cb, x0, y0); x0 = fma(ca, y0, x0);
                        x1 = fma(ca, y1, x1); y1 = fma(cb, x1, y1); x1 = fma(ca, y1, x1);
                    }
                }
                if (sidx < S) {
                    const int l = sidx % 5;
                    const bool beta = (lane >> l) & 1u;
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const double sx = beta ? a[m].x : a[m + 16].x;
                        const double sy = beta ? a[m].y : a[m + 16].y;
                        const double rx = __shfl_xor_sync(0xffffffffu, sx, 1 << l);
                        const double ry = __shfl_xor_sync(0xffffffffu, sy, 1 << l);
                        if (beta) { a[m].x = rx; a[m].y = ry; }
                        else { a[m + 16].x = rx; a[m + 16].y = ry; }
                    }
                }
            }
            if (r + 1 < R) {
                if (r == 0) __syncthreads();
#pragma unroll
                for (int j = 0; j < 32; ++j) sts128(w0 + j * 512, a[j]);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = lds128(w0 + ((j * 512) ^ ((r & 1) << 9)));
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const uint32_t o = k * 128 + tid;
            const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << 12);
            __stcs(out + dst, a[k]);
        }
    }
}

// Variant: "real-split" rounds.  A pair rotation only mixes Re(a) with Im(b) and Im(a) with Re(b); when all
// gates of a round go between two colour classes of register qubits, the state splits into two
// independent real halves, so a thread can hold ONE real component of 64 amplitudes (6 register qubits,
// K_{3,3} = 9 gates per round) in the same 128 registers.  Exchange = 64 LDS.64 + 64 STS.64 per thread.
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(addr), "d"(v) : "memory");
}
template <int G>
__global__ void __launch_bounds__(128, 3) pass_split(const double* __restrict__ in, double* __restrict__ out, int nb, int c, int R, int prefetch,
                                                     double ca, double cb) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = base + warp * 16384u + lane * 8u;
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t a_ = t & ((1ull << abits) - 1), rest = t >> abits;
        const uint64_t obase = (a_ << c) | (rest << (24 - c));
        const double* src = in + (t << 13);
        if (prefetch && tid == 0 && t + gridDim.x < ntiles) prefetch_l2_bulk(in + ((t + gridDim.x) << 13), 65536u);
        double a[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) a[k] = __ldcs(src + k * 128 + tid);
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int gte = 0; gte < G; ++gte) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    double& x0 = a[4 * m + 1];
                    double& y0 = a[4 * m + 2];
                    x0 = fma(ca, y0, x0); y0 = fma(cb, x0, y0); x0 = fma(ca, y0, x0);
                }
            }
            if (r + 1 < R) {
                if (r == 0) __syncthreads();
#pragma unroll
                for (int j = 0; j < 64; ++j) sts64(w0 + j * 256, a[j]);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 64; ++j) a[j] = lds64(w0 + ((j * 256) ^ ((r & 1) << 9)));
            }
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const uint32_t o = (k * 128 + tid) >> 1;  // complex slot
            const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << 12);
            __stcs(out + 2 * dst + (tid & 1), a[k]);
        }
    }
}

// Variant of pass_split: the tile is loaded by ONE bulk asynchronous copy (TMA, cp.async.bulk + mbarrier) into
// the CTA's own tile buffer as soon as the last round has read its data out of it, i.e. the next
// tile's global load overlaps the last round's arithmetic and the global store.
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t phase) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(mbar),
        "r"(phase)
        : "memory");
}
template <int G, int ROWS = 1>
__global__ void __launch_bounds__(128, 3) pass_split_tma(const double* __restrict__ in, double* __restrict__ out, int nb, int c, int R, double ca,
                                                         double cb) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar_storage;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&mbar_storage);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = base + warp * 16384u + lane * 8u;
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint64_t t) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(65536u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(base),
                     "l"(in + (t << 13)), "r"(65536u), "r"(mbar)
                     : "memory");
    };
    auto pad = [](uint32_t x) { return x + (x >> 3) + (x >> 6) + (x >> 9); };
    // ROWS > 1: the tile arrives as ROWS bulk copies of 65536 / ROWS bytes, each placed at its padded slot
    auto issue_rows = [&](uint64_t t) {
        constexpr uint32_t RB_ = 65536u / ROWS;      // bytes per copy
        constexpr int PER = ROWS / 128 > 0 ? ROWS / 128 : 1;
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(65536u) : "memory");
        }
        __syncthreads();
        if (tid * PER < ROWS) {
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const uint32_t row = tid * PER + i;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                                 base + pad(row * (RB_ / 16u)) * 16u),
                             "l"(reinterpret_cast<const char*>(in + (t << 13)) + (size_t)row * RB_), "r"(RB_), "r"(mbar)
                             : "memory");
            }
        }
    };
    if (ROWS == 1) {
        if (tid == 0 && blockIdx.x < ntiles) issue(blockIdx.x);
    } else if (blockIdx.x < ntiles) issue_rows(blockIdx.x);
    uint32_t phase = 0;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t a_ = t & ((1ull << abits) - 1), rest = t >> abits;
        const uint64_t obase = (a_ << c) | (rest << (24 - c));
        double a[64];
        mbar_wait(mbar, phase);
        phase ^= 1;
        // round 0 reads the linear tile: thread tid, register k -> double index k * 128 + tid (conflict free)
        if (ROWS == 1) {
#pragma unroll
            for (int k = 0; k < 64; ++k) a[k] = lds64(base + (k * 128 + tid) * 8u);
        } else {
#pragma unroll
            for (int k = 0; k < 64; ++k) a[k] = lds64(base + pad((uint32_t)(k * 64)) * 16u + pad(tid >> 1) * 16u + (tid & 1) * 8u);
        }
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int gte = 0; gte < G; ++gte) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    double& x0 = a[4 * m + 1];
                    double& y0 = a[4 * m + 2];
                    x0 = fma(ca, y0, x0); y0 = fma(cb, x0, y0); x0 = fma(ca, y0, x0);
                }
            }
            if (r + 1 < R) {
                if (r == 0) __syncthreads();
#pragma unroll
                for (int j = 0; j < 64; ++j) sts64(w0 + j * 256, a[j]);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 64; ++j) a[j] = lds64(w0 + ((j * 256) ^ ((r & 1) << 9)));
                if (r + 2 == R) {  // the last round holds its data: the buffer is free for the next tile
                    __syncthreads();
                    if (ROWS == 1) {
                        if (tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);
                    } else if (t + gridDim.x < ntiles) issue_rows(t + gridDim.x);
                }
            }
        }
        if (R == 1) {
            __syncthreads();
            if (ROWS == 1) {
                if (tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);
            } else if (t + gridDim.x < ntiles) issue_rows(t + gridDim.x);
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const uint32_t o = (k * 128 + tid) >> 1;  // complex slot
            const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << 12);
            __stcs(out + 2 * dst + (tid & 1), a[k]);
        }
    }
}

// Variant: the next tile is staged by per-thread cp.async (LDGSTS, 16 B each, 32 per thread) into a PADDED
// layout (arbitrary placement, no lane constraint on the first round), issued after the last round's loads.
__global__ void __launch_bounds__(128, 3) pass_split_cpasync(const double* __restrict__ in, double* __restrict__ out, int nb, int c, int R, double ca,
                                                             double cb) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = base + warp * 16384u + lane * 8u;
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    auto pad = [](uint32_t x) { return x + (x >> 3) + (x >> 6) + (x >> 9); };
    const uint32_t st0 = base + pad(tid) * 16u;
    auto issue = [&](uint64_t t) {
        const double2* src = reinterpret_cast<const double2*>(in) + (t << 12) + tid;
#pragma unroll
        for (int k = 0; k < 32; ++k)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(st0 + pad((uint32_t)k << 7) * 16u), "l"(src + k * 128) : "memory");
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if (blockIdx.x < ntiles) issue(blockIdx.x);
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t a_ = t & ((1ull << abits) - 1), rest = t >> abits;
        const uint64_t obase = (a_ << c) | (rest << (24 - c));
        double a[64];
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        // round 0 reads the padded tile (here: 8-byte units k * 128 + tid of the dense index, through pad())
#pragma unroll
        for (int k = 0; k < 64; ++k) a[k] = lds64(base + pad((uint32_t)(k * 64)) * 16u + pad(tid >> 1) * 16u + (tid & 1) * 8u);
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int gte = 0; gte < 9; ++gte) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    double& x0 = a[4 * m + 1];
                    double& y0 = a[4 * m + 2];
                    x0 = fma(ca, y0, x0); y0 = fma(cb, x0, y0); x0 = fma(ca, y0, x0);
                }
            }
            if (r + 1 < R) {
                if (r == 0) __syncthreads();
#pragma unroll
                for (int j = 0; j < 64; ++j) sts64(w0 + j * 256, a[j]);
                __syncthreads();
#pragma unroll
                for (int j = 0; j < 64; ++j) a[j] = lds64(w0 + ((j * 256) ^ ((r & 1) << 9)));
                if (r + 2 == R) {
                    __syncthreads();
                    if (t + gridDim.x < ntiles) issue(t + gridDim.x);
                }
            }
        }
        if (R == 1) {
            __syncthreads();
            if (t + gridDim.x < ntiles) issue(t + gridDim.x);
        }
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            const uint32_t o = (k * 128 + tid) >> 1;
            const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << 12);
            __stcs(out + 2 * dst + (tid & 1), a[k]);
        }
    }
}

// Realism ladder for split rounds with the TMA tile load: start from pass_split_tma and add the real kernel's
// features one by one.
//  F & 1 : sign fix-ups (odd half keeps the f = 1 elements negated: LOP3 on 32 of 64 values at load and store)
//  F & 2 : runtime gate mask (nine slot branches), coefficients from the constant bank
//  F & 4 : runtime strides: per-round header from shared memory, 8 high bases + 8 uniform low offsets
//  F & 8 : per-thread base from seven thread-bit strides (second header word)
//  F & 16: global store with runtime 64-bit offsets (output bit positions from the descriptor)
//  F & 32: per-tile output base from a loop over the tile-id bits
struct SplitDesc {
    uint16_t rs[16][8];      // register-bit byte strides per round
    uint16_t ts[16][8];      // thread-bit byte strides per round
    uint32_t mask[16];
    double ca[16][10], cb[16][10];
    uint8_t last_reg_out[8], last_thr_out[8], hi_out[40];
    int nq;
};
__device__ __forceinline__ double flip_sign(double v, uint32_t signmask) {
    return __hiloint2double(__double2hiint(v) ^ (int)signmask, __double2loint(v));
}
template <int F>
__global__ void __launch_bounds__(128, 3) pass_split_real(const double* __restrict__ in, double* __restrict__ out, int nb, int c, int R,
                                                          const __grid_constant__ SplitDesc D, double ca0, double cb0) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar_storage;
    __shared__ __align__(16) uint4 hdr_tab[32];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&mbar_storage);
    const uint32_t hdr_u32 = (uint32_t)__cvta_generic_to_shared(hdr_tab);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t half = tid & 1u, tq = tid >> 1, smask = half << 31;
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    if (tid < 16) {
        hdr_tab[2 * tid] = make_uint4(D.rs[tid][0] | D.rs[tid][1] << 16, D.rs[tid][2] | D.rs[tid][3] << 16, D.rs[tid][4] | D.rs[tid][5] << 16,
                                      D.mask[tid] | 1u << 16 | 1u << 17);
        hdr_tab[2 * tid + 1] = make_uint4(D.ts[tid][0] | D.ts[tid][1] << 16, D.ts[tid][2] | D.ts[tid][3] << 16, D.ts[tid][4] | D.ts[tid][5] << 16,
                                          D.ts[tid][6]);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](uint64_t t) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(65536u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(base),
                     "l"(in + (t << 13)), "r"(65536u), "r"(mbar)
                     : "memory");
    };
    auto thread_base = [&](const uint4& g) -> uint32_t {
        uint32_t b = half << 3;
        if (tq & 1u) b += g.x & 0xffffu;
        if (tq & 2u) b += g.x >> 16;
        if (tq & 4u) b += g.y & 0xffffu;
        if (tq & 8u) b += g.y >> 16;
        if (tq & 16u) b += g.z & 0xffffu;
        if (tq & 32u) b += g.z >> 16;
        return b;
    };
    // fixed-pattern base: warp-private 16 KiB window, 8-byte units (as pass_split)
    const uint32_t w0 = base + warp * 16384u + lane * 8u;
    if (tid == 0 && blockIdx.x < ntiles) issue(blockIdx.x);
    uint32_t phase = 0;
    uint64_t st_d = 0;
    if (F & 16) {
#pragma unroll
        for (int b = 0; b < 6; ++b)
            if ((tq >> b) & 1u) st_d |= 1ull << D.last_thr_out[b];
    }
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        uint64_t obase;
        if (F & 32) {
            obase = 0;
            for (int b = 0; b < D.nq - 12; ++b)
                if ((t >> b) & 1ull) obase |= 1ull << D.hi_out[b];
        } else {
            const uint64_t a_ = t & ((1ull << abits) - 1), rest = t >> abits;
            obase = (a_ << c) | (rest << (24 - c));
        }
        double a[64];
        uint4 hdr = make_uint4(0, 0, 0, 0), thr = hdr;
        if (F & 4) {
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hdr.x), "=r"(hdr.y), "=r"(hdr.z), "=r"(hdr.w) : "r"(hdr_u32));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(thr.x), "=r"(thr.y), "=r"(thr.z), "=r"(thr.w) : "r"(hdr_u32 + 16));
        }
        uint32_t boff = (F & 8) ? base + thread_base(thr) : w0;
        mbar_wait(mbar, phase);
        phase ^= 1;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
            a[k] = lds64(base + (k * 128 + tid) * 8u);
            if ((F & 1) && (__builtin_popcount(k >> 3) & 1)) a[k] = flip_sign(a[k], smask);
        }
        for (int r = 0; r < R; ++r) {
            uint32_t rb[6];
            uint32_t mask = 0x1ff;
            if (F & 4) {
                rb[0] = hdr.x & 0xffffu; rb[1] = hdr.x >> 16; rb[2] = hdr.y & 0xffffu; rb[3] = hdr.y >> 16; rb[4] = hdr.z & 0xffffu; rb[5] = hdr.z >> 16;
                mask = hdr.w & 0xffffu;
            }
            const uint32_t boff1 = boff + 8u - (half << 4);
            if (r > 0) {
                if (F & 4) {
                    uint32_t bh[8], lo[8];
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        uint32_t hi = 0, l = 0;
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if ((cc >> k) & 1) hi += rb[3 + k];
                            if ((cc >> k) & 1) l += rb[k];
                        }
                        bh[cc] = ((__builtin_popcount(cc) & 1) ? boff1 : boff) + hi;
                        lo[cc] = l;
                    }
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        a[j] = lds64(bh[j >> 3] + lo[j & 7]);
                        if ((F & 1) && (__builtin_popcount(j >> 3) & 1)) a[j] = flip_sign(a[j], smask);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        a[j] = lds64(w0 + ((j * 256) ^ ((r & 1) << 9)));
                        if ((F & 1) && (__builtin_popcount(j >> 3) & 1)) a[j] = flip_sign(a[j], smask);
                    }
                }
            }
            if (r + 1 == R && R > 1) {
                __syncthreads();
                if (tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);
            }
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int y = 3; y < 6; ++y) {
                    const int slot = 3 * x + (y - 3);
                    if (!(F & 2) || (mask & (1u << slot))) {
                        const double ca = (F & 2) ? D.ca[r][slot] : ca0, cb = (F & 2) ? D.cb[r][slot] : cb0;
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            const int lowm = m & ((1 << x) - 1);
                            const int midm = (m >> x) & ((1 << (y - x - 1)) - 1);
                            const int him = m >> (y - 1);
                            const int j0 = lowm | (midm << (x + 1)) | (him << (y + 1));
                            const int jA = j0 | (1 << x), jB = j0 | (1 << y);
                            if (__builtin_popcount(jA >> 3) & 1) {
                                double& X = a[jB]; double& Y = a[jA];
                                X = fma(ca, Y, X); Y = fma(cb, X, Y); X = fma(ca, Y, X);
                            } else {
                                double& X = a[jA]; double& Y = a[jB];
                                X = fma(ca, Y, X); Y = fma(cb, X, Y); X = fma(ca, Y, X);
                            }
                        }
                    }
                }
            if (r + 1 < R) {
                uint4 hn = hdr, tn = thr;
                if (F & 4) {
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hn.x), "=r"(hn.y), "=r"(hn.z), "=r"(hn.w) : "r"(hdr_u32 + (r + 1) * 32));
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(tn.x), "=r"(tn.y), "=r"(tn.z), "=r"(tn.w) : "r"(hdr_u32 + (r + 1) * 32 + 16));
                }
                if (r == 0) __syncthreads();
                if (F & 4) {
                    uint32_t bh[8], lo[8];
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc) {
                        uint32_t hi = 0, l = 0;
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            if ((cc >> k) & 1) hi += rb[3 + k];
                            if ((cc >> k) & 1) l += rb[k];
                        }
                        bh[cc] = ((__builtin_popcount(cc) & 1) ? boff1 : boff) + hi;
                        lo[cc] = l;
                    }
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const bool fj = __builtin_popcount(j >> 3) & 1;
                        sts64(bh[j >> 3] + lo[j & 7], ((F & 1) && fj) ? flip_sign(a[j], smask) : a[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const bool fj = __builtin_popcount(j >> 3) & 1;
                        sts64(w0 + j * 256, ((F & 1) && fj) ? flip_sign(a[j], smask) : a[j]);
                    }
                }
                __syncthreads();
                hdr = hn;
                thr = tn;
                if (F & 8) boff = base + thread_base(tn);
            }
        }
        if (R == 1) {
            __syncthreads();
            if (tid == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x);
        }
        if (F & 16) {
            char* pb = reinterpret_cast<char*>(out) + ((obase + st_d) << 4);
            char* pb0 = pb + (half << 3);
            char* pb1 = pb + 8 - (half << 3);
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                uint64_t off = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    if ((j >> k) & 1) off += 16ull << D.last_reg_out[k];
                const bool fj = __builtin_popcount(j >> 3) & 1;
                __stcs(reinterpret_cast<double*>((fj ? pb1 : pb0) + off), ((F & 1) && fj) ? flip_sign(a[j], smask) : a[j]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                const uint32_t o = (k * 128 + tid) >> 1;
                const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << 12);
                const bool fj = __builtin_popcount(k >> 3) & 1;
                __stcs(out + 2 * dst + (tid & 1), ((F & 1) && fj) ? flip_sign(a[k], smask) : a[k]);
            }
        }
    }
}

template <int F>
void run_split_real(const double* a, double* b, int nb, cudaEvent_t e0, cudaEvent_t e1) {
    cudaFuncSetAttribute(pass_split_real<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75008);
    SplitDesc D;
    memset(&D, 0, sizeof(D));
    D.nq = nb;
    auto pad = [](uint32_t x) { return x + (x >> 3) + (x >> 6) + (x >> 9); };
    for (int r = 0; r < 16; ++r) {
        // registers on tile positions 6..11, thread qubits on positions 0..5 (lanes 1..3 on positions 0,1,2: conflict free)
        for (int k = 0; k < 6; ++k) D.rs[r][k] = (uint16_t)(16 * pad(1u << (6 + k)));
        for (int k = 0; k < 6; ++k) D.ts[r][k] = (uint16_t)(16 * pad(1u << k));
        D.mask[r] = 0x1ff;
        for (int k = 0; k < 10; ++k) { D.ca[r][k] = 1e-9; D.cb[r][k] = -1e-9; }
    }
    // output permutation: thread qubits (positions 0..5) -> output bits 0..5; registers (6..11) -> bits 12..17; tile id bits fill the rest
    for (int k = 0; k < 6; ++k) { D.last_thr_out[k] = (uint8_t)k; D.last_reg_out[k] = (uint8_t)(12 + k); }
    {
        int nxt = 6;
        for (int b = 0; b < nb - 12; ++b) {
            while (nxt >= 12 && nxt < 18) ++nxt;
            D.hi_out[b] = (uint8_t)nxt++;
            if (nxt == 12) nxt = 18;
        }
    }
    for (int R : {2, 4, 5}) {
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            pass_split_real<F><<<148 * 3, 128, 75008>>>(a, b, nb, 5, R, D, 1e-9, -1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        printf("split-real F=%2d R=%d : %.3f ms per pass\n", F, R, best);
    }
}

// Realism ladder for the complex round: which feature of the real kernel costs what?
//  F & 1: runtime XOR addressing (per-thread base ^ xor-combination of five runtime register-bit offsets)
//  F & 2: runtime gate mask (ten slot branches) with coefficients fetched from shared memory one gate ahead
//  F & 4: per-round header + per-thread base fetched from shared-memory tables
//  F & 8: global loads in 128-byte segments (lanes 0-2 contiguous, the other lane bits far apart)
struct RealDesc {
    uint32_t rb[16][5];
    uint32_t mask[16];
    double ca[16][10], cb[16][10];
};
template <int F>
__global__ void __launch_bounds__(128, 3) pass_real(const double2* __restrict__ in, double2* __restrict__ out, int nb, int c, int R,
                                                    const __grid_constant__ RealDesc D, double ca0, double cb0) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(16) uint4 hdr_tab[16];
    __shared__ uint16_t base_tab[16][128];
    __shared__ __align__(16) double2 coef_tab[16][11];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = base + warp * 16384u + lane * 16u;
    for (int r = 0; r < 16; ++r) {
        base_tab[r][tid] = (uint16_t)(warp * 16384u + lane * 16u);
        if (tid == 0) hdr_tab[r] = make_uint4(D.rb[r][0] | D.rb[r][1] << 16, D.rb[r][2] | D.rb[r][3] << 16, D.rb[r][4] | D.mask[r] << 16, 1);
        if (tid < 11) coef_tab[r][tid] = make_double2(ca0, cb0);
    }
    __syncthreads();
    const uint32_t hdr_u32 = (uint32_t)__cvta_generic_to_shared(hdr_tab), coef_u32 = (uint32_t)__cvta_generic_to_shared(coef_tab),
                   bt_u32 = (uint32_t)__cvta_generic_to_shared(base_tab) + 2 * tid;
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    // F & 8: thread -> tile index with lanes 0..2 on bits 0..2, lanes 3,4 on bits 10,11, warp on bits 8,9; registers on bits 3..7
    const uint32_t lin8 = (lane & 7u) | ((lane >> 3) << 10) | (warp << 8);
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t a_ = t & ((1ull << abits) - 1), rest = t >> abits;
        const uint64_t obase = (a_ << c) | (rest << (24 - c));
        const double2* src = in + (t << 12);
        if (tid == 0 && t + gridDim.x < ntiles) prefetch_l2_bulk(in + ((t + gridDim.x) << 12), 65536u);
        double2 a[32];
        if (F & 8) {
#pragma unroll
            for (int k = 0; k < 32; ++k) a[k] = __ldcs(src + lin8 + (k << 3));
        } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) a[k] = __ldcs(src + k * 128 + tid);
        }
        uint4 hdr = make_uint4(0, 0, 0, 0);
        uint32_t boff = warp * 16384u + lane * 16u;
        if (F & 4) {
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hdr.x), "=r"(hdr.y), "=r"(hdr.z), "=r"(hdr.w) : "r"(hdr_u32));
            asm volatile("{ .reg .u16 h; ld.shared.u16 h, [%1]; cvt.u32.u16 %0, h; }" : "=r"(boff) : "r"(bt_u32));
        }
        for (int r = 0; r < R; ++r) {
            uint32_t rb[5];
            uint32_t mask;
            if (F & 4) {
                rb[0] = hdr.x & 0xffffu; rb[1] = hdr.x >> 16; rb[2] = hdr.y & 0xffffu; rb[3] = hdr.y >> 16; rb[4] = hdr.z & 0xffffu;
                mask = hdr.z >> 16;
            } else {
#pragma unroll
                for (int k = 0; k < 5; ++k) rb[k] = D.rb[r][k];
                mask = D.mask[r];
            }
            uint32_t cptr = coef_u32 + r * 176;
            double2 cnext = make_double2(ca0, cb0);
            if ((F & 2) && !(F & 16)) cnext = lds128(cptr);
            if (r > 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (F & 1) {
                        uint32_t off = 0;
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if ((j >> k) & 1) off ^= rb[k];
                        a[j] = lds128(base + (boff ^ off));
                    } else {
                        a[j] = lds128(w0 + ((j * 512) ^ ((r & 1) << 9)));
                    }
                }
            }
            if (F & 2) {
#pragma unroll
                for (int x = 0; x < 5; ++x)
#pragma unroll
                    for (int y = x + 1; y < 5; ++y) {
                        const int slot = x * 5 - x * (x + 1) / 2 + (y - x - 1);
                        if (mask & (1u << slot)) {
                            double ca, cb;
                            if (F & 16) {
                                ca = D.ca[r][slot];
                                cb = D.cb[r][slot];
                            } else {
                                ca = cnext.x;
                                cb = cnext.y;
                                cptr += 16;
                                cnext = lds128(cptr);
                            }
#pragma unroll
                            for (int m = 0; m < 8; ++m) {
                                const int j0 = (((m >> (y - 1)) << (y + 1)) | (m & ((1 << (y - 1)) - 1) & ~((1 << x) - 1)) << 1 | (m & ((1 << x) - 1)));
                                const int jA = (j0 | (1 << x)) & 31, jB = (j0 | (1 << y)) & 31;
                                double& x0 = a[jA].x; double& y0 = a[jB].y; double& x1 = a[jB].x; double& y1 = a[jA].y;
                                x0 = fma(ca, y0, x0); y0 = fma(cb, x0, y0); x0 = fma(ca, y0, x0);
                                x1 = fma(ca, y1, x1); y1 = fma(cb, x1, y1); x1 = fma(ca, y1, x1);
                            }
                        }
                    }
            } else if (F & 32) {
                cnext = lds128(cptr);
#pragma unroll
                for (int gte = 0; gte < 5; ++gte) {
                    const double ca = cnext.x, cb = cnext.y;
                    cptr += 16;
                    cnext = lds128(cptr);
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        double& x0 = a[4 * m + 1].x; double& y0 = a[4 * m + 2].y; double& x1 = a[4 * m + 2].x; double& y1 = a[4 * m + 1].y;
                        x0 = fma(ca, y0, x0); y0 = fma(cb, x0, y0); x0 = fma(ca, y0, x0);
                        x1 = fma(ca, y1, x1); y1 = fma(cb, x1, y1); x1 = fma(ca, y1, x1);
                    }
                }
            } else {
#pragma unroll
                for (int gte = 0; gte < 5; ++gte)
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        double& x0 = a[4 * m + 1].x; double& y0 = a[4 * m + 2].y; double& x1 = a[4 * m + 2].x; double& y1 = a[4 * m + 1].y;
                        x0 = fma(ca0, y0, x0); y0 = fma(cb0, x0, y0); x0 = fma(ca0, y0, x0);
                        x1 = fma(ca0, y1, x1); y1 = fma(cb0, x1, y1); x1 = fma(ca0, y1, x1);
                    }
            }
            if (r + 1 < R) {
                uint4 hn = hdr;
                uint32_t bn = boff;
                if (F & 4) {
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hn.x), "=r"(hn.y), "=r"(hn.z), "=r"(hn.w) : "r"(hdr_u32 + (r + 1) * 16));
                    asm volatile("{ .reg .u16 h; ld.shared.u16 h, [%1]; cvt.u32.u16 %0, h; }" : "=r"(bn) : "r"(bt_u32 + (r + 1) * 256));
                }
                if (r == 0) __syncthreads();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (F & 1) {
                        uint32_t off = 0;
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if ((j >> k) & 1) off ^= rb[k];
                        sts128(base + (boff ^ off), a[j]);
                    } else {
                        sts128(w0 + j * 512, a[j]);
                    }
                }
                __syncthreads();
                hdr = hn;
                boff = bn;
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const uint32_t o = k * 128 + tid;
            const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << 12);
            __stcs(out + dst, a[k]);
        }
    }
}

template <int F>
void run_real(const double2* a, double2* b, int nb, cudaEvent_t e0, cudaEvent_t e1) {
    cudaFuncSetAttribute(pass_real<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    RealDesc D;
    for (int r = 0; r < 16; ++r) {
        // register bits on the 512-byte rows of the warp's 16 KiB window (conflict free, warp private)
        for (int k = 0; k < 5; ++k) D.rb[r][k] = 512u << k;
        D.mask[r] = 0x155;  // five of the ten slots
        for (int k = 0; k < 10; ++k) { D.ca[r][k] = 1e-9; D.cb[r][k] = -1e-9; }
    }
    for (int R : {2, 4, 6}) {
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            pass_real<F><<<148 * 3, 128, 65536>>>(a, b, nb, 5, R, D, 1e-9, -1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        printf("real F=%2d R=%d : %.3f ms per pass\n", F, R, best);
    }
}

int main() {
    const int nb = 30;
    double2 *a, *b;
    cudaMalloc(&a, sizeof(double2) << nb);
    cudaMalloc(&b, sizeof(double2) << nb);
    cudaMemset(a, 0, sizeof(double2) << nb);
    cudaMemset(b, 0, sizeof(double2) << nb);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(pass_ideal<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(pass_ideal<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(pass_ideal<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    auto run = [&](int G, int R, int c, int pf) {
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            if (G == 0) pass_ideal<0><<<148 * 3, 128, 65536>>>(a, b, nb, c, R, pf, 1e-9, -1e-9);
            else if (G == 5) pass_ideal<5><<<148 * 3, 128, 65536>>>(a, b, nb, c, R, pf, 1e-9, -1e-9);
            else pass_ideal<8><<<148 * 3, 128, 65536>>>(a, b, nb, c, R, pf, 1e-9, -1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
        printf("G=%d gates/round R=%d rounds chunk=%4d B prefetch=%d : %.3f ms per pass  %.0f GB/s  (%.1f%% of 6541.5)\n", G, R, 16 << c, pf, best,
               gb / (best * 1e-3), 100.0 * gb / (best * 1e-3) / 6541.5);
    };
    for (int pf = 0; pf < 2; ++pf)
        for (int G : {0, 5}) {
            for (int R : {1, 2, 4, 6, 8}) run(G, R, 5, pf);
        }
    run(5, 6, 3, 1);
    run(5, 6, 4, 1);
    run(8, 6, 5, 1);
    run(8, 4, 5, 1);
    run_real<0>(a, b, nb, e0, e1);
    run_real<2>(a, b, nb, e0, e1);
    run_real<18>(a, b, nb, e0, e1);
    run_real<32>(a, b, nb, e0, e1);
    cudaFuncSetAttribute(pass_split<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int R : {1, 2, 3, 4, 5, 6}) {
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            pass_split<9><<<148 * 3, 128, 65536>>>((const double*)a, (double*)b, nb, 5, R, 1, 1e-9, -1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
        printf("split: R=%d rounds x 9 gates = %2d gates : %.3f ms per pass  %.0f GB/s (%.1f%%)\n", R, 9 * R, best, gb / (best * 1e-3),
               100.0 * gb / (best * 1e-3) / 6541.5);
    }
    cudaFuncSetAttribute(pass_split_tma<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int R : {1, 2, 3, 4, 5, 6}) {
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            pass_split_tma<9><<<148 * 3, 128, 65536>>>((const double*)a, (double*)b, nb, 5, R, 1e-9, -1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
        printf("split+TMA load: R=%d rounds x 9 gates = %2d gates : %.3f ms per pass  %.0f GB/s (%.1f%%)\n", R, 9 * R, best, gb / (best * 1e-3),
               100.0 * gb / (best * 1e-3) / 6541.5);
    }
    run_split_real<0>((const double*)a, (double*)b, nb, e0, e1);
    run_split_real<1>((const double*)a, (double*)b, nb, e0, e1);
    run_split_real<2>((const double*)a, (double*)b, nb, e0, e1);
    run_split_real<12>((const double*)a, (double*)b, nb, e0, e1);
    run_split_real<14>((const double*)a, (double*)b, nb, e0, e1);
    run_split_real<16>((const double*)a, (double*)b, nb, e0, e1);
    run_split_real<48>((const double*)a, (double*)b, nb, e0, e1);
    run_split_real<63>((const double*)a, (double*)b, nb, e0, e1);
    cudaFuncSetAttribute(pass_split_tma<9, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75008);
    cudaFuncSetAttribute(pass_split_tma<9, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 75008);
    for (int rows : {512, 128})
        for (int R : {1, 2, 4, 5, 6}) {
            float best = 1e30f;
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                if (rows == 512) pass_split_tma<9, 512><<<148 * 3, 128, 75008>>>((const double*)a, (double*)b, nb, 5, R, 1e-9, -1e-9);
                else pass_split_tma<9, 128><<<148 * 3, 128, 75008>>>((const double*)a, (double*)b, nb, 5, R, 1e-9, -1e-9);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
            printf("split+TMA %3d-row padded load: R=%d rounds x 9 gates : %.3f ms per pass  %.0f GB/s (%.1f%%)\n", rows, R, best, gb / (best * 1e-3),
                   100.0 * gb / (best * 1e-3) / 6541.5);
        }
    cudaFuncSetAttribute(pass_split_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 75008);
    for (int R : {1, 2, 3, 4, 5, 6}) {
        float best = 1e30f;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            pass_split_cpasync<<<148 * 3, 128, 75008>>>((const double*)a, (double*)b, nb, 5, R, 1e-9, -1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
        printf("split+cp.async padded load: R=%d rounds x 9 gates = %2d gates : %.3f ms per pass  %.0f GB/s (%.1f%%)\n", R, 9 * R, best,
               gb / (best * 1e-3), 100.0 * gb / (best * 1e-3) / 6541.5);
    }
    cudaFuncSetAttribute(pass_shfl<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int R : {1})
        for (int S : {0, 1, 2, 3, 4}) {
            float best = 1e30f;
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                pass_shfl<5><<<148 * 3, 128, 65536>>>(a, b, nb, 5, R, S, 1, 1e-9, -1e-9);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
            printf("shfl: R=%d full rounds x (5 gates + %d x (swap + 4 gates)) = %2d gates : %.3f ms per pass  %.0f GB/s (%.1f%%)  %.3f ms/gate\n", R, S,
                   R * (5 + 4 * S), best, gb / (best * 1e-3), 100.0 * gb / (best * 1e-3) / 6541.5, best / (R * (5 + 4 * S)));
        }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
