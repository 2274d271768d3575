// Which feature of the real tile-pass round costs the FP64 / shared-memory overlap?  Three CTAs of 128 threads per SM
// (64 KiB of dynamic shared memory each, like the pass kernel); every CTA loops over rounds of
//   [G gates x 48 DFMA per thread] [store 64 reals per thread] barrier [load 64 reals per thread].
// F & 1: 64-bit shared-memory accesses (64 LDS.64 + 64 STS.64) instead of 128-bit (32 + 32)
// F & 2: sign flips (LOP3 on the high word) of 32 of the 64 values after the load and before the store
// F & 4: gates behind a runtime mask test each, coefficients from the kernel parameter (constant bank)
// F & 8: runtime strides: eight per-thread bases + eight warp-uniform offsets recomputed every round
// F & 16 (with F & 1): planar tile (Re plane, Im plane) and a WARP-uniform half: no sign flips, every lane of a warp
//         touches the same component; the half-1 warps run the gates with negated (uniform) coefficients
// F & 32 (with F & 1, without F & 4): program order of a round = loads of the first gate's 32 operands, first gate,
//         the other 32 loads, gates 2..G-1, stores of the 32 values the last gate does not touch, last gate, other stores
// Prints cycles per round of three tiles (the three co-resident CTAs together); floors: shared pipe 3 x 1024, FP64 3 x G x 48 x 2.13.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/overlap_probe.cu -o overlap_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Desc {
    double ca[16], cb[16];
    uint32_t mask;
    uint32_t st[6];
};
__device__ __forceinline__ double2 lds128(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(a), "d"(x), "d"(y) : "memory"); }
__device__ __forceinline__ double lds64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double flip(double v, uint32_t m) { return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v)); }
__host__ __device__ constexpr int insert_zero(int v, int pos) { return ((v >> pos) << (pos + 1)) | (v & ((1 << pos) - 1)); }

template <int G, int F>
__global__ void __launch_bounds__(128, 3) rounds(int iters, const __grid_constant__ Desc D, long long* cycles, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t tid = threadIdx.x;
    double v[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = 1.0 + 0.001 * j + 1e-6 * tid;
    const uint32_t smask = (tid & 1u) << 31;
    // lane-linear windows: thread t owns 16-byte column t of 32 rows of 2 KiB (conflict free for both widths)
    const uint32_t w128 = base + tid * 16u;
    // interleaved: lane pair (2k, 2k+1) shares a 16-byte slot; planar: warps 0,1 work on the Re plane, warps 2,3 on the Im plane
    const uint32_t w64 = (F & 16) ? base + (tid >> 6) * 32768u + (tid & 63u) * 8u : base + (tid >> 1) * 16u + (tid & 1u) * 8u;
    const uint32_t row64 = (F & 16) ? 512u : 1024u;  // byte stride of register index j
    __syncthreads();
    const long long t0 = clock64();
    if constexpr ((F & 32) != 0) {
        // explicit program order: the first gate starts on its 32 operands while the other 32 loads are in flight, the
        // values the last gate does not touch are stored before it runs
        const double ca = 1e-9, cb = -1e-9;
        auto gate = [&](int x, int y) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int j0 = insert_zero(insert_zero(m, x), y);
                const int jA = j0 | (1 << x), jB = j0 | (1 << y);
                double &p = (__builtin_popcount(jA >> 3) & 1) ? v[jB] : v[jA], &q = (__builtin_popcount(jA >> 3) & 1) ? v[jA] : v[jB];
                p = fma(ca, q, p); q = fma(cb, p, q); p = fma(ca, q, p);
            }
        };
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int s = 1; s < G - 1; ++s) gate(s / 3, 3 + s % 3);
#pragma unroll
            for (int j = 0; j < 64; ++j)
                if (((j >> 2) & 1) == ((j >> 5) & 1)) sts64(w64 + j * 1024, v[j]);   // untouched by the last gate (2, 5)
            gate(2, 5);
#pragma unroll
            for (int j = 0; j < 64; ++j)
                if (((j >> 2) & 1) != ((j >> 5) & 1)) sts64(w64 + j * 1024, v[j]);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 64; ++j)
                if ((j & 1) != ((j >> 3) & 1)) v[j] = lds64(w64 + (j ^ (it & 1) ^ ((it & 1) << 3)) * 1024);  // operands of the first gate (0, 3)
            gate(0, 3);
#pragma unroll
            for (int j = 0; j < 64; ++j)
                if ((j & 1) == ((j >> 3) & 1)) v[j] = lds64(w64 + (j ^ (it & 1) ^ ((it & 1) << 3)) * 1024);
            __syncthreads();
        }
    } else
    for (int it = 0; it < iters; ++it) {
        // gates
        if (F & 4) {
#pragma unroll
            for (int x = 0; x < 3; ++x)
#pragma unroll
                for (int y = 3; y < 6; ++y) {
                    const int slot = 3 * x + (y - 3);
                    if (slot < G && (D.mask & (1u << slot))) {
                        const bool neg = (F & 16) && (tid >> 6);  // warp-uniform: the half-1 warps rotate by the opposite angle
                        const double ca = neg ? -D.ca[slot] : D.ca[slot], cb = neg ? -D.cb[slot] : D.cb[slot];
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            const int j0 = insert_zero(insert_zero(m, x), y);
                            const int jA = j0 | (1 << x), jB = j0 | (1 << y);
                            double &p = (__builtin_popcount(jA >> 3) & 1) ? v[jB] : v[jA], &q = (__builtin_popcount(jA >> 3) & 1) ? v[jA] : v[jB];
                            p = fma(ca, q, p); q = fma(cb, p, q); p = fma(ca, q, p);
                        }
                    }
                }
        } else {
            const double ca = 1e-9, cb = -1e-9;
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    double &p = v[(4 * m + g) & 63], &q = v[(4 * m + g + 2) & 63];
                    p = fma(ca, q, p); q = fma(cb, p, q); p = fma(ca, q, p);
                }
        }
        // store / barrier / load
        if (F & 1) {
            uint32_t bh[8], lo[8];
            if (F & 8) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t hi = 0, l = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { if ((c >> k) & 1) { hi += D.st[3 + k]; l += D.st[k]; } }
                    bh[c] = w64 + hi; lo[c] = l;
                }
            }
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                double x = v[j];
                if ((F & 2) && !(F & 16) && (__builtin_popcount(j >> 3) & 1)) x = flip(x, smask);
                sts64((F & 8) ? bh[j >> 3] + lo[j & 7] : w64 + j * row64, x);
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                v[j] = lds64((F & 8) ? bh[(j >> 3) ^ (it & 1)] + lo[j & 7] : w64 + (j ^ (it & 1)) * row64);
                if ((F & 2) && !(F & 16) && (__builtin_popcount(j >> 3) & 1)) v[j] = flip(v[j], smask);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sts128(w128 + j * 2048, v[2 * j], v[2 * j + 1]);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 32; ++j) { const double2 a = lds128(w128 + (j ^ (it & 1)) * 2048); v[2 * j] = a.x; v[2 * j + 1] = a.y; }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) s += v[j];
    sink[blockIdx.x * 128 + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int G, int F>
void run(long long* cyc, double* sink) {
    const int smem = 65536, iters = 400, grid = 3 * 148;
    cudaFuncSetAttribute(rounds<G, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Desc D{};
    for (int s = 0; s < 16; ++s) { D.ca[s] = 1e-9 * (s + 1); D.cb[s] = -1e-9; }
    D.mask = 0xffffu;
    const uint32_t st[6] = {1024, 2048, 4096, 8192, 16384, 32768};
    for (int k = 0; k < 6; ++k) D.st[k] = st[k];
    rounds<G, F><<<grid, 128, smem>>>(10, D, cyc, sink);
    rounds<G, F><<<grid, 128, smem>>>(iters, D, cyc, sink);
    cudaDeviceSynchronize();
    static long long h[3 * 148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0; long long mx = 0;
    for (int b = 0; b < grid; ++b) { sum += h[b]; if (h[b] > mx) mx = h[b]; }
    printf("G=%2d F=%2d : %7.1f cycles per round of 3 tiles (mean), max %7.1f   floors: shared %d, fp64 %d   %s\n", G, F, sum / grid / iters, (double)mx / iters,
           3 * 1024, (int)(3 * G * 48 * 2.13), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* cyc; double* sink;
    cudaMalloc(&cyc, 3 * 148 * sizeof(long long));
    cudaMalloc(&sink, 3 * 148 * 128 * sizeof(double));
    run<9, 0>(cyc, sink); run<9, 1>(cyc, sink); run<9, 3>(cyc, sink); run<9, 4>(cyc, sink); run<9, 5>(cyc, sink); run<9, 7>(cyc, sink); run<9, 15>(cyc, sink);
    run<0, 0>(cyc, sink); run<0, 1>(cyc, sink); run<0, 15>(cyc, sink);
    run<5, 0>(cyc, sink); run<5, 15>(cyc, sink);
    run<9, 7>(cyc, sink); run<9, 23>(cyc, sink);   // interleaved + sign flips vs planar + warp-uniform half (static strides)
    run<9, 1>(cyc, sink); run<9, 33>(cyc, sink);   // default program order vs loads / stores interleaved with the first / last gate
    return 0;
}
