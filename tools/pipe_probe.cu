// Micro-benchmark of the two on-chip resources the fused tile pass contends for (one SM):
//   * the shared-memory data pipe driven by LDS.128 / STS.128 from W warps,
//   * the FP64 pipe driven by dependent-chain DFMAs with 16 independent chains per thread,
//   * both at once (some warps move data, the others do arithmetic).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pipe_probe.cu -o pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 lds128(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const double2& v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};\n" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// mode bit0: warps < nmem do smem traffic; bit1: warps >= nmem do DFMA.  Each "unit" of memory work
// = 32 STS.128 + barrier-free + 32 LDS.128 per thread (= one exchange of a 32-amplitude block);
// each unit of math = G gates x 48 DFMA per thread.
template <int G>
__global__ void probe(int nmem, int iters, long long* cycles, double* sink, int do_mem, int do_math) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2 a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = make_double2(1.0 + j + lane, 0.5 * j);
    const double ca = 1e-9 * (lane + 1), cb = -1e-9;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nmem) {
        if (do_mem) {
            // each warp owns a 16 KiB window: 32 rows of 512 B, lane-linear (conflict free)
            const uint32_t w0 = base + (uint32_t)warp * 16384u + (uint32_t)lane * 16u;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int j = 0; j < 32; ++j) sts128(w0 + j * 512, a[j]);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = lds128(w0 + ((j * 512) ^ ((it & 1) << 9)));
            }
        }
    } else if (do_math) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int gte = 0; gte < G; ++gte) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    double& x = a[2 * m].x;
                    double& y = a[2 * m + 1].y;
                    x = fma(ca, y, x);
                    y = fma(cb, x, y);
                    x = fma(ca, y, x);
                }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += a[j].x + a[j].y;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
}

// three groups of four warps; token rotation as in tile_pass_grouped_kernel: slot = acquire, 32 STS,
// group barrier, 32 LDS, release; then G gates of arithmetic.  mode 0: free running (group barrier only).
template <int G>
__global__ void rot(int iters, long long* cycles, double* sink, int use_token) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp >> 2, gnext = g == 2 ? 0 : g + 1;
    double2 a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = make_double2(1.0 + j + lane, 0.5 * j);
    const double ca = 1e-9 * (lane + 1), cb = -1e-9;
    const uint32_t w0 = base + (uint32_t)warp * 16384u + (uint32_t)lane * 16u;
    __syncthreads();
    if (use_token && g == 2) asm volatile("bar.arrive %0, 256;" ::"r"(8) : "memory");
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int gte = 0; gte < G; ++gte) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                double& x = a[2 * m].x;
                double& y = a[2 * m + 1].y;
                x = fma(ca, y, x);
                y = fma(cb, x, y);
                x = fma(ca, y, x);
            }
        }
        if (use_token) asm volatile("bar.sync %0, 256;" ::"r"(8 + g) : "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) sts128(w0 + j * 512, a[j]);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = lds128(w0 + ((j * 512) ^ ((it & 1) << 9)));
        if (use_token) asm volatile("bar.arrive %0, 256;" ::"r"(8 + gnext) : "memory");
    }
    if (use_token && g == 0) asm volatile("bar.sync %0, 256;" ::"r"(8) : "memory");
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += a[j].x + a[j].y;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (lane == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
}

template <int G>
void run_rot(long long* cyc, double* sink, int smem) {
    cudaFuncSetAttribute(rot<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[32];
    for (int tok = 0; tok < 2; ++tok) {
        rot<G><<<1, 384, smem>>>(200, cyc, sink, tok);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < 12; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("rotation G=%2d token=%d : %7.1f cycles per round of 3 tiles (pipe floor ~3072, math floor %d)\n", G, tok, mx / 200.0,
               (int)(3 * G * 48 * 2.13));
    }
}

int main() {
    long long* cyc;
    double* sink;
    cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
    cudaMalloc(&sink, 148 * 1024 * sizeof(double));
    const int smem = 12 * 16384;
    cudaFuncSetAttribute(probe<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 200;
    long long h[32];
    auto run = [&](int warps, int nmem, int do_mem, int do_math, const char* label) {
        probe<5><<<1, warps * 32, smem>>>(nmem, iters, cyc, sink, do_mem, do_math);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mmax = 0, fmax = 0;
        for (int w = 0; w < warps; ++w) {
            if (w < nmem) mmax = h[w] > mmax ? h[w] : mmax;
            else fmax = h[w] > fmax ? h[w] : fmax;
        }
        const double mem_bytes = (double)nmem * iters * 32 * 2 * 512;
        const double dfma_warp = (double)(warps - nmem) * iters * 5 * 48;
        printf("%-28s warps=%2d mem_warps=%2d : mem %8lld cyc (%6.1f B/clk)   math %8lld cyc (%5.2f warp-DFMA/clk/SM)\n", label, warps, nmem,
               mmax, do_mem && mmax ? mem_bytes / mmax : 0.0, fmax, do_math && fmax ? dfma_warp / fmax : 0.0);
    };
    run(1, 1, 1, 0, "mem only");  // warm-up
    for (int w : {1, 2, 4, 8, 12}) run(w, w, 1, 0, "mem only");
    for (int w : {1, 2, 4, 8, 12}) run(w, 0, 0, 1, "math only");
    run(12, 4, 1, 1, "4 mem + 8 math");
    run(12, 4, 1, 0, "4 mem (+8 idle)");
    run(12, 4, 0, 1, "8 math (+4 idle)");
    run(12, 8, 1, 1, "8 mem + 4 math");
    run(12, 6, 1, 1, "6 mem + 6 math");
    run(8, 4, 1, 1, "4 mem + 4 math");
    run_rot<2>(cyc, sink, smem);
    run_rot<5>(cyc, sink, smem);
    run_rot<8>(cyc, sink, smem);
    run_rot<10>(cyc, sink, smem);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
