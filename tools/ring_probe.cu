// Does an explicit token ring over the shared-memory phases of three co-resident tile groups recover the FP64 / shared-memory
// overlap that free-running groups lose once every phase carries dependent latency (address set-up, barrier drain, per-gate
// branches)?  One CTA of 384 threads = three groups of four warps; per round and group:
//   [G gates x 48 DFMA, a dependent integer chain of LATG ops between gates] [chain LATS] [64 STS.64] group barrier
//   [chain LATL] [64 LDS.64]
// MODE 0: free running; 1: one token around store+barrier+load; 2: two tokens (stores / loads), ring order 0 -> 1 -> 2.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ring_probe.cu -o tools/_build/ring_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double lds64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;\n" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// dependent chain of n multiply-adds (about 4-5 cycles each, one issue slot each)
__device__ __forceinline__ uint32_t chain(uint32_t x, int n) {
    for (int i = 0; i < n; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(3u), "r"(1u));
    return x;
}

template <int G, int MODE>
__global__ void __launch_bounds__(384, 1) ring(int iters, int latg, int lats, int latl, long long* cycles, double* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    const int tid = threadIdx.x, g = tid >> 7, gt = tid & 127, gnext = g == 2 ? 0 : g + 1;
    double v[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = 1.0 + 0.001 * j + 1e-6 * tid;
    const double ca = 1e-9, cb = -1e-9;
    uint32_t w64 = base + g * 65536u + (gt >> 1) * 16u + (gt & 1u) * 8u;
    __syncthreads();
    // barrier ids: 1..3 group barriers; 4..6 token S (or the single token) handed TO group g; 7..9 token L
    if (MODE >= 1 && g == 2) bar_arrive(4, 256);
    if (MODE == 2 && g == 2) bar_arrive(7, 256);
    const long long t0 = clock64();
    uint32_t acc = tid;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int gte = 0; gte < G; ++gte) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                double &p = v[(4 * m + gte) & 63], &q = v[(4 * m + gte + 2) & 63];
                p = fma(ca, q, p); q = fma(cb, p, q); p = fma(ca, q, p);
            }
            acc = chain(acc, latg);
        }
        acc = chain(acc, lats);
        w64 += acc & 0u;
        if (MODE >= 1) bar_sync(4 + g, 256);
#pragma unroll
        for (int j = 0; j < 64; ++j) sts64(w64 + j * 1024, v[j]);
        if (MODE == 2) bar_arrive(4 + gnext, 256);
        bar_sync(1 + g, 128);
        acc = chain(acc, latl);
        w64 += acc & 0u;
        if (MODE == 2) bar_sync(7 + g, 256);
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = lds64(w64 + (j ^ (it & 1)) * 1024);
        if (MODE == 2) bar_arrive(7 + gnext, 256);
        if (MODE == 1) bar_arrive(4 + gnext, 256);
    }
    if (MODE >= 1 && g == 0) bar_sync(4, 256);
    if (MODE == 2 && g == 0) bar_sync(7, 256);
    const long long t1 = clock64();
    double s = acc;
#pragma unroll
    for (int j = 0; j < 64; ++j) s += v[j];
    sink[blockIdx.x * 384 + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int G, int MODE>
void run(int latg, int lats, int latl, long long* cyc, double* sink) {
    const int smem = 3 * 65536, iters = 300, grid = 148;
    cudaFuncSetAttribute(ring<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ring<G, MODE><<<grid, 384, smem>>>(10, latg, lats, latl, cyc, sink);
    ring<G, MODE><<<grid, 384, smem>>>(iters, latg, lats, latl, cyc, sink);
    cudaDeviceSynchronize();
    static long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0;
    for (int b = 0; b < grid; ++b) sum += h[b];
    printf("G=%d mode=%d lat(g,s,l)=(%2d,%3d,%3d): %7.1f cycles per round of 3 tiles   %s\n", G, MODE, latg, lats, latl, sum / grid / iters,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* cyc; double* sink;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&sink, 148 * 384 * sizeof(double));
    const int lat[4][3] = {{0, 0, 0}, {8, 40, 60}, {10, 60, 100}, {12, 100, 150}};
    for (auto& l : lat) {
        run<9, 0>(l[0], l[1], l[2], cyc, sink);
        run<9, 1>(l[0], l[1], l[2], cyc, sink);
        run<9, 2>(l[0], l[1], l[2], cyc, sink);
    }
    return 0;
}
