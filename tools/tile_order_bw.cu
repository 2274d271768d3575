// Micro-benchmark (B200): the fused tile pass reduced to its data movement.  Every CTA reads
// contiguous 64 KiB tiles (32 x LDG.128 per thread) and writes them as scattered chunks of
// 2^c amplitudes (32 x STG.128 per thread), exactly like tile_pass_kernel with one round.
// Question: does the ORDER in which tiles are traversed matter?  "near": the tiles in flight at the
// same time write ADJACENT chunks (together they fill whole 64 KiB output regions); "far": they
// write chunks that are 2^(24-c) amplitudes apart.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/tile_order_bw.cu -o tile_order_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int MINB>
__global__ void __launch_bounds__(128, MINB) tilecopy(const double2* __restrict__ in, double2* __restrict__ out, int nb, int c, int near_mode,
                                                      int streaming, int spread = 0) {
    const uint64_t ntiles = 1ull << (nb - 12);
    const int abits = 12 - c;
    const uint32_t tid = threadIdx.x;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        uint64_t a, rest;
        if (near_mode) {
            a = t & ((1ull << abits) - 1);
            rest = t >> abits;
        } else {
            a = t >> (nb - 12 - abits);
            rest = t & ((1ull << (nb - 12 - abits)) - 1);
        }
        // spread: the tile's own upper bits go to the TOP output bits (every chunk of a tile in a different
        // 2^(nb-12+c)-amplitude region, i.e. a different DRAM/TLB page), the rest right above bit 12
        const uint64_t obase = spread ? ((a << c) | (rest << 12)) : ((a << c) | (rest << (24 - c)));
        const int hshift = spread ? nb - (12 - c) : 12;
        const double2* src = in + (t << 12);
        double2 v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = streaming ? __ldcs(src + k * 128 + tid) : src[k * 128 + tid];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const uint32_t o = k * 128 + tid;
            const uint64_t dst = obase | (o & ((1u << c) - 1)) | ((uint64_t)(o >> c) << hshift);
            if (streaming) __stcs(out + dst, v[k]);
            else out[dst] = v[k];
        }
    }
}

int main() {
    const int nb = 30;  // 16 GiB per buffer
    double2 *a, *b;
    cudaMalloc(&a, sizeof(double2) << nb);
    cudaMalloc(&b, sizeof(double2) << nb);
    cudaMemset(a, 0, sizeof(double2) << nb);
    cudaMemset(b, 0, sizeof(double2) << nb);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int streaming = 1; streaming >= 1; --streaming)
        for (int near_mode = 1; near_mode >= 0; --near_mode)
            for (int c = 2; c <= 5; ++c) {
                float best = 1e30f;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(e0);
                    tilecopy<3><<<148 * 3, 128>>>(a, b, nb, c, near_mode, streaming);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (it > 0 && ms < best) best = ms;
                }
                const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
                printf("%s %s chunk=%5d B  %.3f ms  %.1f GB/s\n", streaming ? "ldcs/stcs" : "plain    ", near_mode ? "near" : "far ", 16 << c, best,
                       gb / (best * 1e-3));
            }
    for (int spread = 0; spread < 2; ++spread)
        for (int near_mode = 1; near_mode >= 0; --near_mode)
            for (int c = 3; c <= 6; ++c) {
                float best = 1e30f;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(e0);
                    tilecopy<3><<<148 * 3, 128>>>(a, b, nb, c, near_mode, 1, spread);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (it > 0 && ms < best) best = ms;
                }
                const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
                printf("spread=%d %s chunk=%5d B  %.3f ms  %.1f GB/s\n", spread, near_mode ? "near" : "far ", 16 << c, best, gb / (best * 1e-3));
            }
    // more CTAs per SM (fewer registers needed by a plain copy): is 3 x 128 threads enough to saturate HBM?
    for (int c = 3; c <= 4; ++c) {
        float best = 1e30f;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            tilecopy<6><<<148 * 6, 128>>>(a, b, nb, c, 1, 1);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
        printf("6 CTAs/SM near chunk=%5d B  %.3f ms  %.1f GB/s\n", 16 << c, best, gb / (best * 1e-3));
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
