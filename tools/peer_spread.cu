// Micro-benchmark (2 B200s, one process): do kernel stores into LOCAL / PEER memory slow down when the 512 chunks (128 B) a
// CTA writes per 64 KiB tile are spread over many 2 MiB pages (address-translation reach), as the fused exchange passes do?
// Tile t (64 KiB, read contiguously) is written as 512 chunks; chunk c of tile t goes to
//   dst = ((c_hi * ntiles + t) * 2^lo + c_lo) * 128 B,   c_lo = c mod 2^lo, c_hi = c >> lo
// lo = 9: the tile stays contiguous (1 page per tile); lo = 0: 512 chunks with stride ntiles * 128 B (512 pages per tile).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void __launch_bounds__(256) spread_store(const double2* __restrict__ src, double2* __restrict__ dst, int log_tiles, int lo) {
    const uint64_t ntiles = 1ull << log_tiles;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        for (int i = threadIdx.x; i < 4096; i += 256) {
            const uint64_t c = i >> 3, e = i & 7;  // chunk (8 amplitudes = 128 B), element
            const uint64_t c_lo = c & ((1ull << lo) - 1), c_hi = c >> lo;
            const uint64_t j = ((((c_hi << log_tiles) + t) << lo) + c_lo) * 8 + e;
            __stcs(dst + j, src[(t << 12) + i]);
        }
    }
}

int main() {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    for (int nb : {27, 29}) {  // 2 GiB, 8 GiB
        double2 *a, *b, *loc;
        cudaSetDevice(1);
        cudaMalloc(&b, sizeof(double2) << nb);
        cudaMemset(b, 0, sizeof(double2) << nb);
        cudaSetDevice(0);
        cudaDeviceEnablePeerAccess(1, 0);
        cudaMalloc(&a, sizeof(double2) << nb);
        cudaMalloc(&loc, sizeof(double2) << nb);
        cudaMemset(a, 0, sizeof(double2) << nb);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int peer = 0; peer < 2; ++peer)
            for (int lo : {9, 6, 3, 0}) {
                float best = 1e30f;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(e0);
                    spread_store<<<148 * 4, 256>>>(a, peer ? b : loc, nb - 12, lo);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (it > 0 && ms < best) best = ms;
                }
                printf("buffer %4.0f GiB  %s  pages per tile %4d : %.3f ms  %.1f GB/s written\n", (double)(sizeof(double2) << nb) / (1ull << 30),
                       peer ? "PEER " : "local", 1 << (9 - lo), best, (double)(sizeof(double2) << nb) / 1e9 / (best * 1e-3));
            }
        cudaFree(a); cudaFree(loc);
        cudaSetDevice(1); cudaFree(b); cudaSetDevice(0);
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
