// Micro-benchmark (2 B200s, one process): bandwidth of kernel STORES into a peer GPU's memory over NVLink
// as a function of the contiguous chunk size and the store flavour.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int MODE>
__global__ void peer_store(const double2* __restrict__ src, double2* __restrict__ dst, int nb, int c, int r) {
    const uint64_t dim = 1ull << nb;
    const int hb = nb - c;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t lo = i & ((1ull << c) - 1), hi = i >> c;
        const uint64_t rot = ((hi << r) | (hi >> (hb - r))) & ((1ull << hb) - 1);
        const uint64_t j = (rot << c) | lo;
        const double2 v = src[i];
        if (MODE == 0) dst[j] = v;
        else if (MODE == 1) __stcs(dst + j, v);
        else __stcg(dst + j, v);
    }
}

int main() {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
    const int nb = 27;  // 2 GiB
    double2 *a, *b;
    cudaSetDevice(1);
    cudaMalloc(&b, sizeof(double2) << nb);
    cudaMemset(b, 0, sizeof(double2) << nb);
    cudaSetDevice(0);
    cudaDeviceEnablePeerAccess(1, 0);
    cudaMalloc(&a, sizeof(double2) << nb);
    cudaMemset(a, 0, sizeof(double2) << nb);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int grid : {148 * 16, 148 * 2, 32}) {
        for (int mode = 0; mode < 3; ++mode)
            for (int c : {0, 3, 5, 8}) {
                float best = 1e30f;
                for (int it = 0; it < 3; ++it) {
                    cudaEventRecord(e0);
                    if (mode == 0) peer_store<0><<<grid, 256>>>(a, b, nb, c, 10);
                    else if (mode == 1) peer_store<1><<<grid, 256>>>(a, b, nb, c, 10);
                    else peer_store<2><<<grid, 256>>>(a, b, nb, c, 10);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (it > 0 && ms < best) best = ms;
                }
                printf("grid=%5d mode=%s chunk=%5d B  %.3f ms  %.1f GB/s to peer\n", grid, mode == 0 ? "st   " : (mode == 1 ? "st.cs" : "st.cg"),
                       16 << c, best, (double)(sizeof(double2) << nb) / 1e9 / (best * 1e-3));
            }
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
