// Micro-benchmark (B200): bandwidth of a bit-permuting copy as a function of the contiguous
// chunk size on the WRITE side (reads are fully contiguous) and on the READ side.
// Decides how many tile qubits consecutive passes must share (SchedConfig::need_shared).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/chunk_bw.cu -o gpurun_out/chunk_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

// dst index: low c bits kept, then the remaining (nb - c) bits rotated by r (a transpose of the high part)
__global__ void permcopy(const double2* __restrict__ src, double2* __restrict__ dst, int nb, int c, int r, int scatter_write) {
    const uint64_t dim = 1ull << nb;
    const int hb = nb - c;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t lo = i & ((1ull << c) - 1), hi = i >> c;
        const uint64_t rot = ((hi << r) | (hi >> (hb - r))) & ((1ull << hb) - 1);
        const uint64_t j = (rot << c) | lo;
        if (scatter_write) dst[j] = src[i];
        else dst[i] = src[j];
    }
}

int main() {
    const int nb = 28;  // 4 GiB per buffer
    double2 *a, *b;
    cudaMalloc(&a, sizeof(double2) << nb);
    cudaMalloc(&b, sizeof(double2) << nb);
    cudaMemset(a, 0, sizeof(double2) << nb);
    cudaMemset(b, 0, sizeof(double2) << nb);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int sw = 1; sw >= 0; --sw)
        for (int c = 0; c <= 8; ++c) {
            float best = 1e30f;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                permcopy<<<148 * 16, 256>>>(a, b, nb, c, 12, sw);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            const double gb = 2.0 * (double)(sizeof(double2) << nb) / 1e9;
            printf("%s chunk=%5d B  %.3f ms  %.1f GB/s\n", sw ? "scatter-write" : "gather-read  ", 16 << c, best, gb / (best * 1e-3));
        }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
