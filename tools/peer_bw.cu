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

// MODE 3: the same scattered chunks written by bulk-TMA stores from shared memory (cp.async.bulk.global.shared::cta):
// a CTA stages 4 KiB (256 amplitudes) per step, one thread issues 4096 / chunk_bytes bulk stores, two staging buffers.
__global__ void __launch_bounds__(256) peer_store_bulk(const double2* __restrict__ src, double2* __restrict__ dst, int nb, int c, int r) {
    __shared__ __align__(128) double2 stage[2][256];
    const uint64_t nblk = 1ull << (nb - 8);
    const int hb = nb - c;
    int buf = 0;
    for (uint64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x, buf ^= 1) {
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the buffer we are about to fill is free
        __syncthreads();
        stage[buf][threadIdx.x] = src[(blk << 8) + threadIdx.x];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const int per = 1 << c;  // amplitudes per chunk
            for (int k = 0; k < 256; k += per) {
                const uint64_t i = (blk << 8) + k;
                const uint64_t hi = i >> c;
                const uint64_t rot = ((hi << r) | (hi >> (hb - r))) & ((1ull << hb) - 1);
                const uint64_t j = rot << c;
                const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&stage[buf][k]);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + j), "r"(sa), "r"(per * 16) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    // bulk-TMA stores from shared memory
    for (int grid : {148 * 8, 148 * 2})
        for (int c : {3, 4, 5, 6, 8}) {
            float best = 1e30f;
            for (int it = 0; it < 3; ++it) {
                cudaEventRecord(e0);
                peer_store_bulk<<<grid, 256>>>(a, b, nb, c, 10);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            printf("grid=%5d mode=bulk  chunk=%5d B  %.3f ms  %.1f GB/s to peer\n", grid, 16 << c, best, (double)(sizeof(double2) << nb) / 1e9 / (best * 1e-3));
        }
    // both directions at once (the real exchange): GPU 1 stores into GPU 0 while GPU 0 stores into GPU 1
    {
        double2 *a2, *b2;
        cudaSetDevice(1);
        cudaDeviceEnablePeerAccess(0, 0);
        cudaMalloc(&b2, sizeof(double2) << nb);
        cudaMemset(b2, 0, sizeof(double2) << nb);
        cudaStream_t s1;
        cudaStreamCreate(&s1);
        cudaEvent_t f0, f1;
        cudaEventCreate(&f0);
        cudaEventCreate(&f1);
        cudaSetDevice(0);
        cudaMalloc(&a2, sizeof(double2) << nb);
        cudaMemset(a2, 0, sizeof(double2) << nb);
        cudaDeviceSynchronize();
        for (int c : {3, 5, 8}) {
            float best = 1e30f;
            for (int it = 0; it < 3; ++it) {
                cudaSetDevice(1);
                cudaEventRecord(f0, s1);
                peer_store<1><<<148 * 16, 256, 0, s1>>>(b2, a2, nb, c, 10);
                cudaEventRecord(f1, s1);
                cudaSetDevice(0);
                cudaEventRecord(e0);
                peer_store<1><<<148 * 16, 256>>>(a, b, nb, c, 10);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                cudaEventSynchronize(f1);
                float ms0, ms1;
                cudaEventElapsedTime(&ms0, e0, e1);
                cudaSetDevice(1);
                cudaEventElapsedTime(&ms1, f0, f1);
                cudaSetDevice(0);
                const float ms = ms0 > ms1 ? ms0 : ms1;
                if (it > 0 && ms < best) best = ms;
            }
            printf("both directions st.cs chunk=%5d B  %.3f ms  %.1f GB/s per direction\n", 16 << c, best, (double)(sizeof(double2) << nb) / 1e9 / (best * 1e-3));
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
