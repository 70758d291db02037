// Microbenchmark: HBM bandwidth of random S-byte chunks (the access pattern of a gather of row segments), read only and
// read + write, against sequential.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o chunk_bw chunk_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

// each group of S/16 threads moves U chunks (U independent 16-byte loads per thread in flight): chunk index -> pseudo-random
// source chunk (and destination chunk when WRITE); total_chunks is a power of two
template <bool WRITE, bool RANDOM, int U>
__global__ void __launch_bounds__(256) k_chunks(const uint4 *src, uint4 *dst, uint64_t n_chunks, uint32_t tpc /*threads per chunk*/, uint64_t total_chunks, uint4 *sink) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t c0 = (t / tpc) * U, lane = t % tpc;
    if (c0 >= n_chunks) return;
    uint4 v[U];
    uint64_t dc[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const uint64_t c = c0 + u;
        const uint64_t sc = RANDOM ? mix(c) & (total_chunks - 1) : c;
        dc[u] = RANDOM ? mix(c ^ 0x55555555) & (total_chunks - 1) : c;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + sc * tpc + lane));
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (WRITE) dst[dc[u] * tpc + lane] = v[u];
        else if (v[u].x == 0x12345678u && v[u].y == 0x9abcdef0u) *sink = v[u];
    }
}

int main() {
    const uint64_t bytes = 8ull << 30;
    uint4 *a, *b, *sink;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 64);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (uint32_t S : {256u, 512u, 1024u, 2048u, 4096u, 16384u}) {
        const uint32_t tpc = S / 16;
        const uint64_t total = bytes / S, n = total / 2;  // move 4 GiB per launch
        const uint32_t grid = (uint32_t)((n / 8 * tpc + 255) / 256);
        for (int mode = 0; mode < 4; mode++) {
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k_chunks<false, true, 8><<<grid, 256>>>(a, b, n, tpc, total, sink);
                if (mode == 1) k_chunks<true, true, 8><<<grid, 256>>>(a, b, n, tpc, total, sink);
                if (mode == 2) k_chunks<false, false, 8><<<grid, 256>>>(a, b, n, tpc, total, sink);
                if (mode == 3) k_chunks<true, false, 8><<<grid, 256>>>(a, b, n, tpc, total, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            const double moved = (double)n * S * ((mode & 1) ? 2 : 1);
            printf("{\"chunk_bytes\": %u, \"mode\": \"%s\", \"GBps\": %.1f}\n", S, mode == 0 ? "random read" : mode == 1 ? "random read + random write" : mode == 2 ? "sequential read" : "sequential copy", moved / best / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
