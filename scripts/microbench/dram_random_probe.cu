// Random-granule DRAM bandwidth of one B200 (development aid; not part of the product library).
// The step kernel's traffic is isolated 32-byte sectors and short runs of them, not streams, so the copy bandwidth in
// MEASURED_PEAKS.json overstates what such a pattern can reach.  This probe reads (and optionally writes back) granules
// of G bytes at random G-aligned offsets of a buffer much larger than L2, with every SM fully occupied and several
// independent loads in flight per thread, and reports GB/s per granule size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/dram_random_probe scripts/microbench/dram_random_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// lanesPerGranule = G / 16 consecutive lanes read one granule (16 bytes each)
template <int UNROLL, bool WRITE>
__global__ void probe(uint4 *buf, uint64_t nGranules, int lanesPerGranule, int iters, uint64_t seed, unsigned *sink) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t grp = tid / lanesPerGranule, sub = tid % lanesPerGranule;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint4 v[UNROLL];
        uint64_t off[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const uint64_t g = mix(seed + grp * 0x9E3779B97F4A7C15ULL + (uint64_t)(it * UNROLL + k)) % nGranules;
            off[k] = g * lanesPerGranule + sub;
            v[k] = buf[off[k]];
        }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            acc += v[k].x ^ v[k].w;
            if (WRITE) { v[k].x += 1; buf[off[k]] = v[k]; }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main(int argc, char **argv) {
    const double gb = argc > 1 ? atof(argv[1]) : 16.0;
    const size_t bytes = (size_t)(gb * (1ull << 30));
    uint4 *buf; unsigned *sink;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int threads = 256, blocks = 148 * 8 * 4, iters = 256;
    constexpr int U = 4;
    printf("{\"buffer_gb\": %.1f, \"rows\": [\n", gb);
    bool first = true;
    for (int w = 0; w < 2; ++w)
        for (int G = 32; G <= 1024; G *= 2) {
            const int lpg = G / 16;
            const uint64_t nGran = bytes / G;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (w) probe<U, true><<<blocks, threads>>>(buf, nGran, lpg, iters, 1234 + rep, sink);
                else probe<U, false><<<blocks, threads>>>(buf, nGran, lpg, iters, 1234 + rep, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double moved = (double)blocks * threads * iters * U * 16.0 * (w ? 2.0 : 1.0);
            printf("%s  {\"granule_bytes\": %d, \"mode\": \"%s\", \"ms\": %.3f, \"GBps\": %.1f}", first ? "" : ",\n", G,
                   w ? "read+writeback" : "read", best, moved / best / 1e6);
            first = false;
        }
    printf("\n]}\n");
    return cudaGetLastError() != cudaSuccess;
}
