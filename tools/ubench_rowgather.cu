// The memory roof of sad_match's access pattern, without the arithmetic: how many 256-byte descriptor rows per second
// can a B200 gather through L1 when they are read the way eval_batch reads them -- eight lanes per row (lane `sub`
// takes the 16-byte chunks sub and 8 + sub, so every LDG.128 covers four full 128-byte lines), VISO_EVAL_DEPTH x 2
// loads in flight per warp, 4 warps per CTA, 8 CTAs per SM, 15.5 KB of shared memory per CTA taken from L1 -- and
// with sad_match's reuse: a CTA ("tile") draws its rows from 250 candidates scattered over one 2040-row set
// (one image's descriptors, 522 KB) and reads each about 4.3 times (27 queries x 40 candidates).
// Sets are 2000 x 2040 rows = 1 GB, as in the 1000-frame benchmark (nothing survives in L2 between tiles of
// different sets; tiles of the same set share through L2 as in the real kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_rowgather.cu -o tools/_bin/ubench_rowgather
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define ROWS_PER_SET 2040
#define CAND 250
#define QUERIES 27
#define PER_QUERY 40
#ifndef DEPTH
#define DEPTH 2
#endif

__global__ void __launch_bounds__(128, 8) gather(const uint4* __restrict__ desc, const unsigned short* __restrict__ cand,
                                                  const unsigned short* __restrict__ lists, unsigned* out, int tiles_per_set)
{
    extern __shared__ unsigned short s_list[];   // QUERIES x PER_QUERY staged indices + padding to 15.5 KB
    const int tile = blockIdx.x, set = tile / tiles_per_set;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 7, g = lane >> 3;
    const unsigned short* c = cand + (size_t)tile * CAND;
    const unsigned short* l = lists + (size_t)tile * QUERIES * PER_QUERY;
    for (int i = threadIdx.x; i < QUERIES * PER_QUERY; i += blockDim.x) s_list[i] = c[l[i]];   // row index inside the set
    __syncthreads();
    const uint4* base = desc + (size_t)set * ROWS_PER_SET * 16 + sub;
    unsigned acc = 0;
    for (int q = warp; q < QUERIES; q += 4) {
        const unsigned short* ql = s_list + q * PER_QUERY;
        for (int b = 0; b < PER_QUERY; b += 4 * DEPTH) {   // 4 rows per step, DEPTH steps in flight
            uint4 ra[DEPTH], rb[DEPTH];
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) {
                const int e = min(b + 4 * s + g, PER_QUERY - 1);
                const uint4* rp = base + (size_t)ql[e] * 16;
                ra[s] = __ldg(rp); rb[s] = __ldg(rp + 8);
            }
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) acc += ra[s].x ^ ra[s].y ^ ra[s].z ^ ra[s].w ^ rb[s].x ^ rb[s].y ^ rb[s].z ^ rb[s].w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;   // keeps the loads alive
}


// Variant for the round-2 design question: the tile's 250 candidate rows are first copied into shared memory
// (64 KB per CTA, coalesced), the list-driven reads then hit shared memory.  WARPS warps per CTA, as many CTAs per SM
// as 64 KB + lists allow (3).
#ifndef CWARPS
#define CWARPS 8
#endif
__global__ void __launch_bounds__(CWARPS * 32) gather_cached(const uint4* __restrict__ desc, const unsigned short* __restrict__ cand,
                                                               const unsigned short* __restrict__ lists, unsigned* out, int tiles_per_set)
{
    extern __shared__ uint4 s_rows[];            // CAND x 16 uint4, then QUERIES x PER_QUERY list entries (staged slots)
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_rows + CAND * 16);
    const int tile = blockIdx.x, set = tile / tiles_per_set;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 7, g = lane >> 3;
    const unsigned short* c = cand + (size_t)tile * CAND;
    const unsigned short* l = lists + (size_t)tile * QUERIES * PER_QUERY;
    for (int i = threadIdx.x; i < QUERIES * PER_QUERY; i += blockDim.x) s_list[i] = l[i];     // slot inside the cache
    const uint4* setbase = desc + (size_t)set * ROWS_PER_SET * 16;
    for (int i = threadIdx.x; i < CAND * 16; i += blockDim.x) s_rows[i] = __ldg(setbase + (size_t)c[i >> 4] * 16 + (i & 15));
    __syncthreads();
    unsigned acc = 0;
    for (int q = warp; q < QUERIES; q += CWARPS) {
        const unsigned short* ql = s_list + q * PER_QUERY;
        for (int b = 0; b < PER_QUERY; b += 4 * DEPTH) {
            uint4 ra[DEPTH], rb[DEPTH];
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) {
                const int e = min(b + 4 * s + g, PER_QUERY - 1);
                const uint4* rp = s_rows + (size_t)ql[e] * 16 + sub;
                ra[s] = rp[0]; rb[s] = rp[8];
            }
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) acc += ra[s].x ^ ra[s].y ^ ra[s].z ^ ra[s].w ^ rb[s].x ^ rb[s].y ^ rb[s].z ^ rb[s].w;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main()
{
    const int sets = 2000, tiles_per_set = 78 * 3 / 2, tiles = sets * tiles_per_set;   // ~ one job and a half per set
    uint4* desc; cudaMalloc(&desc, (size_t)sets * ROWS_PER_SET * 256);
    cudaMemset(desc, 1, (size_t)sets * ROWS_PER_SET * 256);
    std::vector<unsigned short> cand((size_t)tiles * CAND), lists((size_t)tiles * QUERIES * PER_QUERY);
    unsigned s = 7;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (int t = 0; t < tiles; ++t) {
        // a tile's candidates: 14 grid rows x ~18 consecutive cell-sorted records each, like a staged neighbourhood
        for (int r = 0; r < 14; ++r) {
            const int start = rnd() % (ROWS_PER_SET - 18);
            for (int k = 0; k < 18 && r * 18 + k < CAND; ++k) cand[(size_t)t * CAND + r * 18 + k] = (unsigned short)(start + k);
        }
        for (int i = 0; i < QUERIES * PER_QUERY; ++i) lists[(size_t)t * QUERIES * PER_QUERY + i] = (unsigned short)(rnd() % CAND);
    }
    unsigned short *d_cand, *d_lists; unsigned* d_out;
    cudaMalloc(&d_cand, cand.size() * 2); cudaMalloc(&d_lists, lists.size() * 2); cudaMalloc(&d_out, 4);
    cudaMemcpy(d_cand, cand.data(), cand.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_lists, lists.data(), lists.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 15872;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather<<<tiles, 128, smem>>>(desc, d_cand, d_lists, d_out, tiles_per_set);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 3;
    for (int r = 0; r < reps; ++r) gather<<<tiles, 128, smem>>>(desc, d_cand, d_lists, d_out, tiles_per_set);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rows = (double)tiles * QUERIES * PER_QUERY;
    printf("%d tiles, %.1f M row reads per launch: %.3f ms -> %.1f G rows/s = %.2f TB/s through L1 (%s)\n", tiles, rows / 1e6, ms / reps,
           rows / (ms / reps * 1e-3) / 1e9, rows * 256 / (ms / reps * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    {
        const size_t smem2 = (size_t)CAND * 256 + QUERIES * PER_QUERY * 2;
        cudaFuncSetAttribute(gather_cached, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        gather_cached<<<tiles, CWARPS * 32, smem2>>>(desc, d_cand, d_lists, d_out, tiles_per_set);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) gather_cached<<<tiles, CWARPS * 32, smem2>>>(desc, d_cand, d_lists, d_out, tiles_per_set);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("shared-memory row cache (%d warps per CTA, %zu B per CTA): %.3f ms -> %.1f G rows/s (%s)\n", CWARPS, smem2, ms / reps,
               rows / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
