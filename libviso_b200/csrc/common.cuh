/*
 * common.cuh -- warp / block primitives shared by the kernel translation units.
 */
#ifndef VISO_COMMON_CUH_
#define VISO_COMMON_CUH_

#include <cuda_runtime.h>
#include <math_constants.h>
#include <limits.h>

#define FULL 0xffffffffu


__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ unsigned warp_sum_u(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

/* ordered compaction step for a CTA of (blockDim.x) threads: returns this thread's output slot (valid only when
 * flag) and advances *base_io (a per-thread copy of the running total, identical in all threads). */
__device__ __forceinline__ int block_compact_slot(bool flag, int& base_io, int* warp_tot /* smem[32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned b = __ballot_sync(FULL, flag);
    int pre = __popc(b & ((1u << lane) - 1));
    if (lane == 0) warp_tot[warp] = __popc(b);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        int c = warp_tot[w];
        if (w < warp) off += c;
        tot += c;
    }
    __syncthreads();
    int slot = base_io + off + pre;
    base_io += tot;
    return slot;
}

/* candidate-grid cell of a coordinate (clamped into the border cells) */
__device__ __forceinline__ int cell_coord(float v, int g)
{
    int c = __float2int_rd(v * (1.0f / VISO_GRID_CS));
    return min(max(c, 0), g - 1);
}

#endif
