// Instruction-throughput microbenchmark for the packed-integer SAD building blocks on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_simd ubench_simd.cu && ./ubench_simd
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void k(unsigned* out, unsigned seed)
{
    unsigned a0 = threadIdx.x * 2654435761u + seed, a1 = a0 ^ 0x9e3779b9u, a2 = a0 * 3u, a3 = a1 * 5u;
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    unsigned b = blockIdx.x * 40503u + seed;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) { c0 += __vsadu4(a0, b); c1 += __vsadu4(a1, b); c2 += __vsadu4(a2, b); c3 += __vsadu4(a3, b); }
            if (OP == 1) { c0 += __vminu2(a0, b); c1 += __vminu2(a1, b); c2 += __vminu2(a2, b); c3 += __vminu2(a3, b); }
            if (OP == 2) { c0 = __usad(a0, b, c0); c1 = __usad(a1, b, c1); c2 = __usad(a2, b, c2); c3 = __usad(a3, b, c3); }
            if (OP == 3) { c0 = __vminu2(c0, a0 + u); c1 = __vminu2(c1, a1 + u); c2 = __vminu2(c2, a2 + u); c3 = __vminu2(c3, a3 + u); }
            if (OP == 4) { c0 += a0 ^ b; c1 += a1 ^ b; c2 += a2 ^ b; c3 += a3 ^ b; }
            b += 0x01010101u * (u + 1);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3;
}

template <int OP> void run(const char* name, int ops_per_inner)
{
    unsigned* d;
    cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 8, 1024>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<148 * 8, 1024>>>(d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = 148.0 * 8 * 32 * ITERS * 8 * ops_per_inner;
    printf("%-28s %8.3f ms  %7.2f G warp-instr/s  = %5.2f warp-instr/clk/SM @1.965GHz (counting %d instr per inner op group)\n", name, ms,
           warp_instr / ms / 1e6, warp_instr / ms / 1e6 / 148 / 1.965, ops_per_inner);
    cudaFree(d);
}

int main()
{
    run<0>("VABSDIFF4.U8.ACC x4 (+1 add)", 5);
    run<1>("VIMNMX.U16x2 + IADD x4 (+1)", 9);
    run<2>("VABSDIFF.U32 acc x4 (+1)", 5);
    run<3>("VIMNMX.U16x2 + IADD(u) x4 (+1)", 9);
    run<4>("LOP3 + IADD x4 (+1)", 9);
    return 0;
}
