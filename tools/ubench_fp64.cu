// FP64 issue-rate micro-benchmark (the FP64 peak is not in MEASURED_PEAKS.json): DFMA, and the DMUL + DADD
// pairs that the estimation kernels actually execute (they are compiled with -fmad=false to follow the reference's
// unfused arithmetic).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench_fp64.cu -o tools/_bin/ubench_fp64
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void k(double* out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = __fma_rn(x[i], a, b);
            else x[i] = __dadd_rn(__dmul_rn(x[i], a), b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> static void run(const char* name, int sms)
{
    const int ctas = sms * 8, thr = 256, iters = 4096;
    double* out; cudaMalloc(&out, (size_t)ctas * thr * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<ctas, thr>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 5.0 * ctas * thr * (double)iters * 8 * 2;
    printf("%s: %.2f TFLOP/s (%.3f ms per launch)\n", name, flops / (ms * 1e-3) / 1e12, ms / 5);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<0>("DFMA (2 flops per instruction)", p.multiProcessorCount);
    run<1>("DMUL + DADD (unfused, 2 flops per 2 instructions)", p.multiProcessorCount);
    return 0;
}
