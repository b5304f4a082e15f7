// Micro-benchmark of the Harris detector kernels alone (tools, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -std=c++17 -Ilibviso_b200/csrc [-DHARRIS_...] tools/ubench_detect.cu -o tools/_bin/ubench_detect
// 250 frames x 2 images of 1241 x 376 blurred noise, 2040 features, KITTI bins; prints the average time of one launch pair.
#include "../libviso_b200/csrc/detect.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv)
{
    const int W = 1241, H = 376, NI = argc > 1 ? atoi(argv[1]) : 500, nbx = 24, nby = 5, nfeat = 2040;
    HarrisCfg c{};
    c.w = W; c.h = H; c.pitch = W; c.nbinx = nbx; c.nbiny = nby; c.sx = W / nbx; c.sy = H / nby; c.per = nfeat / (nbx * nby);
    c.k = 0.04f;
    const double sc = 1.0 / (16.0 * 3 * 255.0);
    c.f0 = (float)(6.0 * sc); c.f1 = (float)(4.0 * sc); c.f2 = (float)(1.0 * sc);
    const size_t bytes = (size_t)W * H, slots = (size_t)nbx * nby * c.per;
    std::vector<unsigned char> img(bytes * 8);
    unsigned s = 12345;
    std::vector<float> nz(bytes * 8);
    for (auto& v : nz) { s = s * 1664525u + 1013904223u; v = (float)(s >> 24); }
    for (size_t i = 0; i < img.size(); ++i) {   // 3-tap blur along the row: correlated like a real image
        size_t a = i ? i - 1 : i, b = i + 1 < img.size() ? i + 1 : i;
        img[i] = (unsigned char)((nz[a] + 2 * nz[i] + nz[b]) / 4);
    }
    unsigned char* d_img; float2 *d_kp, *d_tmp; int *d_n, *d_bc, *d_flag; DetectJob* d_jobs;
    cudaMalloc(&d_img, bytes * NI); cudaMalloc(&d_kp, NI * slots * 8); cudaMalloc(&d_tmp, NI * slots * 8);
    cudaMalloc(&d_n, NI * 4); cudaMalloc(&d_bc, NI * nbx * nby * 4); cudaMalloc(&d_flag, 4); cudaMalloc(&d_jobs, NI * sizeof(DetectJob));
    for (int i = 0; i < NI; ++i) cudaMemcpy(d_img + i * bytes, img.data() + (i % 8) * bytes, bytes, cudaMemcpyHostToDevice);
    int one = 1; cudaMemcpy(d_flag, &one, 4, cudaMemcpyHostToDevice);
    std::vector<DetectJob> jobs(NI);
    for (int i = 0; i < NI; ++i)
        jobs[i] = DetectJob{d_img + i * bytes, d_kp + i * slots, d_n + i, d_tmp + i * slots, nullptr, nullptr, d_bc + i * nbx * nby, d_flag};
    cudaMemcpy(d_jobs, jobs.data(), NI * sizeof(DetectJob), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) viso_launch_detect(d_jobs, NI, c, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) viso_launch_detect(d_jobs, NI, c, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<int> n(NI); cudaMemcpy(n.data(), d_n, NI * 4, cudaMemcpyDeviceToHost);
    printf("%s: %d images %.3f ms per launch (%.2f us per image), n[0]=%d err=%s\n", argv[0], NI, ms / reps, 1e3 * ms / reps / NI, n[0],
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
