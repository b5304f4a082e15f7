/*
 * geometry.cu -- elementwise and small-matrix geometry: triangulate_rectified (viso.cpp:1137-1162, mvg.cpp:172-192),
 * projectPoints (viso.cpp:326-333), collect_matches (viso.cpp:501-514), triangulate_dlt (mvg.cpp:124-169),
 * solveRigidMotion (estimation.cpp:29-51), with their launch wrappers.
 */
#include "viso_dev.h"
#include "common.cuh"

/* triangulate_rectified<double>, viso.cpp:1146-1152 */
__global__ void triangulate_f64_kernel(const double* __restrict__ x, int m, int stride, double* __restrict__ X, ParamDev P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double u1 = x[i], v1 = x[stride + i], u2 = x[2 * stride + i];
    const double d = u1 - u2;
    X[i] = P.base * (u1 - P.cu) / d;
    X[stride + i] = P.base * (v1 - P.cv) / d;
    X[2 * stride + i] = P.f * P.base / d;
}

/* triangulate_rectified (float), mvg.cpp:184-190 */
__global__ void triangulate_f32_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int m, double f,
                                       double base, double c1u, double c1v, float* __restrict__ X)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double d = fmaxf(__fsub_rn(x1[i], x2[i]), 0.0001f);
    X[i] = (float)((x1[i] - c1u) * base / d);
    X[m + i] = (float)((x1[m + i] - c1v) * base / d);
    X[2 * m + i] = (float)(f * base / d);
}

/* projectPoints, viso.cpp:326-333: x = h2e(P * e2h(X)); w == 0 raises the error flag (misc.h:118-119) */
__global__ void project_kernel(const double* __restrict__ X, int n, const double* __restrict__ Pm, double* __restrict__ x,
                               int* err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double Xh[4] = {X[i], X[n + i], X[2 * n + i], 1.0};
    double xh[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += Pm[r * 4 + k] * Xh[k];
        xh[r] = s;
    }
    if (fabs(xh[2]) == 0) { atomicOr(err, 4); return; }
    x[i] = xh[0] / xh[2];
    x[n + i] = xh[1] / xh[2];
}

/* collect_matches (Mat x, 4 x m), viso.cpp:501-514, + triangulate_rectified<double>, viso.cpp:1146-1152 */
__global__ void collect_tri_kernel(const float2* __restrict__ kp1, int n1, const float2* __restrict__ kp2, int n2,
                                   const int* __restrict__ matches, int m, double* __restrict__ x, double* __restrict__ X,
                                   ParamDev P, int* err)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m) return;
    const int i1 = matches[3 * p], i2 = matches[3 * p + 1];
    if (i1 < 0 || i1 >= n1 || i2 < 0 || i2 >= n2) { atomicOr(err, 8); return; }
    const float2 a = kp1[i1], b = kp2[i2];
    const double u1 = a.x, v1 = a.y, u2 = b.x, v2 = b.y;
    if (x) { x[p] = u1; x[m + p] = v1; x[2 * m + p] = u2; x[3 * m + p] = v2; }
    if (X) {
        const double d = u1 - u2;
        X[p] = P.base * (u1 - P.cu) / d;
        X[m + p] = P.base * (v1 - P.cv) / d;
        X[2 * m + p] = P.f * P.base / d;
    }
}

/* ------------------------------------------------------------------------------------------------ small-matrix geometry */

/* cyclic Jacobi eigen-decomposition of a symmetric N x N matrix (N <= 4) in registers; V's columns = eigenvectors */
template <int N>
__device__ __forceinline__ void jacobi_eig(double A[N][N], double V[N][N])
{
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
#pragma unroll
        for (int p = 0; p < N; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) off += A[p][q] * A[p][q];
        if (off < 1e-300) break;
#pragma unroll
        for (int p = 0; p < N; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p][q];
                if (apq == 0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), sn = t * c;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
}

/*
 * triangulate_dlt, mvg.cpp:124-169: per point the 4x4 system A = [x1*P1(2,:)-P1(0,:); y1*P1(2,:)-P1(1,:); same for
 * camera 2] in double, X = last right singular vector / its 4th component (d = 1 when |vt33| < DBL_MIN, :163).  The
 * singular vector is the eigenvector of A^T A with the smallest eigenvalue (cyclic Jacobi); it is unique only up to
 * rounding, so parity with cv::SVD / the oracle is a tolerance, not bit equality.  One thread per point.
 */
__global__ void __launch_bounds__(128) triangulate_dlt_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int m,
                                                              const double* __restrict__ P1, const double* __restrict__ P2,
                                                              float* __restrict__ X)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double A[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        A[0][k] = x1[i] * P1[8 + k] - P1[k];
        A[1][k] = x1[m + i] * P1[8 + k] - P1[4 + k];
        A[2][k] = x2[i] * P2[8 + k] - P2[k];
        A[3][k] = x2[m + i] * P2[8 + k] - P2[4 + k];
    }
    double M[4][4], V[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += A[k][r] * A[k][c];
            M[r][c] = s;
        }
    jacobi_eig<4>(M, V);
    double v0 = V[0][0], v1 = V[1][0], v2 = V[2][0], v3 = V[3][0], best = M[0][0];
#pragma unroll
    for (int c = 1; c < 4; ++c)
        if (M[c][c] < best) { best = M[c][c]; v0 = V[0][c]; v1 = V[1][c]; v2 = V[2][c]; v3 = V[3][c]; }
    const double d = (fabs(v3) < 2.2250738585072014e-308) ? 1.0 : v3;
    X[i] = (float)((float)v0 / d);
    X[m + i] = (float)((float)v1 / d);
    X[2 * m + i] = (float)((float)v2 / d);
}

/*
 * solveRigidMotion, estimation.cpp:29-51 (Kabsch): C = (A - mean A)(B - mean B)^T, C = U S V^T,
 * R = U diag(1, 1, det(U V^T)) V^T, t = mean A - R mean B, i.e. T maps B onto A.  One CTA: block-reduced means and
 * covariance in double, thread 0 does the 3x3 SVD (Jacobi on C^T C, U = C V / sigma with Gram-Schmidt completion).
 * T: 4 x 4 float row-major.
 */
__global__ void __launch_bounds__(256) rigid_motion_kernel(const float* __restrict__ A, const float* __restrict__ B, int n,
                                                           float* __restrict__ T)
{
    __shared__ double red[256];
    __shared__ double stat[15]; /* mean A (3), mean B (3), C (9) */
    auto block_sum = [&](double v) {
        red[threadIdx.x] = v;
        __syncthreads();
        for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        const double r = red[0];
        __syncthreads();
        return r;
    };
    for (int r = 0; r < 3; ++r) {
        double sa = 0, sb = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { sa += A[r * n + i]; sb += B[r * n + i]; }
        const double ta = block_sum(sa), tb = block_sum(sb);
        if (threadIdx.x == 0) { stat[r] = ta / n; stat[3 + r] = tb / n; }
    }
    __syncthreads();
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int i = threadIdx.x; i < n; i += blockDim.x) s += (A[r * n + i] - stat[r]) * (B[c * n + i] - stat[3 + c]);
            const double t = block_sum(s);
            if (threadIdx.x == 0) stat[6 + r * 3 + c] = t;
        }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double C[3][3], M[3][3], V[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C[r][c] = stat[6 + r * 3 + c];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += C[k][r] * C[k][c];
            M[r][c] = s;
        }
    jacobi_eig<3>(M, V);
    int ord[3] = {0, 1, 2}; /* descending eigenvalue */
    for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (M[ord[b]][ord[b]] > M[ord[a]][ord[a]]) { const int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double Vs[3][3], U[3][3];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][ord[c]];
    for (int c = 0; c < 3; ++c) {
        double u[3];
        for (int r = 0; r < 3; ++r) { double s = 0; for (int k = 0; k < 3; ++k) s += C[r][k] * Vs[k][c]; u[r] = s; }
        for (int pc = 0; pc < c; ++pc) {
            double dot = 0;
            for (int r = 0; r < 3; ++r) dot += u[r] * U[r][pc];
            for (int r = 0; r < 3; ++r) u[r] -= dot * U[r][pc];
        }
        double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (nrm < 1e-12) { /* rank deficient: complete the basis */
            if (c == 2) {
                u[0] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
                u[1] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
                u[2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
            } else {
                double e[3] = {0, 0, 0};
                e[c == 0 ? 0 : (fabs(U[0][0]) < 0.9 ? 0 : 1)] = 1;
                for (int pc = 0; pc < c; ++pc) {
                    double dot = 0;
                    for (int r = 0; r < 3; ++r) dot += e[r] * U[r][pc];
                    for (int r = 0; r < 3; ++r) e[r] -= dot * U[r][pc];
                }
                for (int r = 0; r < 3; ++r) u[r] = e[r];
            }
            nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        }
        for (int r = 0; r < 3; ++r) U[r][c] = u[r] / nrm;
    }
    double UVt[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += U[r][k] * Vs[c][k];
            UVt[r][c] = s;
        }
    const double det = UVt[0][0] * (UVt[1][1] * UVt[2][2] - UVt[1][2] * UVt[2][1]) -
                       UVt[0][1] * (UVt[1][0] * UVt[2][2] - UVt[1][2] * UVt[2][0]) +
                       UVt[0][2] * (UVt[1][0] * UVt[2][1] - UVt[1][1] * UVt[2][0]);
    const double dg[3] = {1, 1, det};
    double R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += U[r][k] * dg[k] * Vs[c][k];
            R[r][c] = s;
        }
    for (int r = 0; r < 3; ++r) {
        double t = stat[r];
        for (int k = 0; k < 3; ++k) t -= R[r][k] * stat[3 + k];
        for (int c = 0; c < 3; ++c) T[r * 4 + c] = (float)R[r][c];
        T[r * 4 + 3] = (float)t;
    }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

/* ------------------------------------------------------------------------------------------------ launchers */

cudaError_t viso_launch_triangulate_f64(const double* x, int m, int stride, double* X, ParamDev p, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    triangulate_f64_kernel<<<(m + 255) / 256, 256, 0, s>>>(x, m, stride, X, p);
    return cudaGetLastError();
}

cudaError_t viso_launch_triangulate_f32(const float* x1, const float* x2, int m, double f, double base, double c1u,
                                        double c1v, float* X, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    triangulate_f32_kernel<<<(m + 255) / 256, 256, 0, s>>>(x1, x2, m, f, base, c1u, c1v, X);
    return cudaGetLastError();
}

cudaError_t viso_launch_project(const double* X, int n, const double* P, double* x, int* err_flag, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    project_kernel<<<(n + 255) / 256, 256, 0, s>>>(X, n, P, x, err_flag);
    return cudaGetLastError();
}

cudaError_t viso_launch_collect_tri(const float2* kp1, int n1, const float2* kp2, int n2, const int* matches, int m,
                                    double* x, double* X, ParamDev p, int* err_flag, cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    collect_tri_kernel<<<(m + 255) / 256, 256, 0, s>>>(kp1, n1, kp2, n2, matches, m, x, X, p, err_flag);
    return cudaGetLastError();
}

cudaError_t viso_launch_triangulate_dlt(const float* x1, const float* x2, int m, const double* P1, const double* P2, float* X,
                                        cudaStream_t s)
{
    if (m <= 0) return cudaSuccess;
    triangulate_dlt_kernel<<<(m + 127) / 128, 128, 0, s>>>(x1, x2, m, P1, P2, X);
    return cudaGetLastError();
}

cudaError_t viso_launch_rigid_motion(const float* A, const float* B, int n, float* T, cudaStream_t s)
{
    rigid_motion_kernel<<<1, 256, 0, s>>>(A, B, n, T);
    return cudaGetLastError();
}
