/*
 * estimation.cu -- RANSAC + Gauss-Newton on the stereo reprojection error (compute_J, get_inliers, minimize_reproj,
 * ransac_minimize_reproj: viso.cpp:1401-1623) with their launch wrappers.
 *
 * Built with -fmad=false: the FP64 code must evaluate exactly the reference's expressions (separate multiply and
 * add, as a stock x86-64 build of the reference does) so that Jacobians, normal equations, LU pivots and inlier
 * decisions agree bit for bit with the CPU path.  sin / cos follow glibc's libm operation for operation
 * (glibc_sincos.h), so the rotation entries are the CPU's as well.
 */
#include "viso_dev.h"
#include "common.cuh"
#include "glibc_sincos.h"

#include <algorithm>
#include <cstdlib>

/* ------------------------------------------------------------------------------------------------ estimation */

struct Rot {
    double r00, r01, r02, r10, r11, r12, r20, r21, r22;
    double rdrx10, rdrx11, rdrx12, rdrx20, rdrx21, rdrx22;
    double rdry00, rdry01, rdry02, rdry10, rdry11, rdry12, rdry20, rdry21, rdry22;
    double rdrz00, rdrz01, rdrz10, rdrz11, rdrz20, rdrz21;
    double tx, ty, tz;
};

/* viso.cpp:1406-1424: the rotation (and its derivatives) from the six trigonometric values */
__device__ __forceinline__ void rot_from_sincos(double sx, double cx, double sy, double cy, double sz, double cz,
                                                const double* tr, Rot& R, bool derivs)
{
    R.tx = tr[3]; R.ty = tr[4]; R.tz = tr[5];
    R.r00 = +cy * cz;                R.r01 = -cy * sz;                R.r02 = +sy;
    R.r10 = +sx * sy * cz + cx * sz; R.r11 = -sx * sy * sz + cx * cz; R.r12 = -sx * cy;
    R.r20 = -cx * sy * cz + sx * sz; R.r21 = +cx * sy * sz + sx * cz; R.r22 = +cx * cy;
    if (derivs) {
        R.rdrx10 = +cx * sy * cz - sx * sz; R.rdrx11 = -cx * sy * sz - sx * cz; R.rdrx12 = -cx * cy;
        R.rdrx20 = +sx * sy * cz + cx * sz; R.rdrx21 = -sx * sy * sz + cx * cz; R.rdrx22 = -sx * cy;
        R.rdry00 = -sy * cz;      R.rdry01 = +sy * sz;      R.rdry02 = +cy;
        R.rdry10 = +sx * cy * cz; R.rdry11 = -sx * cy * sz; R.rdry12 = +sx * sy;
        R.rdry20 = -cx * cy * cz; R.rdry21 = +cx * cy * sz; R.rdry22 = -cx * sy;
        R.rdrz00 = -cy * sz;                R.rdrz01 = -cy * cz;
        R.rdrz10 = -sx * sy * sz + cx * cz; R.rdrz11 = -sx * sy * cz - cx * sz;
        R.rdrz20 = +cx * sy * sz + sx * cz; R.rdrz21 = +cx * sy * cz - sx * sz;
    }
}

__device__ __forceinline__ void make_rot(const double* tr, Rot& R, bool derivs)
{
    const double rx = tr[0], ry = tr[1], rz = tr[2];
    const double sx = viso_sc::sin_glibc(rx), cx = viso_sc::cos_glibc(rx), sy = viso_sc::sin_glibc(ry);
    const double cy = viso_sc::cos_glibc(ry), sz = viso_sc::sin_glibc(rz), cz = viso_sc::cos_glibc(rz);
    rot_from_sincos(sx, cx, sy, cy, sz, cz, tr, R, derivs);
}

/* prediction of one point, viso.cpp:1441-1443, 1452, 1486-1489 */
__device__ __forceinline__ void predict_point(const Rot& R, const ParamDev& P, double X1p, double Y1p, double Z1p,
                                              double pred[4])
{
    const double X1c = R.r00 * X1p + R.r01 * Y1p + R.r02 * Z1p + R.tx;
    const double Y1c = R.r10 * X1p + R.r11 * Y1p + R.r12 * Z1p + R.ty;
    const double Z1c = R.r20 * X1p + R.r21 * Y1p + R.r22 * Z1p + R.tz;
    const double X2c = X1c - P.base;
    pred[0] = P.f * X1c / Z1c + P.cu;
    pred[1] = P.f * Y1c / Z1c + P.cv;
    pred[2] = P.f * X2c / Z1c + P.cu;
    pred[3] = P.f * Y1c / Z1c + P.cv;
}

/* inlier test, viso.cpp:1527-1533 */
__device__ __forceinline__ bool inlier_point(const Rot& R, const ParamDev& P, const double* X, const double* obs,
                                             int stride, int i)
{
    double pred[4];
    predict_point(R, P, X[i], X[stride + i], X[2 * stride + i], pred);
    const double e0 = obs[i] - pred[0], e1 = obs[stride + i] - pred[1];
    const double e2 = obs[2 * stride + i] - pred[2], e3 = obs[3 * stride + i] - pred[3];
    const double err2 = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    return err2 < P.thr2;
}

/* a point in the two camera frames, viso.cpp:1441-1445 */
struct CamPt { double X1c, Y1c, Z1c, X2c; };
__device__ __forceinline__ CamPt cam_point(const Rot& R, const ParamDev& P, double X1p, double Y1p, double Z1p)
{
    CamPt c;
    c.X1c = R.r00 * X1p + R.r01 * Y1p + R.r02 * Z1p + R.tx;
    c.Y1c = R.r10 * X1p + R.r11 * Y1p + R.r12 * Z1p + R.ty;
    c.Z1c = R.r20 * X1p + R.r21 * Y1p + R.r22 * Z1p + R.tz;
    c.X2c = c.X1c - P.base;
    return c;
}

/* column j (0..5) of the four Jacobian rows of one point, viso.cpp:1455-1484 (literal; no shortcuts for the constant
 * derivative columns so that non-finite inputs propagate exactly as in the reference) */
__device__ __forceinline__ void jac_column(const Rot& R, const ParamDev& P, const CamPt& c, double X1p, double Y1p,
                                           double Z1p, double weight, int j, double col[4])
{
    double X1cd, Y1cd, Z1cd;
    switch (j) {
    case 0: X1cd = 0;
        Y1cd = R.rdrx10 * X1p + R.rdrx11 * Y1p + R.rdrx12 * Z1p;
        Z1cd = R.rdrx20 * X1p + R.rdrx21 * Y1p + R.rdrx22 * Z1p;
        break;
    case 1: X1cd = R.rdry00 * X1p + R.rdry01 * Y1p + R.rdry02 * Z1p;
        Y1cd = R.rdry10 * X1p + R.rdry11 * Y1p + R.rdry12 * Z1p;
        Z1cd = R.rdry20 * X1p + R.rdry21 * Y1p + R.rdry22 * Z1p;
        break;
    case 2: X1cd = R.rdrz00 * X1p + R.rdrz01 * Y1p;
        Y1cd = R.rdrz10 * X1p + R.rdrz11 * Y1p;
        Z1cd = R.rdrz20 * X1p + R.rdrz21 * Y1p;
        break;
    case 3: X1cd = 1; Y1cd = 0; Z1cd = 0; break;
    case 4: X1cd = 0; Y1cd = 1; Z1cd = 0; break;
    default: X1cd = 0; Y1cd = 0; Z1cd = 1; break;
    }
    col[0] = weight * P.f * (X1cd * c.Z1c - c.X1c * Z1cd) / (c.Z1c * c.Z1c);
    col[1] = weight * P.f * (Y1cd * c.Z1c - c.Y1c * Z1cd) / (c.Z1c * c.Z1c);
    col[2] = weight * P.f * (X1cd * c.Z1c - c.X2c * Z1cd) / (c.Z1c * c.Z1c);
    col[3] = weight * P.f * (Y1cd * c.Z1c - c.Y1c * Z1cd) / (c.Z1c * c.Z1c);
}

/* weighted residuals of one point (column 6 of its four rows), viso.cpp:1486-1494 */
__device__ __forceinline__ void residual_column(const ParamDev& P, const CamPt& c, double weight, const double ob[4],
                                                double col[4])
{
    double pred[4];
    pred[0] = P.f * c.X1c / c.Z1c + P.cu;
    pred[1] = P.f * c.Y1c / c.Z1c + P.cv;
    pred[2] = P.f * c.X2c / c.Z1c + P.cu;
    pred[3] = P.f * c.Y1c / c.Z1c + P.cv;
#pragma unroll
    for (int r = 0; r < 4; ++r) col[r] = weight * (ob[r] - pred[r]);
}

/* Jacobian rows + weighted residuals of one point, viso.cpp:1441-1495.  out: 4 rows x 7 (6 Jacobian columns + residual). */
__device__ __forceinline__ void point_rows(const Rot& R, const ParamDev& P, double X1p, double Y1p, double Z1p,
                                           double weight, const double ob[4], double out[4][7])
{
    const CamPt c = cam_point(R, P, X1p, Y1p, Z1p);
    double col[4];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        jac_column(R, P, c, X1p, Y1p, Z1p, weight, j, col);
#pragma unroll
        for (int r = 0; r < 4; ++r) out[r][j] = col[r];
    }
    residual_column(P, c, weight, ob, col);
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r][6] = col[r];
}

/* weight, viso.cpp:1449: read from observe column i (the LOOP index, not active[i]) */
__device__ __forceinline__ double weight_of(const ParamDev& P, double obs0_col_i)
{
    return 1.0 / (fabs(obs0_col_i - P.cu) / fabs(P.cu) + 0.05);
}

/* cv::solve(JtJ, Jtr, p, DECOMP_LU) == OpenCV hal LUImpl<double>, m = 6, one right-hand side (viso.cpp:1602).
 * A is the full symmetric 6x6.  Returns false when a pivot is < 100*DBL_EPSILON. */
__device__ __forceinline__ bool lu_solve6(double A[6][6], double b[6])
{
    const double eps = 2.220446049250313e-16 * 100;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        int k = i;
        double best = fabs(A[i][i]);
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            const double v = fabs(A[j][i]);
            if (v > best) { best = v; k = j; }
        }
        if (best < eps) return false;
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            if (k == j) {
#pragma unroll
                for (int c = i; c < 6; ++c) { const double t = A[i][c]; A[i][c] = A[j][c]; A[j][c] = t; }
                const double t = b[i]; b[i] = b[j]; b[j] = t;
            }
        }
        const double d = -1 / A[i][i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            const double alpha = A[j][i] * d;
#pragma unroll
            for (int c = i + 1; c < 6; ++c) A[j][c] += alpha * A[i][c];
            b[j] += alpha * b[i];
        }
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = b[i];
#pragma unroll
        for (int c = i + 1; c < 6; ++c) s -= A[i][c] * b[c];
        b[i] = s / A[i][i];
    }
    return true;
}


/* lu_solve6 spread over a warp: lane c < 6 holds column c of A (a[r] = A[r][c]), lane 6 the right-hand side; the other
 * lanes idle along.  Every scalar operation of lu_solve6 is performed once, by the lane that owns its result, on the same
 * operands in the same order -- the pivot search and -1 / pivot on the pivot column's lane, alpha = A[j][i] * d there as
 * well (broadcast), A[j][c] += alpha * A[i][c] on lane c, the back substitution on every lane alike from shuffled
 * operands -- so the solution is bit-identical; columns left of the pivot are swapped along with the rest, which
 * lu_solve6 does not do, but nothing reads them again.  x[6] and the return value are the same on all lanes. */
__device__ __forceinline__ bool lu_solve6_warp(const double* sums /* [27]: upper triangle row-major, then J^T r */, int lane,
                                               double x[6])
{
    const double eps = 2.220446049250313e-16 * 100;
    double a[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        int t = 21 + r; /* lane 6 (and the idle lanes): the right-hand side */
        if (lane < 6) {
            const int lo = min(r, lane), hi = max(r, lane);
            t = lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
        }
        a[r] = sums[t];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        int k = i;
        double best = fabs(a[i]);
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            const double v = fabs(a[j]);
            if (v > best) { best = v; k = j; }
        }
        k = __shfl_sync(FULL, k, i);
        if (__shfl_sync(FULL, best < eps ? 1 : 0, i)) return false;
#pragma unroll
        for (int j = i + 1; j < 6; ++j)
            if (k == j) { const double t = a[i]; a[i] = a[j]; a[j] = t; }
        const double d = __shfl_sync(FULL, -1 / a[i], i);
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
            const double alpha = __shfl_sync(FULL, a[j] * d, i);
            if (lane > i) a[j] += alpha * a[i];
        }
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double sv = __shfl_sync(FULL, a[i], 6);
#pragma unroll
        for (int c = i + 1; c < 6; ++c) sv -= __shfl_sync(FULL, a[i], c) * x[c];
        x[i] = sv / __shfl_sync(FULL, a[i], i);
    }
    return true;
}

/* NaN pivots: fabs(NaN) > best is false and best < eps is false, exactly like the reference's
 * std::abs comparisons -- the solve "succeeds" with NaN output (and viso.cpp:1610 then reports convergence). */

__device__ __forceinline__ void sample_from_seeds(const uint32_t* seeds, int N, int s[3])
{
    const uint32_t r0 = seeds[0], r1 = seeds[1], r2 = seeds[2];
    int a = (int)(((unsigned long long)r0 * (unsigned long long)N) >> 32);
    int b = (int)(((unsigned long long)r1 * (unsigned long long)(N - 1)) >> 32);
    int c = (int)(((unsigned long long)r2 * (unsigned long long)(N - 2)) >> 32);
    if (b >= a) b++;
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    if (c >= lo) c++;
    if (c >= hi) c++;
    int s0 = lo, s1 = hi, s2 = c;
    if (s2 < s0) { const int t = s2; s2 = s1; s1 = s0; s0 = t; }
    else if (s2 < s1) { const int t = s2; s2 = s1; s1 = t; }
    s[0] = s0; s[1] = s1; s[2] = s2;
}

__constant__ int c_pair_a[27] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 5, 0, 1, 2, 3, 4, 5};
__constant__ int c_pair_b[27] = {0, 1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 2, 3, 4, 5, 3, 4, 5, 4, 5, 5, 6, 6, 6, 6, 6, 6};

#ifndef VISO_HYP_PER_CTA
#define VISO_HYP_PER_CTA 8 /* hypotheses per CTA of ransac_hyp_kernel, 4 lanes each: one warp, so that a hypothesis that runs all 100
                             iterations (0.8 % do, and the kernel lasts as long as they) holds one warp's registers, not four */
#endif
#ifndef VISO_HYP_MINB
#define VISO_HYP_MINB 12
#endif

/*
 * FOUR LANES per hypothesis: tr = 0, 3-point sample, Gauss-Newton (minimize_reproj with 3 active points),
 * viso.cpp:1555-1562 + 1583-1623.
 *
 * There are only ransac_iter x frame pairs hypotheses (50 k per 1000-frame sequence), each a dependent FP64 chain
 * of a few thousand instructions per iteration (84 IEEE divisions, six sin / cos, a 6 x 6 LU): one thread per
 * hypothesis leaves the SMs at ~0.3 IPC.  A quad splits the iteration without changing a single operation:
 *   lanes 0..2  sin / cos of one angle each (same function, same argument as make_rot), broadcast by shuffle;
 *   lanes 0..2  the 4 Jacobian rows + residuals of one sample point each (point_rows), written to shared memory;
 *   lanes 0..3  7 of the 21 + 6 normal-equation sums each, every sum SEQUENTIALLY over rows 0..11 (bit-identical
 *               to cv::mulTransposed / J^T r);
 *   lane 0      the LU solve and the convergence test (viso.cpp:1602-1617), broadcast by shuffle.
 */
__global__ void __launch_bounds__(VISO_HYP_PER_CTA * 4, VISO_HYP_MINB)
ransac_hyp_kernel(const RansacProb* __restrict__ probs, ParamDev P, int it_cap, int* __restrict__ strag, int strag_cap)
{
    __shared__ double rows_s[VISO_HYP_PER_CTA][12][7];
    __shared__ double sums_s[VISO_HYP_PER_CTA][28];
    const RansacProb& pb = probs[blockIdx.y];
    const int hl = threadIdx.x >> 2, q = threadIdx.x & 3, lane = threadIdx.x & 31;
    const int hId = blockIdx.x * VISO_HYP_PER_CTA + hl;
    const unsigned qmask = 0xfu << (lane & ~3);
    const int q0 = lane & ~3; /* first lane of the quad */
    const int n = *pb.n;
    if (hId >= pb.H || n < pb.min_n || n < 1) return; /* quad uniform */
    int s[3];
    if (pb.table) { s[0] = pb.table[3 * hId]; s[1] = pb.table[3 * hId + 1]; s[2] = pb.table[3 * hId + 2]; }
    else sample_from_seeds(pb.seeds + 3 * hId, n, s);
    const int S = pb.stride;
    bool bad_index = false;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (s[i] < 0 || s[i] >= n) bad_index = true;
    /* this lane's sample point (lane 3 mirrors point 2 and does not write) */
    const int pi = q < 3 ? q : 2;
    const int a = bad_index ? 0 : s[pi];
    const double Xp = pb.X[a], Yp = pb.X[S + a], Zp = pb.X[2 * S + a];
    double ob[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) ob[r] = pb.obs[r * S + a];
    const double w = weight_of(P, pb.obs[min(pi, n - 1)]); /* columns 0,1,2: the LOOP index, viso.cpp:1449 */
    double tr[6] = {0, 0, 0, 0, 0, 0};
    int ok = 0;
    bool unfinished = !bad_index; /* left the loop at the cap without a verdict */
    if (!bad_index) {
        for (int it = 0; it < it_cap; ++it) {
            /* make_rot, viso.cpp:1406-1424: lane 0 -> rx, lane 1 -> ry, lane 2 -> rz */
            const double ang = tr[q < 3 ? q : 0];
            const double sv = viso_sc::sin_glibc(ang), cv = viso_sc::cos_glibc(ang);
            Rot R;
            {
                const double sx = __shfl_sync(qmask, sv, q0), cx = __shfl_sync(qmask, cv, q0);
                const double sy = __shfl_sync(qmask, sv, q0 + 1), cy = __shfl_sync(qmask, cv, q0 + 1);
                const double sz = __shfl_sync(qmask, sv, q0 + 2), cz = __shfl_sync(qmask, cv, q0 + 2);
                rot_from_sincos(sx, cx, sy, cy, sz, cz, tr, R, true);
            }
            if (q < 3) {
                double rows[4][7];
                point_rows(R, P, Xp, Yp, Zp, w, ob, rows);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 7; ++c) rows_s[hl][4 * q + r][c] = rows[r][c];
            }
            __syncwarp(qmask);
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int t = q + 4 * j;
                if (t < 27) {
                    const int sa = c_pair_a[t], sb = c_pair_b[t];
                    double acc = 0;
#pragma unroll
                    for (int k = 0; k < 12; ++k) acc += rows_s[hl][k][sa] * rows_s[hl][k][sb];
                    sums_s[hl][t] = acc;
                }
            }
            __syncwarp(qmask);
            int flag = 0; /* 0 continue, 1 converged, 2 singular */
            double p[6] = {0, 0, 0, 0, 0, 0};
            if (q == 0) {
                double A[6][6], b[6];
                int t = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = r; c < 6; ++c) { A[r][c] = sums_s[hl][t]; A[c][r] = sums_s[hl][t]; ++t; }
#pragma unroll
                for (int r = 0; r < 6; ++r) b[r] = sums_s[hl][21 + r];
                if (!lu_solve6(A, b)) flag = 2;
                else {
                    flag = 1;
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                        if (b[j] > P.thresh) flag = 0; /* fabs(p > thresh), viso.cpp:1610 */
#pragma unroll
                    for (int j = 0; j < 6; ++j) p[j] = b[j];
                }
            }
            flag = __shfl_sync(qmask, flag, q0);
            if (flag == 2) { ok = 0; unfinished = false; break; }
            if (flag == 1) { ok = 1; unfinished = false; break; }
#pragma unroll
            for (int j = 0; j < 6; ++j) tr[j] = tr[j] + __shfl_sync(qmask, p[j], q0);
            __syncwarp(qmask); /* rows_s / sums_s are rewritten by the next iteration */
        }
    }
    if (q == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) pb.hyp_tr[6 * hId + j] = tr[j];
        pb.hyp_ok[hId] = ok;
        pb.hyp_count[hId] = -1;
        if (unfinished && it_cap < 100) { /* ransac_hyp_cont_kernel takes it from here (tr after it_cap iterations) */
            const int slot = atomicAdd(strag, 1);
            if (slot < strag_cap) { strag[2 + 2 * slot] = blockIdx.y; strag[3 + 2 * slot] = hId; }
        }
    }
}

/*
 * The hypotheses ransac_hyp_kernel left unfinished at its iteration cap, ONE WARP each, iterations it0 .. 99.
 *
 * Why: 0.8 % of the three-point samples never converge and run all 100 iterations (viso.cpp:1590), 4.5 % need more than
 * 8 (median 4).  A quad executes ~3600 instructions per iteration whatever its neighbours do, so the 100-iteration
 * tail alone lasted ~0.8 ms (at 7 % warp occupancy) and set the duration of the launch.  Here the same operations are
 * spread over a warp instead of a quad -- nothing is re-associated, every value is computed by the expression the quad
 * kernel uses:
 *   lanes 0..2    sin and cos of one angle each, broadcast;
 *   lanes 0..17   (sample point, parameter j): column j of the point's four Jacobian rows (3 divisions per lane
 *                 instead of 28);
 *   lanes 18..20  the residual column of one point each;
 *   lanes 0..26   one normal-equation sum each, sequentially over rows 0..11;
 *   lane 0        the LU solve and the convergence test, broadcast.
 * The straggler list is (problem, hypothesis) pairs; the state is hyp_tr (tr after it0 iterations).
 */
#ifndef VISO_CONT_LU_WARP
#define VISO_CONT_LU_WARP 1
#endif
#ifndef VISO_CONT_WARPS
#define VISO_CONT_WARPS 2
#endif
__global__ void __launch_bounds__(VISO_CONT_WARPS * 32)
ransac_hyp_cont_kernel(const RansacProb* __restrict__ probs, ParamDev P, int it0, const int* __restrict__ strag, int strag_cap)
{
    __shared__ double rows_s[VISO_CONT_WARPS][12][7];
    __shared__ double sums_s[VISO_CONT_WARPS][28];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int count = min(strag[0], strag_cap);
    for (int e = blockIdx.x * VISO_CONT_WARPS + w; e < count; e += gridDim.x * VISO_CONT_WARPS) {
        const RansacProb& pb = probs[strag[2 + 2 * e]];
        const int hId = strag[3 + 2 * e];
        const int n = *pb.n;
        int s[3];
        if (pb.table) { s[0] = pb.table[3 * hId]; s[1] = pb.table[3 * hId + 1]; s[2] = pb.table[3 * hId + 2]; }
        else sample_from_seeds(pb.seeds + 3 * hId, n, s);
        const int S = pb.stride;
        /* this lane's sample point: (point, column) for lanes 0..17, the point of the residual column for 18..20 */
        const int pi = lane < 18 ? lane / 6 : (lane < 21 ? lane - 18 : 0);
        const int jc = lane < 18 ? lane - 6 * pi : 6;
        const int a = s[pi]; /* in range: the quad kernel only lists hypotheses with valid samples */
        const double Xp = pb.X[a], Yp = pb.X[S + a], Zp = pb.X[2 * S + a];
        double ob[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) ob[r] = pb.obs[r * S + a];
        const double wgt = weight_of(P, pb.obs[min(pi, n - 1)]); /* columns 0,1,2: the LOOP index, viso.cpp:1449 */
        double tr[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) tr[j] = pb.hyp_tr[6 * hId + j];
        int ok = 0;
        for (int it = it0; it < 100; ++it) {
            const double ang = tr[lane < 3 ? lane : 0];
            const double sv = viso_sc::sin_glibc(ang), cv = viso_sc::cos_glibc(ang);
            Rot R;
            {
                const double sx = __shfl_sync(FULL, sv, 0), cx = __shfl_sync(FULL, cv, 0);
                const double sy = __shfl_sync(FULL, sv, 1), cy = __shfl_sync(FULL, cv, 1);
                const double sz = __shfl_sync(FULL, sv, 2), cz = __shfl_sync(FULL, cv, 2);
                rot_from_sincos(sx, cx, sy, cy, sz, cz, tr, R, true);
            }
            if (lane < 21) {
                const CamPt c = cam_point(R, P, Xp, Yp, Zp);
                double col[4];
                if (lane < 18) jac_column(R, P, c, Xp, Yp, Zp, wgt, jc, col);
                else residual_column(P, c, wgt, ob, col);
#pragma unroll
                for (int r = 0; r < 4; ++r) rows_s[w][4 * pi + r][jc] = col[r];
            }
            __syncwarp();
            if (lane < 27) {
                const int sa = c_pair_a[lane], sb = c_pair_b[lane];
                double acc = 0;
#pragma unroll
                for (int k = 0; k < 12; ++k) acc += rows_s[w][k][sa] * rows_s[w][k][sb];
                sums_s[w][lane] = acc;
            }
            __syncwarp();
            int flag; /* 0 continue, 1 converged, 2 singular: the same on all lanes */
            double p[6] = {0, 0, 0, 0, 0, 0};
#if VISO_CONT_LU_WARP
            if (!lu_solve6_warp(sums_s[w], lane, p)) flag = 2;
            else {
                flag = 1;
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    if (p[j] > P.thresh) flag = 0; /* fabs(p > thresh), viso.cpp:1610 */
            }
#else
            flag = 0;
            if (lane == 0) {
                double A[6][6], b[6];
                int t = 0;
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = r; c < 6; ++c) { A[r][c] = sums_s[w][t]; A[c][r] = sums_s[w][t]; ++t; }
#pragma unroll
                for (int r = 0; r < 6; ++r) b[r] = sums_s[w][21 + r];
                if (!lu_solve6(A, b)) flag = 2;
                else {
                    flag = 1;
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                        if (b[j] > P.thresh) flag = 0; /* fabs(p > thresh), viso.cpp:1610 */
#pragma unroll
                    for (int j = 0; j < 6; ++j) p[j] = b[j];
                }
            }
#endif
#if !VISO_CONT_LU_WARP
            flag = __shfl_sync(FULL, flag, 0);
#pragma unroll
            for (int j = 0; j < 6; ++j) p[j] = __shfl_sync(FULL, p[j], 0);
#endif
            if (flag == 2) { ok = 0; break; }
            if (flag == 1) { ok = 1; break; }
#pragma unroll
            for (int j = 0; j < 6; ++j) tr[j] = tr[j] + p[j];
            __syncwarp(); /* rows_s / sums_s are rewritten by the next iteration */
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 6; ++j) pb.hyp_tr[6 * hId + j] = tr[j];
            pb.hyp_ok[hId] = ok;
        }
    }
}

/* One warp per hypothesis: support-set size, viso.cpp:1563 (get_inliers, :1509-1537) */
#ifndef VISO_SCORE_MINB
#define VISO_SCORE_MINB 4 /* 64 registers, 32 warps per SM: the loop waits on L1-hit latency (long scoreboard), measured -0.08 ms per 1000 frames */
#endif
__global__ void __launch_bounds__(256, VISO_SCORE_MINB) ransac_score_kernel(const RansacProb* __restrict__ probs, ParamDev P)
{
    const RansacProb& pb = probs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int hId = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n = *pb.n;
    if (hId >= pb.H || n < pb.min_n || n < 1) return;
    if (!pb.hyp_ok[hId]) return;
    double tr[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) tr[j] = pb.hyp_tr[6 * hId + j];
    Rot R;
    make_rot(tr, R, false);
    int cnt = 0;
    for (int i = lane; i < n; i += 32) cnt += inlier_point(R, P, pb.X, pb.obs, pb.stride, i) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    if (lane == 0) pb.hyp_count[hId] = cnt;
}

#define VISO_GN_PARALLEL_MIN 1024 /* active sets above this size use the parallel normal-equation sums */
#define VISO_GN_STAGE 512 /* Jacobian rows staged in shared memory per round of the sequential sums (28 KB) */

/*
 * Block-cooperative minimize_reproj (viso.cpp:1583-1623) over an arbitrary active set.  Per iteration:
 * all threads write Jacobian rows + residuals to `scratch` ([4*na][7]); lanes 0..26 of warp 0 then form the 21
 * JtJ sums and 6 Jt*r sums SEQUENTIALLY in row order (bit-identical to cv::mulTransposed / the oracle's
 * J^T r); thread 0 solves and decides.  tr_s: shared double[6], in/out.  Returns 1 converged / 0 failed.
 */
__device__ int gn_block(const double* __restrict__ X, const double* __restrict__ obs, int stride, int n,
                        const int* __restrict__ active, int na, double* tr_s, const ParamDev& P,
                        double* __restrict__ scratch, double* sums_s /* smem[27] */, int* flag_s /* smem */,
                        double* stage_s /* smem[VISO_GN_STAGE * 7] */)
{
    for (int it = 0; it < 100; ++it) {
        double tr[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) tr[j] = tr_s[j];
        Rot R;
        make_rot(tr, R, true);
        if (na > VISO_GN_PARALLEL_MIN) {
            /* Large active sets (BASELINE configs[3]: ~7000 inliers of 10 000): the 27 sums are formed in parallel --
             * every thread accumulates the products of its own points (i = tid, tid + 256, ...), then a fixed
             * shuffle tree per warp and the warps in order.  Deterministic, but not cv::mulTransposed's row order:
             * the refined tr agrees with the sequential path to ~1e-13 relative (north_star's tolerance is 1e-6). */
            double acc[27];
#pragma unroll
            for (int t = 0; t < 27; ++t) acc[t] = 0;
            for (int i = threadIdx.x; i < na; i += blockDim.x) {
                const int a = active[i];
                double ob[4], rows[4][7];
#pragma unroll
                for (int r = 0; r < 4; ++r) ob[r] = obs[r * stride + a];
                const double w = weight_of(P, obs[i]); /* column i, viso.cpp:1449 */
                point_rows(R, P, X[a], X[stride + a], X[2 * stride + a], w, ob, rows);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int t = 0;
#pragma unroll
                    for (int ja = 0; ja < 6; ++ja)
#pragma unroll
                        for (int jc = ja; jc < 6; ++jc) { acc[t] += rows[r][ja] * rows[r][jc]; ++t; }
#pragma unroll
                    for (int ja = 0; ja < 6; ++ja) acc[21 + ja] += rows[r][ja] * rows[r][6];
                }
            }
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
            for (int t = 0; t < 27; ++t) {
                double v = acc[t];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                if (lane == 0) stage_s[warp * 27 + t] = v;
            }
            __syncthreads();
            if (threadIdx.x < 27) {
                double sum = 0;
                for (int w2 = 0; w2 < nw; ++w2) sum += stage_s[w2 * 27 + threadIdx.x];
                sums_s[threadIdx.x] = sum;
            }
        } else {
            for (int i = threadIdx.x; i < na; i += blockDim.x) {
                const int a = active[i];
                double ob[4], rows[4][7];
    #pragma unroll
                for (int r = 0; r < 4; ++r) ob[r] = obs[r * stride + a];
                const double w = weight_of(P, obs[i]); /* column i, viso.cpp:1449 */
                point_rows(R, P, X[a], X[stride + a], X[2 * stride + a], w, ob, rows);
    #pragma unroll
                for (int r = 0; r < 4; ++r)
    #pragma unroll
                    for (int c = 0; c < 7; ++c) scratch[(size_t)(4 * i + r) * 7 + c] = rows[r][c];
            }
            __syncthreads();
            {
                /* the rows come back through shared memory, VISO_GN_STAGE rows at a time (coalesced, all threads), so the
                 * 27 sequential accumulation chains read with shared-memory latency instead of L2 latency; the order of
                 * the additions is unchanged: chunks in order, rows ascending inside a chunk */
                const int a = threadIdx.x < 27 ? c_pair_a[threadIdx.x] : 0, b = threadIdx.x < 27 ? c_pair_b[threadIdx.x] : 0;
                double s = 0;
                const int rowsN = 4 * na;
                for (int k0 = 0; k0 < rowsN; k0 += VISO_GN_STAGE) {
                    const int cnt = min(VISO_GN_STAGE, rowsN - k0);
                    for (int e = threadIdx.x; e < cnt * 7; e += blockDim.x) stage_s[e] = scratch[(size_t)k0 * 7 + e];
                    __syncthreads();
                    if (threadIdx.x < 27) {
    #pragma unroll 8
                        for (int k = 0; k < cnt; ++k) s += stage_s[k * 7 + a] * stage_s[k * 7 + b];
                    }
                    __syncthreads();
                }
                if (threadIdx.x < 27) sums_s[threadIdx.x] = s;
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) { /* warp 0: the LU with one column per lane (lu_solve6_warp), decision by every lane alike */
            double b[6];
            int flag;
            if (!lu_solve6_warp(sums_s, threadIdx.x, b)) flag = 2;
            else {
                bool conv = true;
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    if (b[j] > P.thresh) { conv = false; break; }
                flag = conv ? 1 : 0;
            }
            if (threadIdx.x == 0) {
                if (flag == 0) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) tr_s[j] = tr_s[j] + b[j];
                }
                *flag_s = flag;
            }
        }
        __syncthreads();
        const int flag = *flag_s;
        __syncthreads();
        if (flag == 1) return 1;
        if (flag == 2) return 0;
    }
    return 0;
}

/* ordered inlier list of `tr` over all n points (block cooperative); returns the count (same in all threads) */
__device__ int inliers_block(const double* __restrict__ X, const double* __restrict__ obs, int stride, int n,
                             const double* tr_s, const ParamDev& P, int* __restrict__ out, int* warp_tot)
{
    double tr[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) tr[j] = tr_s[j];
    Rot R;
    make_rot(tr, R, false);
    int base = 0;
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const bool flag = i < n && inlier_point(R, P, X, obs, stride, i);
        const int slot = block_compact_slot(flag, base, warp_tot);
        if (flag) out[slot] = i;
    }
    return base;
}

/* One CTA per problem: viso.cpp:1564-1579 */
#ifndef VISO_FINAL_MINB
#define VISO_FINAL_MINB 4 /* 64 registers (the spills of the row evaluation are hidden: the CTA waits on barriers and on the sequential sums); 1 -> 2 -> 4 CTAs per SM: 0.43 -> 0.28 -> 0.22 ms, 5 and more lose again */
#endif
/* MINB = resident CTAs per SM the instance is compiled for: VISO_FINAL_MINB (64 registers) for the pipeline's problems
 * (max_n = keypoints per image, a few hundred correspondences, sequential sums), 1 (all the registers it wants) for
 * problems of more than 4096 correspondences, whose parallel
 * normal-equation sums keep 27 accumulators per thread (BASELINE configs[3]: 1.29 instead of 1.45 ms per call) */
template <int MINB>
__global__ void __launch_bounds__(256, MINB) ransac_final_kernel(const RansacProb* __restrict__ probs, ParamDev P)
{
    __shared__ int warp_tot[32];
    __shared__ int best_cnt_s[256], best_idx_s[256];
    __shared__ double tr_s[6];
    __shared__ double sums_s[27];
    __shared__ double stage_s[VISO_GN_STAGE * 7];
    __shared__ int flag_s;
    const RansacProb& pb = probs[blockIdx.x];
    const int n = *pb.n;
    viso_record_dev* rec = pb.rec;
    if (n < pb.min_n || n < 1) {
        if (threadIdx.x == 0) {
            for (int j = 0; j < 6; ++j) rec->tr[j] = pb.tr_init[j];
            rec->ok = 0; rec->n_inliers = 0; rec->best_hyp = -1; rec->n_circ = n;
        }
        return;
    }
    /* first hypothesis with the strictly largest support (viso.cpp:1564: '>' against an initially empty set) */
    int bc = 0, bi = INT_MAX;
    for (int hId = threadIdx.x; hId < pb.H; hId += blockDim.x) {
        if (pb.hyp_ok[hId]) {
            const int c = pb.hyp_count[hId];
            if (c > bc) { bc = c; bi = hId; }
        }
    }
    best_cnt_s[threadIdx.x] = bc; best_idx_s[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const int c2 = best_cnt_s[threadIdx.x + o], i2 = best_idx_s[threadIdx.x + o];
            if (c2 > best_cnt_s[threadIdx.x] || (c2 == best_cnt_s[threadIdx.x] && i2 < best_idx_s[threadIdx.x])) {
                best_cnt_s[threadIdx.x] = c2; best_idx_s[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    const int best_cnt = best_cnt_s[0];
    const int best_hyp = best_cnt > 0 ? best_idx_s[0] : -1;
    if (threadIdx.x < 6) tr_s[threadIdx.x] = best_hyp >= 0 ? pb.hyp_tr[6 * best_hyp + threadIdx.x] : pb.tr_init[threadIdx.x];
    __syncthreads();
    int n_act = 0;
    if (best_hyp >= 0) n_act = inliers_block(pb.X, pb.obs, pb.stride, n, tr_s, P, pb.active, warp_tot);
    __syncthreads();
    int ok = 0, n_inl = n_act;
    const int* list = pb.active;
    if (n_act >= 6) {
        ok = gn_block(pb.X, pb.obs, pb.stride, n, pb.active, n_act, tr_s, P, pb.scratch, sums_s, &flag_s, stage_s);
        if (ok) {
            n_inl = inliers_block(pb.X, pb.obs, pb.stride, n, tr_s, P, pb.inliers, warp_tot);
            list = pb.inliers;
        }
    }
    __syncthreads();
    if (list != pb.inliers) /* failure: the reference leaves best_inliers = RANSAC support set */
        for (int i = threadIdx.x; i < n_act; i += blockDim.x) pb.inliers[i] = pb.active[i];
    if (threadIdx.x == 0) {
        for (int j = 0; j < 6; ++j) rec->tr[j] = tr_s[j];
        rec->ok = ok; rec->n_inliers = n_inl; rec->best_hyp = best_hyp; rec->n_circ = n;
    }
}

/* standalone minimize_reproj (viso_minimize_reproj): one CTA */
__global__ void __launch_bounds__(256) gn_kernel(const double* X, const double* obs, int stride, const int* active,
                                                 int na, double* tr, int* ok, double* scratch, ParamDev P)
{
    __shared__ double tr_s[6];
    __shared__ double sums_s[27];
    __shared__ double stage_s[VISO_GN_STAGE * 7];
    __shared__ int flag_s;
    if (threadIdx.x < 6) tr_s[threadIdx.x] = tr[threadIdx.x];
    __syncthreads();
    const int r = gn_block(X, obs, stride, stride, active, na, tr_s, P, scratch, sums_s, &flag_s, stage_s);
    __syncthreads();
    if (threadIdx.x < 6) tr[threadIdx.x] = tr_s[threadIdx.x];
    if (threadIdx.x == 0) *ok = r;
}

/* standalone get_inliers (viso_get_inliers): one CTA */
__global__ void __launch_bounds__(256) inliers_kernel(const double* X, const double* obs, int n, int stride,
                                                      const double* tr, int* inliers, int* count, ParamDev P)
{
    __shared__ int warp_tot[32];
    __shared__ double tr_s[6];
    if (threadIdx.x < 6) tr_s[threadIdx.x] = tr[threadIdx.x];
    __syncthreads();
    const int c = inliers_block(X, obs, stride, n, tr_s, P, inliers, warp_tot);
    if (threadIdx.x == 0) *count = c;
}

/* ------------------------------------------------------------------------------------------------ launchers */

cudaError_t viso_launch_ransac(const RansacProb* probs, int n_probs, int max_H, int max_n, ParamDev p, int* strag,
                               int hyp_it_cap, int sm_count, cudaStream_t s, int* launches)
{
    if (n_probs <= 0 || max_H <= 0) return cudaSuccess;
    const int it_cap = strag ? std::max(1, std::min(100, hyp_it_cap)) : 100;
    const long long strag_cap = (long long)n_probs * max_H;
    cudaError_t e;
    if (it_cap < 100) {
        e = viso_launch_zero(strag, 1, s);
        if (e != cudaSuccess) return e;
        if (launches) *launches += 1;
    }
    ransac_hyp_kernel<<<dim3((max_H + VISO_HYP_PER_CTA - 1) / VISO_HYP_PER_CTA, n_probs), VISO_HYP_PER_CTA * 4, 0, s>>>(
        probs, p, it_cap, strag, (int)strag_cap);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (it_cap < 100) {
        /* one warp per unfinished hypothesis where they fit (their number is only known on the device: a grid-stride
         * loop over the list takes whatever is there) */
        const long long want = (strag_cap + VISO_CONT_WARPS - 1) / VISO_CONT_WARPS;
        const int grid = (int)std::max(1LL, std::min(want, (long long)std::max(1, sm_count) * (16 / VISO_CONT_WARPS)));
        ransac_hyp_cont_kernel<<<grid, VISO_CONT_WARPS * 32, 0, s>>>(probs, p, it_cap, strag, (int)strag_cap);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (launches) *launches += 1;
    }
    ransac_score_kernel<<<dim3((max_H + 7) / 8, n_probs), 256, 0, s>>>(probs, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (max_n > 4 * VISO_GN_PARALLEL_MIN) ransac_final_kernel<1><<<n_probs, 256, 0, s>>>(probs, p);
    else ransac_final_kernel<VISO_FINAL_MINB><<<n_probs, 256, 0, s>>>(probs, p);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

cudaError_t viso_launch_gn(const double* X, const double* obs, int stride, const int* active, int na, double* tr,
                           int* ok, double* scratch, ParamDev p, cudaStream_t s)
{
    gn_kernel<<<1, 256, 0, s>>>(X, obs, stride, active, na, tr, ok, scratch, p);
    return cudaGetLastError();
}

cudaError_t viso_launch_inliers(const double* X, const double* obs, int n, int stride, const double* tr, int* inliers,
                                int* count, ParamDev p, cudaStream_t s)
{
    inliers_kernel<<<1, 256, 0, s>>>(X, obs, n, stride, tr, inliers, count, p);
    return cudaGetLastError();
}

/* test hook: the device sin / cos of glibc_sincos.h over an array (tests/test_gpu_parity.py compares with the host libm) */
__global__ void sincos_probe_kernel(const double* __restrict__ x, int n, double* __restrict__ s, double* __restrict__ c)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        s[i] = viso_sc::sin_glibc(x[i]);
        c[i] = viso_sc::cos_glibc(x[i]);
    }
}

cudaError_t viso_launch_sincos_probe(const double* x, int n, double* s, double* c, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    sincos_probe_kernel<<<std::min((n + 255) / 256, 1024), 256, 0, st>>>(x, n, s, c); /* grid-stride test hook */
    return cudaGetLastError();
}
