/*
 * detect.cu -- HarrisBinnedFeatureDetector::detectImpl on the device (reference src/viso.cpp:911-979).
 *
 *   harris_bin_kernel      one CTA per (image, bin): cv::cornerHarris(block 3, aperture 5, k) evaluated for the bin's
 *                          pixels only, never materialised in HBM, then the bin's n/(nbinx*nbiny) strongest |response|
 *   harris_compact_kernel  one CTA per image: concatenates the bins in the reference's order (binx outer, biny inner)
 *
 * Arithmetic: cornerHarris is float32 and OpenCV does not define its rounding; include/viso_b200.h (viso_detect_harris) fixes one
 * canonical operation order and this kernel follows it operation by operation -- the translation unit is compiled with
 * -fmad=false -- so responses, kept sets and keypoint order are reproducible bit for bit.
 * Order rule: per bin the kept elements are the largest by (|response|, x, y), emitted ascending.
 *
 * Mapping: a warp walks down the rows of a 64-pixel-wide strip, two adjacent columns per lane; the 5-row Sobel window
 * and the 3-row box window live in registers, horizontal neighbours come from the adjacent lane (5 shuffles per
 * column and row).  Pixels are read straight from global memory (each row of a strip is two sectors).  58 of the 64
 * columns produce responses (3 halo columns each side), a KITTI bin (51 x 75) is one strip, split in two row segments.
 */
#include "viso_dev.h"
#include "common.cuh"

#ifndef HARRIS_WARPS
#define HARRIS_WARPS 2
#endif
#define HARRIS_MAXIMA 256        /* column-segment maxima used for the lower bound of the cut value */
#define HARRIS_STRIP 58          /* response columns per warp strip */
#define HARRIS_DIRECT 256        /* candidate lists longer than this are first cut by bisection on the value */
#ifndef HARRIS_AHEAD
#define HARRIS_AHEAD 3           /* pixel rows requested ahead of the row being processed */
#endif
#ifndef HARRIS_UNROLL
#define HARRIS_UNROLL 5          /* = the Sobel window height: the window shift becomes register renaming */
#endif
#define HARRIS_PRAGMA_(x) _Pragma(#x)
#define HARRIS_PRAGMA(x) HARRIS_PRAGMA_(x)

namespace {

/* column stride of the response array: the rows padded to whole warp segments, odd so that the two-column lanes of a
 * warp spread over 16 banks (an even stride would put them on 4 or 8) */
__host__ __device__ __forceinline__ int harris_col_stride(int sy)
{
    return (((sy + HARRIS_WARPS - 1) / HARRIS_WARPS) * HARRIS_WARPS) | 1;
}

__device__ __forceinline__ int reflect_clamp(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return min(max(i, 0), n - 1);
}

struct Tri { float a, b, c; };   /* xx, xy, yy (or their sums) */

/* register state of one warp strip: 5 rows of the two Sobel row passes and 2 rows of box row sums, for two columns */
struct WalkState {
    float r0[5], r1[5], t0[5], t1[5];
    Tri prev0, prev1, cur0, cur1;
    float p0[HARRIS_AHEAD], p1[HARRIS_AHEAD];   /* pixels of the current row and the rows after it (requested early) */
    const unsigned char* pf;                    /* fast rows: the pixel (row to request next, first column) */
    unsigned* vp;                               /* fast rows: where the next response of the first column goes */
};

struct WalkLane {
    int lx0, lx1;                /* reflected image columns of the lane's two pixels */
    bool left0, left1, right0, right1, valid0, valid1;
};

/* One pixel row enters the windows.  MODE 0: Sobel row pass only; 1: + covariance row yi-2; 2: + response row yi-3.
 * GENERAL rows handle everything (reflected rows and columns, rows past the segment); the others are rows whose
 * 64 x (5 + HARRIS_AHEAD) pixel neighbourhood is inside the image and whose response row is inside the bin (see
 * harris_bin_kernel): running pointers, no selects, no masks. */
template <int MODE, bool GENERAL>
__device__ __forceinline__ void walk_row(WalkState& w, const WalkLane& ln, const unsigned char* __restrict__ img,
                                         const HarrisCfg& c, int yi, int y0, int rb, int sp, unsigned* vcol, unsigned& best)
{
    /* a row further down is requested before this row's arithmetic */
    float q0, q1;
    if (GENERAL) {
        const unsigned char* nrow = img + (size_t)reflect_clamp(yi + HARRIS_AHEAD, c.h) * c.pitch;
        q0 = __ldg(nrow + ln.lx0); q1 = __ldg(nrow + ln.lx1);
    } else {
        q0 = __ldg(w.pf); q1 = __ldg(w.pf + 1);
        w.pf += c.pitch;
    }
    const float p0 = w.p0[0], p1 = w.p1[0];
    const float L0 = __shfl_up_sync(FULL, p0, 1), L1 = __shfl_up_sync(FULL, p1, 1);
    const float R0 = __shfl_down_sync(FULL, p0, 1), R1 = __shfl_down_sync(FULL, p1, 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) { w.r0[i] = w.r0[i + 1]; w.r1[i] = w.r1[i + 1]; w.t0[i] = w.t0[i + 1]; w.t1[i] = w.t1[i + 1]; }
    w.r0[4] = (R0 - L0) + 2.0f * (p1 - L1);
    w.r1[4] = (R1 - L1) + 2.0f * (R0 - p0);
    float t = c.f0 * p0; t = t + c.f1 * (L1 + p1); t = t + c.f2 * (L0 + R0); w.t0[4] = t;
    t = c.f0 * p1; t = t + c.f1 * (p0 + R0); t = t + c.f2 * (L1 + R1); w.t1[4] = t;
#pragma unroll
    for (int i = 0; i + 1 < HARRIS_AHEAD; ++i) { w.p0[i] = w.p0[i + 1]; w.p1[i] = w.p1[i + 1]; }
    w.p0[HARRIS_AHEAD - 1] = q0; w.p1[HARRIS_AHEAD - 1] = q1;
    if (MODE == 0) return;
    float dx0 = c.f0 * w.r0[2]; dx0 = dx0 + c.f1 * (w.r0[1] + w.r0[3]); dx0 = dx0 + c.f2 * (w.r0[0] + w.r0[4]);
    float dx1 = c.f0 * w.r1[2]; dx1 = dx1 + c.f1 * (w.r1[1] + w.r1[3]); dx1 = dx1 + c.f2 * (w.r1[0] + w.r1[4]);
    float dy0 = 2.0f * (w.t0[3] - w.t0[1]); dy0 = dy0 + (w.t0[4] - w.t0[0]);
    float dy1 = 2.0f * (w.t1[3] - w.t1[1]); dy1 = dy1 + (w.t1[4] - w.t1[0]);
    const Tri m0{dx0 * dx0, dx0 * dy0, dy0 * dy0}, m1{dx1 * dx1, dx1 * dy1, dy1 * dy1};
    Tri l{__shfl_up_sync(FULL, m1.a, 1), __shfl_up_sync(FULL, m1.b, 1), __shfl_up_sync(FULL, m1.c, 1)};
    Tri r{__shfl_down_sync(FULL, m0.a, 1), __shfl_down_sync(FULL, m0.b, 1), __shfl_down_sync(FULL, m0.c, 1)};
    Tri lm1 = m0, rm0 = m1;
    if (GENERAL) {
        if (ln.left0) l = m1;        /* covariance column -1 is column 1 (BORDER_REFLECT_101 of cv::boxFilter) */
        if (ln.right1) r = m0;       /* column w is column w-2 */
        if (ln.left1) lm1 = r;
        if (ln.right0) rm0 = l;
    }
    Tri new0{(l.a + m0.a) + rm0.a, (l.b + m0.b) + rm0.b, (l.c + m0.c) + rm0.c};
    Tri new1{(lm1.a + m1.a) + r.a, (lm1.b + m1.b) + r.b, (lm1.c + m1.c) + r.c};
    const int yc = yi - 2;
    if (GENERAL && yc == c.h) { new0 = w.prev0; new1 = w.prev1; }   /* covariance row h is row h-2 */
    if (MODE == 2) {
        const int yr = yc - 1;
        if (GENERAL && yr == 0) { w.prev0 = new0; w.prev1 = new1; }   /* row -1 is row 1 */
        const float a0 = (w.prev0.a + w.cur0.a) + new0.a, b0 = (w.prev0.b + w.cur0.b) + new0.b, c0 = (w.prev0.c + w.cur0.c) + new0.c;
        const float a1 = (w.prev1.a + w.cur1.a) + new1.a, b1 = (w.prev1.b + w.cur1.b) + new1.b, c1 = (w.prev1.c + w.cur1.c) + new1.c;
        const float tr0 = a0 + c0, tr1 = a1 + c1;
        const float h0 = (a0 * c0 - b0 * b0) - (c.k * tr0) * tr0;
        const float h1 = (a1 * c1 - b1 * b1) - (c.k * tr1) * tr1;
        const unsigned v0 = __float_as_uint(h0) & 0x7fffffffu, v1 = __float_as_uint(h1) & 0x7fffffffu;
        if (GENERAL) {
            if (ln.valid0 && yr < rb) { vcol[yr - y0] = v0; best = max(best, v0); }
            if (ln.valid1 && yr < rb) { vcol[sp + yr - y0] = v1; best = max(best, v1); }
        } else {
            if (ln.valid0) { w.vp[0] = v0; best = max(best, v0); }
            if (ln.valid1) { w.vp[sp] = v1; best = max(best, v1); }
            ++w.vp;
        }
    }
    w.prev0 = w.cur0; w.prev1 = w.cur1; w.cur0 = new0; w.cur1 = new1;
}

/* responses of rows [ra, min(ra + rows, rb)) x strip columns, written (as the bit pattern of |response|) into vals in
 * scan order (column stride sp >= rows of the bin); returns the largest value this lane produced.  Every trip count
 * (rows, nfast) is the same for all warps of the CTA: the compiler can then prove the shuffles convergent. */
__device__ __forceinline__ unsigned harris_walk(const unsigned char* __restrict__ img, const HarrisCfg& c, int x0, int y0,
                                                int xs, int ncol, int ra, int rb, int rows, int nfast, int sp,
                                                unsigned* vals, int lane)
{
    const int cbase = x0 + xs - 3;
    const int q0 = 2 * lane, cx0 = cbase + q0, cx1 = cx0 + 1;
    WalkLane ln;
    ln.lx0 = reflect_clamp(cx0, c.w); ln.lx1 = reflect_clamp(cx1, c.w);
    ln.left0 = cx0 == 0; ln.left1 = cx1 == 0; ln.right0 = cx0 == c.w - 1; ln.right1 = cx1 == c.w - 1;
    ln.valid0 = q0 >= 3 && q0 < 3 + ncol; ln.valid1 = q0 + 1 >= 3 && q0 + 1 < 3 + ncol;
    unsigned* vcol = vals + (xs + q0 - 3) * sp;        /* the reference's scan order: x outer, y inner */
    WalkState w;
#pragma unroll
    for (int i = 0; i < 5; ++i) { w.r0[i] = w.r1[i] = w.t0[i] = w.t1[i] = 0.f; }
    w.prev0 = w.prev1 = w.cur0 = w.cur1 = Tri{0.f, 0.f, 0.f};
    unsigned best = 0;
    /* pixel rows ra-3 .. : four rows fill the Sobel window, covariance rows ra-1 and ra fill the box window (for
     * ra == 0 "row -1" is computed from reflected pixels and then replaced by row 1, see walk_row) */
    int yi = ra - 3;
#pragma unroll
    for (int i = 0; i < HARRIS_AHEAD; ++i) {
        const unsigned char* row = img + (size_t)reflect_clamp(yi + i, c.h) * c.pitch;
        w.p0[i] = __ldg(row + ln.lx0); w.p1[i] = __ldg(row + ln.lx1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i, ++yi) walk_row<0, true>(w, ln, img, c, yi, y0, rb, sp, vcol, best);
#pragma unroll
    for (int i = 0; i < 2; ++i, ++yi) walk_row<1, true>(w, ln, img, c, yi, y0, rb, sp, vcol, best);
    /* response rows: the first one may be image row 0 (general), then nfast rows without any special case */
    walk_row<2, true>(w, ln, img, c, yi, y0, rb, sp, vcol, best);
    ++yi;
    w.pf = img + (size_t)max(yi + HARRIS_AHEAD, 0) * c.pitch + ln.lx0;
    w.vp = vcol + (yi - 3 - y0);
HARRIS_PRAGMA(unroll HARRIS_UNROLL)
    for (int i = 0; i < nfast; ++i, ++yi) walk_row<2, false>(w, ln, img, c, yi, y0, rb, sp, vcol, best);
#pragma unroll 1
    for (int i = 1 + nfast; i < rows; ++i, ++yi) walk_row<2, true>(w, ln, img, c, yi, y0, rb, sp, vcol, best);
    return best;
}

#ifndef HARRIS_MINB
#define HARRIS_MINB 8 /* 126 registers, 16 warps per SM: 4.4 -> 4.1 ms per 2000 images (96 registers without the bound, 135 at 4 - 6: both slower) */
#endif
__global__ void __launch_bounds__(HARRIS_WARPS * 32, HARRIS_MINB)
harris_bin_kernel(const DetectJob* __restrict__ jobs, HarrisCfg c)
{
    extern __shared__ unsigned smem_u[];
    const DetectJob job = jobs[blockIdx.y];
    if (*job.detect == 0) return;
    const int sp = harris_col_stride(c.sy);
    const int npx = c.sx * sp;
    unsigned* vals = smem_u;
    const int npx4 = (npx + 3) & ~3;
    unsigned* mx = vals + npx4;
    unsigned short* cand = reinterpret_cast<unsigned short*>(mx + HARRIS_MAXIMA);
    __shared__ int s_ncand, s_count;
    __shared__ unsigned s_cut, s_max;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nt = HARRIS_WARPS * 32, nw = HARRIS_WARPS;
    const int bin = blockIdx.x, binx = bin / c.nbiny, biny = bin % c.nbiny;
    const int x0 = binx * c.sx, y0 = biny * c.sy;

    for (int i = tid; i < HARRIS_MAXIMA; i += nt) mx[i] = 0;
    if (tid < npx4 - npx) vals[npx + tid] = 0;
    /* the pad rows of the columns (sp - sy of them) are never written */
    for (int i = tid; i < c.sx * (sp - c.sy); i += nt) vals[(i / (sp - c.sy)) * sp + c.sy + i % (sp - c.sy)] = 0;
    if (tid == 0) { s_ncand = 0; s_cut = 0; s_max = 0; }
    __syncthreads();
    /* strip by strip; the warps split a strip's rows evenly */
    const int nstrip = (c.sx + HARRIS_STRIP - 1) / HARRIS_STRIP, rows_per = (c.sy + nw - 1) / nw, ntask = nstrip * nw;
    for (int strip = 0; strip < nstrip; ++strip) {
        const int xs = strip * HARRIS_STRIP, ncol = min(HARRIS_STRIP, c.sx - xs);
        const int ra = y0 + warp * rows_per, rb = y0 + c.sy;
        /* rows 1 .. nfast of every warp's segment need no special case: the strip's 64 columns are inside the image,
         * the row requested ahead (ra + 3 + i + HARRIS_AHEAD, largest for the last warp) exists, and the response row
         * is inside the bin (only the last nw - 1 rows of the last segment can fall into the column pad) */
        const bool cols_inside = x0 + xs - 3 >= 0 && x0 + xs - 3 + 63 <= c.w - 1;
        const int ra_last = y0 + (nw - 1) * rows_per;
        const int nfast = cols_inside ? max(0, min(c.h - 4 - HARRIS_AHEAD - ra_last, rows_per - nw)) : 0;
        const unsigned best = harris_walk(job.img, c, x0, y0, xs, ncol, ra, rb, rows_per, nfast, sp, vals, lane);
        const int task = strip * nw + warp;
        if (task * 32 + lane < HARRIS_MAXIMA) mx[task * 32 + lane] = best;
    }
    __syncthreads();

    /* lower bound of the cut: the per-th largest of the column-segment maxima (each is attained by some pixel, so at
     * least `per` pixels reach it); 0 when fewer maxima than `per` exist */
    const int nmax = min(ntask * 32, HARRIS_MAXIMA);
    if (c.per <= nmax)
        for (int i = tid; i < nmax; i += nt) {
            const unsigned v = mx[i];
            int rank = 0;
            for (int j = 0; j < nmax; ++j) {
                const unsigned o = mx[j];
                rank += (o > v) || (o == v && j < i);
            }
            if (rank == c.per - 1) s_cut = v;
        }
    __syncthreads();
    unsigned cut = max(s_cut, 1u);      /* responses equal to zero are never keypoints (viso.cpp:956) */
    {
        const uint4* v4 = reinterpret_cast<const uint4*>(vals);
        for (int i = tid; i < (npx + 3) / 4; i += nt) {   /* the pad words are 0 */
            const uint4 v = v4[i];
            if (max(max(v.x, v.y), max(v.z, v.w)) < cut) continue;
            if (v.x >= cut) cand[atomicAdd(&s_ncand, 1)] = (unsigned short)(4 * i);
            if (v.y >= cut) cand[atomicAdd(&s_ncand, 1)] = (unsigned short)(4 * i + 1);
            if (v.z >= cut) cand[atomicAdd(&s_ncand, 1)] = (unsigned short)(4 * i + 2);
            if (v.w >= cut) cand[atomicAdd(&s_ncand, 1)] = (unsigned short)(4 * i + 3);
        }
    }
    __syncthreads();
    const int ncand = s_ncand;

    if (ncand > HARRIS_DIRECT && ncand > c.per) {
        /* exact per-th largest value by bisection: count(v >= lo) >= per > count(v >= hi) */
        unsigned vmax = 0;
        for (int i = tid; i < ncand; i += nt) vmax = max(vmax, vals[cand[i]]);
        vmax = __reduce_max_sync(FULL, vmax);
        if (lane == 0) atomicMax(&s_max, vmax);
        __syncthreads();
        unsigned lo = cut, hi = s_max + 1u;
        while (hi - lo > 1u) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            __syncthreads();
            if (tid == 0) s_count = 0;
            __syncthreads();
            int n = 0;
            for (int i = tid; i < ncand; i += nt) n += vals[cand[i]] >= mid;
            n = (int)warp_sum_u((unsigned)n);
            if (lane == 0 && n) atomicAdd(&s_count, n);
            __syncthreads();
            if (s_count >= c.per) lo = mid; else hi = mid;
        }
        cut = lo;
    }

    /* exact order among the candidates that reach the cut: key = (value, position in the reference's scan order) */
    const int cnt = min(c.per, ncand);
    float2* out = job.tmp + (size_t)bin * c.per;
    for (int i = tid; i < ncand; i += nt) {
        const unsigned pi = cand[i], vi = vals[pi];
        if (vi < cut) continue;
        int rank = 0;
        for (int j = 0; j < ncand; ++j) {
            const unsigned pj = cand[j], vj = vals[pj];
            rank += (vj > vi) || (vj == vi && pj > pi);
        }
        if (rank < c.per) {
            out[cnt - 1 - rank] = make_float2((float)(x0 + (int)pi / sp), (float)(y0 + (int)pi % sp));
            if (job.resp_tmp) job.resp_tmp[(size_t)bin * c.per + cnt - 1 - rank] = __uint_as_float(vi);
        }
    }
    if (tid == 0) job.bin_count[bin] = cnt;
}

__global__ void __launch_bounds__(128) harris_compact_kernel(const DetectJob* __restrict__ jobs, int nbins, int per)
{
    extern __shared__ int s_off[];   /* nbins + 1 */
    const DetectJob job = jobs[blockIdx.x];
    if (*job.detect == 0) return;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nbins; ++b) { s_off[b] = run; run += job.bin_count[b]; }
        s_off[nbins] = run;
        *job.n = run;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins * per; i += blockDim.x) {
        const int b = i / per, j = i - b * per;
        if (j < s_off[b + 1] - s_off[b]) {
            job.kp[s_off[b] + j] = job.tmp[i];
            if (job.resp) job.resp[s_off[b] + j] = job.resp_tmp[i];
        }
    }
}

} // namespace

size_t viso_harris_cells(const HarrisCfg& c)
{
    return (size_t)c.sx * harris_col_stride(c.sy);
}

size_t viso_harris_smem(const HarrisCfg& c)
{
    const size_t npx = viso_harris_cells(c);
    return ((npx + 3) & ~(size_t)3) * 4 + HARRIS_MAXIMA * 4 + ((npx * 2 + 3) & ~(size_t)3);
}

cudaError_t viso_launch_detect(const DetectJob* jobs, int n_jobs, const HarrisCfg& c, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    const size_t smem = viso_harris_smem(c);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;   /* checked with a message by the callers */
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(harris_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int nbins = c.nbinx * c.nbiny;
    harris_bin_kernel<<<dim3(nbins, n_jobs), HARRIS_WARPS * 32, smem, s>>>(jobs, c);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    harris_compact_kernel<<<n_jobs, 128, (nbins + 1) * sizeof(int), s>>>(jobs, nbins, c.per);
    return cudaGetLastError();
}
