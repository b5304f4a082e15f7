/*
 * kernels.cu -- hand-written sm_100a kernels for the libviso hot path.
 *
 * Compile with -fmad=false: the FP64 estimation code must evaluate exactly the reference's expressions
 * (separate multiply and add, as a stock x86-64 build of the reference does) so that Jacobians, normal
 * equations, LU pivots and inlier decisions agree bit for bit with the CPU path.  Device sin/cos are the only
 * operations that can differ from glibc in the last place.
 *
 * Kernel inventory (DESIGN.md has the roofline for each):
 *   pack_desc_kernel      f32 cv::Mat descriptors -> biased u16 rows (+ row sum), domain check       [HBM]
 *   grid_build_kernel     counting sort of a keypoint set into 16-px cells                            [latency]
 *   sad_match_kernel      match_desc (viso.cpp:668-722): radius search + top-K + Sampson + SAD argmin  [L1/ALU/HBM]
 *   compact_sort_kernel   Match compaction, libstdc++ introsort order (viso.cpp:724), collect_matches
 *                         (viso.cpp:501-514) + triangulate_rectified<double> (viso.cpp:1137-1154)      [latency]
 *   circle_kernel         match_circle (viso.cpp:206-243) + circular gather (viso.cpp:1291-1305)       [latency]
 *   ransac_hyp_kernel     3-point Gauss-Newton per hypothesis (viso.cpp:1555-1562, 1583-1623)          [FP64]
 *   ransac_score_kernel   get_inliers count per hypothesis (viso.cpp:1509-1537)                        [FP64]
 *   ransac_final_kernel   first-best selection, refine on the support set, final support (viso.cpp:1564-1576)
 *   gn_kernel / inliers_kernel / triangulate / project / circle tables: standalone entry points
 */
#include "viso_dev.h"
#include "introsort.h"

#include <math_constants.h>
#include <stdlib.h>
#include <algorithm>

#define FULL 0xffffffffu

/* ------------------------------------------------------------------------------------------------ helpers */

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

/* ------------------------------------------------------------------------------------------------ pack */

/*
 * f32 cv::Mat descriptor rows (viso.cpp:999-1002) -> biased u16 rows (v + 1024, pad elements 0), plus the u32 sum
 * of the packed row.  One warp per row, lane l owns elements 4l..4l+3.  Also the domain check (integer valued,
 * |v| <= 1023).
 */
__global__ void __launch_bounds__(256) pack_desc_kernel(const PackJob* __restrict__ jobs, int dlen, int* err)
{
    const PackJob job = jobs[blockIdx.y];
    if (job.from_image && *job.from_image) return;
    const int n = *job.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + warp; row < n; row += gridDim.x * 8) {
        const float* src = job.d + (size_t)row * dlen;
        unsigned u[4];
        unsigned sum = 0;
        int bad = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int k = lane * 4 + e;
            unsigned v = 0;
            if (k < dlen) {
                float f = __ldg(src + k);
                float r = truncf(f);
                if (!(f == r) || !(fabsf(f) <= 1023.f)) bad = 1;
                else v = (unsigned)((int)r + 1024);
            }
            u[e] = v;
            sum += v;
        }
        uint2 w;
        w.x = u[0] | (u[1] << 16);
        w.y = u[2] | (u[3] << 16);
        reinterpret_cast<uint2*>(job.out + (size_t)row * VISO_DESC_U16)[lane] = w;
        const unsigned tot = warp_sum_u(sum);
        if (lane == 0) job.rsum[row] = tot;
        if (bad) atomicOr(err, 1);
    }
}

/* ------------------------------------------------------------------------------------------------ extract */

/*
 * MyFeatureExtractor::computeImpl, viso.cpp:1004-1024, fused with the packing above: for keypoint k with
 * p = Point2i(kp.pt) (cv::saturate_cast = round half to even, :1013), element (i, j), i, j in [-R, R] row-major, is
 *     (p.y+i > 0 && p.y+i < rows && p.x+j > 0 && p.x+j < cols) ? sobel_x(p.y+i, p.x+j) : 0        (:1018)
 * with sobel_x = cv::Sobel(image, CV_32F, 1, 0, 3, 1, 0, BORDER_REFLECT_101) (:1010), i.e. the integer
 *     (I(y-1,x+1) + 2 I(y,x+1) + I(y+1,x+1)) - (I(y-1,x-1) + 2 I(y,x-1) + I(y+1,x-1)),  index -1 -> 1, n -> n-2.
 * Values are integers in [-1020, 1020]; the row is written biased (+1024) in the u16 layout with its sum.
 * One warp per keypoint.  The (2R+3)^2 source window is first staged in shared memory with the reflection already
 * applied (6 byte loads per lane, ~3 image rows per load instruction instead of 11), then lane l computes elements
 * 4l..4l+3 from the staged window.
 */
#define VISO_EXTRACT_WIN 16 /* staged window pitch; supports radius <= 6 (2R+3 <= 15) */

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1); /* far outside: any in-range pixel, the sample is masked to 0 anyway */
}

#define VISO_EXTRACT_KPW 4 /* keypoints per warp iteration: their image loads are all in flight together */

template <int radius>
__global__ void __launch_bounds__(256) extract_desc_kernel(const ExtractJob* __restrict__ jobs, int width, int height, int pitch)
{
    __shared__ unsigned char win_s[8][VISO_EXTRACT_KPW][VISO_EXTRACT_WIN * VISO_EXTRACT_WIN];
    const ExtractJob job = jobs[blockIdx.y];
    if (!*job.from_image) return;
    const int n = *job.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int side = 2 * radius + 1, dlen = side * side, wside = side + 2, wn = wside * wside;
    constexpr int NL = (wn + 31) / 32;
    for (int row0 = (blockIdx.x * 8 + warp) * VISO_EXTRACT_KPW; row0 < n; row0 += gridDim.x * 8 * VISO_EXTRACT_KPW) {
        int px[VISO_EXTRACT_KPW], py[VISO_EXTRACT_KPW];
        unsigned char pix[VISO_EXTRACT_KPW][NL];
        /* the image is touched once per frame, so these loads mostly miss to DRAM: issue all of them first */
#pragma unroll
        for (int q = 0; q < VISO_EXTRACT_KPW; ++q) {
            const float2 kp = job.kp[min(row0 + q, n - 1)];
            px[q] = __float2int_rn(kp.x); py[q] = __float2int_rn(kp.y);
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                const int idx = min(lane + 32 * j, wn - 1);
                const int wy = idx / wside, wx = idx - wy * wside;
                const int iy = reflect101(py[q] - radius - 1 + wy, height), ix = reflect101(px[q] - radius - 1 + wx, width);
                pix[q][j] = __ldg(job.img + (size_t)iy * pitch + ix);
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < VISO_EXTRACT_KPW; ++q)
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                const int idx = lane + 32 * j;
                if (idx < wn) {
                    const int wy = idx / wside, wx = idx - wy * wside;
                    win_s[warp][q][wy * VISO_EXTRACT_WIN + wx] = pix[q][j];
                }
            }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < VISO_EXTRACT_KPW; ++q) {
            const int row = row0 + q;
            if (row >= n) break; /* warp uniform */
            const unsigned char* win = win_s[warp][q];
            unsigned u[4];
            unsigned sum = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = lane * 4 + e;
                unsigned v = 0;
                if (k < dlen) {
                    const int r = k / side, c = k - r * side;
                    const int y = py[q] + r - radius, x = px[q] + c - radius;
                    int sob = 0;
                    if (y > 0 && y < height && x > 0 && x < width) {
                        const unsigned char* w0 = win + r * VISO_EXTRACT_WIN + c; /* window row r = image row y-1 */
                        sob = ((int)w0[2] + 2 * (int)w0[VISO_EXTRACT_WIN + 2] + (int)w0[2 * VISO_EXTRACT_WIN + 2]) -
                              ((int)w0[0] + 2 * (int)w0[VISO_EXTRACT_WIN] + (int)w0[2 * VISO_EXTRACT_WIN]);
                    }
                    v = (unsigned)(sob + 1024);
                }
                u[e] = v;
                sum += v;
            }
            uint2 w;
            w.x = u[0] | (u[1] << 16);
            w.y = u[2] | (u[3] << 16);
            reinterpret_cast<uint2*>(job.out + (size_t)row * VISO_DESC_U16)[lane] = w;
            const unsigned tot = warp_sum_u(sum);
            if (lane == 0) job.rsum[row] = tot;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ grid */

__device__ __forceinline__ int cell_coord(float v, int g)
{
    int c = __float2int_rd(v * (1.0f / VISO_GRID_CS));
    return min(max(c, 0), g - 1);
}

/*
 * Counting sort of one keypoint set into 16-px cells (one CTA per set).  Emits, in cell order, the candidate
 * records the matcher streams: srec = (x, y, original index, row sum).
 */
__global__ void __launch_bounds__(512) grid_build_kernel(const GridJob* __restrict__ jobs, GridCfg g)
{
    extern __shared__ int sm[];
    const int ncell = g.gx * g.gy;
    int* hist = sm;               /* ncell + 1 */
    int* cursor = sm + ncell + 1; /* ncell */
    const GridJob job = jobs[blockIdx.x];
    const int n = *job.n;
    for (int c = threadIdx.x; c <= ncell; c += blockDim.x) hist[c] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float2 p = job.xy[i];
        atomicAdd(&hist[cell_coord(p.y, g.gy) * g.gx + cell_coord(p.x, g.gx)], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) { /* exclusive scan: each lane owns a contiguous chunk */
        const int lane = threadIdx.x;
        const int chunk = (ncell + 31) / 32;
        const int b = lane * chunk, e = min(b + chunk, ncell);
        int s = 0;
        for (int c = b; c < e; ++c) s += hist[c];
        int incl = warp_incl_scan(s, lane);
        int run = incl - s;
        for (int c = b; c < e; ++c) { int h = hist[c]; hist[c] = run; run += h; }
        if (lane == 31) hist[ncell] = incl;
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= ncell; c += blockDim.x) {
        int v = hist[c];
        job.cell_start[c] = v;
        if (c < ncell) cursor[c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float2 p = job.xy[i];
        int pos = atomicAdd(&cursor[cell_coord(p.y, g.gy) * g.gx + cell_coord(p.x, g.gx)], 1);
        job.srec[pos] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), (unsigned)i, job.rsum[i]);
    }
}

/* ------------------------------------------------------------------------------------------------ match */

/* sampsonDistance + algebricDistance, viso.cpp:652-666, 390-407: literal operation order (no FMA) */
__device__ __forceinline__ double sampson_dev(const double* F, float p1x, float p1y, float p2x, float p2y)
{
    double Fx0 = F[0] * p1x + F[1] * p1y + F[2];
    double Fx1 = F[3] * p1x + F[4] * p1y + F[5];
    double Ftx0 = F[0] * p2x + F[3] * p2y + F[6];
    double Ftx1 = F[1] * p2x + F[4] * p2y + F[7];
    float a0 = p1x, a1 = p1y, a2 = 1.f, b0 = p2x, b1 = p2y, b2 = 1.f;
    double alg = b0 * F[0] * a0 + b0 * F[1] * a1 + b0 * F[2] * a2 +
                 b1 * F[3] * a0 + b1 * F[4] * a1 + b1 * F[5] * a2 +
                 b2 * F[6] * a0 + b2 * F[7] * a1 + b2 * F[8] * a2;
    float ad = (float)alg;
    float ad2 = __fmul_rn(ad, ad);
    return (double)ad2 / (Fx0 * Fx0 + Fx1 * Fx1 + Ftx0 * Ftx0 + Ftx1 * Ftx1);
}

__device__ __forceinline__ float l1_dist(float qx, float qy, float tx, float ty)
{
    return __fadd_rn(fabsf(__fsub_rn(tx, qx)), fabsf(__fsub_rn(ty, qy)));
}

__device__ __forceinline__ bool key_greater(float d1, int i1, float d2, int i2)
{
    return d1 > d2 || (d1 == d2 && i1 > i2);
}

/* per-warp shared scratch.  The top-K threshold search (hist / tie arrays) and the scanned-candidate list are never
 * live at the same time, so they share storage. */
struct WarpScratch {
    int rowS0[32];
    int rowPre[33];
    int pad_[3];
    union {
        uint4 list[VISO_LIST_CAP];             /* (target index, L1 distance bits, row sum, -) */
        struct {
            unsigned hist[VISO_HIST_BINS];
            float tieD[VISO_TIE_CAP];
            int tieI[VISO_TIE_CAP];
        } sel;
    };
};

struct QueryGeom {
    float qx, qy, r, slack;
    int cy0, cy1;
};

__device__ __forceinline__ float geom_slack(float qx, float qy, float r)
{
    return 1.0f + 4e-6f * (fabsf(qx) + fabsf(qy) + r);
}

__device__ __forceinline__ QueryGeom make_geom(GridCfg g, float qx, float qy, float r)
{
    QueryGeom q;
    q.qx = qx; q.qy = qy; q.r = r;
    q.slack = geom_slack(qx, qy, r);
    q.cy0 = cell_coord(qy - r - q.slack, g.gy);
    q.cy1 = cell_coord(qy + r + q.slack, g.gy);
    return q;
}

/* ---- candidate visitors: f(in, dist, rec) is called by all lanes of the warp, 32 points per call;
 *      rec = (x, y, original index, row sum) of this lane's point, dist its L1 distance to the query ---- */

/* Generic visitor: walks the spans of the target grid rows under the L1 diamond straight from global memory.
 * Handles any radius / point count. */
struct GlobalVisitor {
    const SetView& t;
    GridCfg g;
    QueryGeom q;
    WarpScratch& ws;
    int lane;
    int total0;
    bool one_group;

    /* Spans of the (up to 32) grid rows rg..rg+31 that overlap the diamond: lane l owns row rg+l.  Leaves the
     * flattened prefix table in ws and returns the number of points in those spans (same value in all lanes). */
    __device__ __forceinline__ int setup_rows(int rg)
    {
        const int cy = rg + lane;
        int s0 = 0, len = 0;
        if (cy <= q.cy1) {
            const float lo = (cy == 0) ? -CUDART_INF_F : (float)(cy * VISO_GRID_CS);
            const float hi = (cy == g.gy - 1) ? CUDART_INF_F : (float)((cy + 1) * VISO_GRID_CS);
            const float dymin = fmaxf(0.f, fmaxf(lo - q.qy, q.qy - hi));
            const float rem = q.r - dymin + q.slack;
            if (rem >= 0.f) {
                const int cx0 = cell_coord(q.qx - rem, g.gx), cx1 = cell_coord(q.qx + rem, g.gx);
                s0 = __ldg(t.cell_start + cy * g.gx + cx0);
                len = __ldg(t.cell_start + cy * g.gx + cx1 + 1) - s0;
            }
        }
        const int incl = warp_incl_scan(len, lane);
        __syncwarp();
        ws.rowS0[lane] = s0;
        ws.rowPre[lane] = incl - len;
        if (lane == 31) ws.rowPre[32] = incl;
        __syncwarp();
        return __shfl_sync(FULL, incl, 31);
    }

    /* upper bound on the in-radius count: the number of points in the visited spans */
    __device__ __forceinline__ int bound()
    {
        one_group = q.cy1 - q.cy0 < 32;
        total0 = setup_rows(q.cy0);
        int b = total0;
        if (!one_group)
            for (int rg = q.cy0 + 32; rg <= q.cy1; rg += 32) b += setup_rows(rg);
        return b;
    }

    template <class Fn> __device__ __forceinline__ void rows(int total, Fn&& f)
    {
        int row = 0;
        for (int base = 0; base < total; base += 32) {
            const int fl = base + lane;
            const bool in = fl < total;
            float dist = CUDART_INF_F;
            uint4 rec = make_uint4(0, 0, 0xffffffffu, 0);
            if (in) {
                while (fl >= ws.rowPre[row + 1]) ++row;
                const int p = ws.rowS0[row] + (fl - ws.rowPre[row]);
                rec = __ldg(t.srec + p);
                dist = l1_dist(q.qx, q.qy, __uint_as_float(rec.x), __uint_as_float(rec.y));
            }
            f(in, dist, rec);
        }
    }

    template <class Fn> __device__ __forceinline__ void all(Fn&& f)
    {
        if (one_group) { rows(total0, f); return; } /* the row table of the only group is still in place */
        for (int rg = q.cy0; rg <= q.cy1; rg += 32) {
            const int total = setup_rows(rg);
            rows(total, f);
        }
    }
};

__device__ __forceinline__ int dist_bin(float dist, float scale)
{
    return min(VISO_HIST_BINS - 1, (int)(dist * scale));
}

/* running result of one query (warp-uniform values) */
struct BestState {
    unsigned b1, b2;     /* smallest / second smallest SAD with multiplicity; 0xffffffff = none */
    unsigned bdist;      /* float bits of the L1 distance of the best (>= +0, so unsigned order == float order) */
    int bidx;
};

/*
 * Phase 2: exact SAD of the n listed candidates against the query, 32 candidates per batch.
 *
 * Eight lanes share one 256-byte descriptor row: lane `sub` of the group reads the 16-byte chunks sub and 8 + sub,
 * so each of the two LDG.128 of a step covers exactly one full 128-byte line per row (4 rows = 4 lines per
 * instruction, no partially used sector); four rows per step, eight steps per batch.  Each lane holds the matching
 * two chunks of the query row in registers (qa, qb).  Per element pair one VIMNMX.U16x2 + one
 * add:  sum|a-b| = sum(a) + sum(b) - 2*sum(min(a,b)), with the row sums precomputed by the pack kernel.  The
 * eight per-step partial sums of a lane are then transposed-reduced across the 8 lanes of a row group (7 SHFL), so
 * that lane (g, sub) ends up with the complete sum for candidate 4*sub + g of the batch, and the batch is folded
 * into the running (best, second best) with REDUX min/max -- ties on the SAD go to the largest (L1, index) key,
 * i.e. the last one in the reference's scan order (viso.cpp:703).
 */
/* one batch of NS*4 candidates starting at list entry `base` (NS = 8: up to 32, NS = 4: up to 16).  Branch free:
 * entries past the end of the list are clamped to the last one (a repeated L1-hit load) and masked afterwards, so
 * that all 2*NS row loads of the batch can be in flight together. */
/* where the candidates of a query come from: target index of entry e, and (index, L1 distance bits, row sum) */
struct ListAcc { /* the warp's uint4 list (generic kernel) */
    const WarpScratch& ws;
    __device__ __forceinline__ unsigned index(int e) const { return ws.list[e].x; }
    __device__ __forceinline__ uint4 entry(int e) const { return ws.list[e]; }
};
struct TileAcc { /* region indices into the staged neighbourhood (tile kernel): the distance is recomputed */
    const unsigned short* ql;
    const uint4* reg;
    float qx, qy;
    __device__ __forceinline__ unsigned index(int e) const { return reg[ql[e]].z; }
    __device__ __forceinline__ uint4 entry(int e) const
    {
        const uint4 r = reg[ql[e]];
        return make_uint4(r.z, __float_as_uint(l1_dist(qx, qy, __uint_as_float(r.x), __uint_as_float(r.y))), r.w, 0u);
    }
};

/* NB = butterfly width in steps (8: up to 32 candidates, 4: up to 16), NL <= NB = steps whose rows are actually
 * loaded and evaluated (the rest contribute 0 and are masked) */
template <int NB, int NL, class Acc>
__device__ __forceinline__ void eval_batch(const uint4* __restrict__ tbase, const Acc& acc_, int base, int n, int lane,
                                           const uint4& qa, const uint4& qb, unsigned qsum, BestState& st)
{
    const int sub = lane & 7, g = lane >> 3;
    const bool b0 = sub & 1, b1 = sub & 2, b2 = sub & 4;
    /* lane L fetches the record of candidate base + L once; the steps and the final fold get it by shuffle */
    const uint4 ent = acc_.entry(min(base + lane, n - 1));
    unsigned part[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) part[s] = 0;
#pragma unroll
    for (int h = 0; h < NL; h += VISO_EVAL_DEPTH) { /* VISO_EVAL_DEPTH steps = 2*VISO_EVAL_DEPTH row loads in flight */
        uint4 ra[VISO_EVAL_DEPTH], rb[VISO_EVAL_DEPTH];
#pragma unroll
        for (int s = 0; s < VISO_EVAL_DEPTH; ++s) {
            if (h + s < NL) {
                const unsigned idx = __shfl_sync(FULL, ent.x, 4 * (h + s) + g);
                const uint4* rp = tbase + (size_t)idx * (VISO_DESC_U16 / 8);
                ra[s] = __ldg(rp);
                rb[s] = __ldg(rp + 8);
            }
        }
#pragma unroll
        for (int s = 0; s < VISO_EVAL_DEPTH; ++s) {
            if (h + s < NL) {
                const unsigned acc = __vminu2(qa.x, ra[s].x) + __vminu2(qa.y, ra[s].y) + __vminu2(qa.z, ra[s].z) + __vminu2(qa.w, ra[s].w) +
                                     __vminu2(qb.x, rb[s].x) + __vminu2(qb.y, rb[s].y) + __vminu2(qb.z, rb[s].z) + __vminu2(qb.w, rb[s].w);
                part[h + s] = (acc & 0xffffu) + (acc >> 16); /* 16 elements x 2047 < 65536: no carry between halves */
            }
        }
    }
    /* transposed reduction over the 8 lanes of a row group: lane (g, sub) ends with candidate 4*step + g where
     * step = sub (NB = 8) or sub & 3 (NB = 4; lanes sub and sub^4 then hold the same candidate) */
    unsigned r2[2];
    if (NB == 8) {
        unsigned r4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned lo = part[2 * j], hi = part[2 * j + 1];
            r4[j] = (b0 ? hi : lo) + __shfl_xor_sync(FULL, b0 ? lo : hi, 1);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned lo = r4[2 * j], hi = r4[2 * j + 1];
            r2[j] = (b1 ? hi : lo) + __shfl_xor_sync(FULL, b1 ? lo : hi, 2);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned lo = part[2 * j], hi = part[2 * j + 1];
            r2[j] = (b0 ? hi : lo) + __shfl_xor_sync(FULL, b0 ? lo : hi, 1);
        }
    }
    unsigned tot;
    int slot; /* position of this lane's candidate inside the batch */
    bool mine = true;
    if (NB == 8) {
        tot = (b2 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, b2 ? r2[0] : r2[1], 4);
        slot = 4 * sub + g;
    } else {
        const unsigned t2 = (b1 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, b1 ? r2[0] : r2[1], 2);
        tot = t2 + __shfl_xor_sync(FULL, t2, 4);
        slot = 4 * (sub & 3) + g;
        mine = !b2; /* the duplicate lanes sit out */
    }
    const unsigned cidx = __shfl_sync(FULL, ent.x, slot), cdist = __shfl_sync(FULL, ent.y, slot);
    const unsigned csum = __shfl_sync(FULL, ent.z, slot);
    unsigned sad = 0xffffffffu, dbits = 0;
    int idx = -1;
    if (mine && base + slot < n) {
        sad = qsum + csum - 2u * tot;
        dbits = cdist;
        idx = (int)cidx;
    }
    const unsigned m1 = __reduce_min_sync(FULL, sad);
    const unsigned ties = __ballot_sync(FULL, sad == m1);
    unsigned m2 = m1;
    if (__popc(ties) < 2) m2 = __reduce_min_sync(FULL, sad == m1 ? 0xffffffffu : sad);
    const unsigned kd = __reduce_max_sync(FULL, sad == m1 ? dbits : 0u);
    const int ki = __reduce_max_sync(FULL, (sad == m1 && dbits == kd) ? idx : -1);
    if (m1 < st.b1) {
        st.b2 = min(st.b1, m2); st.b1 = m1; st.bdist = kd; st.bidx = ki;
    } else if (m1 == st.b1) {
        st.b2 = st.b1;
        if (kd > st.bdist || (kd == st.bdist && ki > st.bidx)) { st.bdist = kd; st.bidx = ki; }
    } else if (m1 < st.b2) {
        st.b2 = m1;
    }
}

template <class Acc>
__device__ __forceinline__ void eval_list(const uint16_t* __restrict__ tdesc, const Acc& acc, int n, int lane,
                                          const uint4& qa, const uint4& qb, unsigned qsum, BestState& st)
{
    const uint4* tbase = reinterpret_cast<const uint4*>(tdesc) + (lane & 7);
    int base = 0;
    for (; n - base > 24; base += 32) eval_batch<8, 8>(tbase, acc, base, n, lane, qa, qb, qsum, st);
    const int rest = n - base; /* 0..24: rows are loaded in units of 8 candidates */
    if (rest > 16) eval_batch<8, 6>(tbase, acc, base, n, lane, qa, qb, qsum, st);
    else if (rest > 8) eval_batch<4, 4>(tbase, acc, base, n, lane, qa, qb, qsum, st);
    else if (rest > 0) eval_batch<4, 2>(tbase, acc, base, n, lane, qa, qb, qsum, st);
}

/* viso.cpp:711-722: the ratio test and the dense output record (best_idx, best_d1, best_d2, valid) */
__device__ __forceinline__ void write_result(const MatchJob& job, const MatchParamsDev& P, int q, const BestState& st)
{
    int valid = 0;
    const int b1 = st.bidx >= 0 ? (int)st.b1 : INT_MAX;
    const int b2 = st.b2 == 0xffffffffu ? INT_MAX : (int)st.b2;
    if (st.bidx >= 0) {
        if (P.second_best) {
            const double d2 = (b2 == INT_MAX) ? 1.7976931348623157e308 : (double)b2;
            valid = ((double)b1 < d2 * P.ratio) ? 1 : 0; /* viso.cpp:715 */
        } else
            valid = 1;
    }
    job.out[q] = make_int4(st.bidx, b1, b2, valid);
}

/*
 * match_desc for one query, viso.cpp:686-722, by one warp.
 *
 * Reference semantics restated set-wise: with D0 = L1(query, target 0) if that is <= radius
 * (else +inf), the scanned candidates are the K smallest keys (L1, index) among
 *     L = { j : L1_j <= radius and L1_j < D0 }
 * (target 0 terminates the reference's scan, viso.cpp:693, and sorts first inside its distance group, so exactly
 * the strictly closer points are scanned).  Over that set: best = min SAD, ties to the LARGEST key (the last one
 * in scan order, viso.cpp:703), best_d2 = second smallest SAD with multiplicity.  Sampson-gated candidates
 * (viso.cpp:695-701) still occupy a top-K slot but are not compared.  All of it is order independent.
 *
 * Phase 1: candidate generation through the visitor V (32 points per step, one per lane); when more than K points
 * can be in range the exact top-K cut is found with a 128-bin histogram over the L1 distance plus an exact rank
 * search inside the cut bin; the Sampson gate is applied and the survivors are appended to the per-warp list.
 * Phase 2: eval_list() whenever the list may overflow on the next step, and once at the end.
 * Returns the number of (query, candidate) pairs that reached the SAD.
 */
template <class V>
__device__ __forceinline__ unsigned match_query(V& vis, const MatchJob& job, const MatchParamsDev& P, WarpScratch& ws,
                                                int lane, const uint4 qrec)
{
    const int q = (int)qrec.z;
    const float qx = __uint_as_float(qrec.x), qy = __uint_as_float(qrec.y);
    const unsigned qsum = qrec.w;
    const uint4* qp = reinterpret_cast<const uint4*>(job.q.desc + (size_t)q * VISO_DESC_U16) + (lane & 7);
    const uint4 qa = __ldg(qp), qb = __ldg(qp + 8);
    const float r = P.radius;
    const int K = P.K;
    const float bscale = (float)VISO_HIST_BINS / (r + 1.0f);
    unsigned pairs = 0;

    BestState st;
    st.b1 = 0xffffffffu; st.b2 = 0xffffffffu; st.bdist = 0; st.bidx = -1;

    /* index 0 terminator */
    const float2 t0 = __ldg(job.t.xy);
    const float d0 = l1_dist(qx, qy, t0.x, t0.y);
    const float D0 = (d0 <= r) ? d0 : CUDART_INF_F;

    /* top-K threshold: (Tbin, Td, Ti); candidates with bin < Tbin, or bin == Tbin and key <= (Td,Ti) */
    int Tbin = INT_MAX;
    float Td = CUDART_INF_F;
    int Ti = INT_MAX;
    if (vis.bound() > K) {
        for (int b = lane; b < VISO_HIST_BINS; b += 32) ws.sel.hist[b] = 0;
        __syncwarp();
        int cnt = 0;
        vis.all([&](bool in, float dist, uint4 rec) {
            const bool inL = in && dist <= r && dist < D0;
            if (inL) atomicAdd(&ws.sel.hist[dist_bin(dist, bscale)], 1u);
            cnt += __popc(__ballot_sync(FULL, inL));
        });
        __syncwarp();
        if (cnt > K) {
            /* find the bin where the cumulative count reaches K */
            unsigned c[4];
            unsigned s = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) { c[e] = ws.sel.hist[lane * 4 + e]; s += c[e]; }
            const int incl = warp_incl_scan((int)s, lane);
            const unsigned hit = __ballot_sync(FULL, incl >= K);
            const int hl = __ffs(hit) - 1;
            int tb = 0, before = 0;
            if (lane == hl) {
                int run = incl - (int)s;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (run + (int)c[e] >= K) { tb = lane * 4 + e; before = run; break; }
                    run += (int)c[e];
                }
            }
            tb = __shfl_sync(FULL, tb, hl);
            before = __shfl_sync(FULL, before, hl);
            const int nb = (int)ws.sel.hist[tb];
            const int m = K - before; /* 1..nb keys of bin tb are kept */
            Tbin = tb;
            if (m < nb) {
                if (nb <= VISO_TIE_CAP) {
                    int fill = 0;
                    vis.all([&](bool in, float dist, uint4 rec) {
                        const bool hitb = in && dist <= r && dist < D0 && dist_bin(dist, bscale) == tb;
                        const unsigned bm = __ballot_sync(FULL, hitb);
                        if (hitb) {
                            const int o = fill + __popc(bm & ((1u << lane) - 1));
                            ws.sel.tieD[o] = dist;
                            ws.sel.tieI[o] = (int)rec.z;
                        }
                        fill += __popc(bm);
                    });
                    __syncwarp();
                    /* the key of rank m-1 inside the bin */
                    float selD = 0.f; int selI = 0; bool have = false;
                    for (int e = lane; e < nb; e += 32) {
                        const float de = ws.sel.tieD[e]; const int ie = ws.sel.tieI[e];
                        int rank = 0;
                        for (int o = 0; o < nb; ++o) rank += key_greater(de, ie, ws.sel.tieD[o], ws.sel.tieI[o]) ? 1 : 0;
                        if (rank == m - 1) { selD = de; selI = ie; have = true; }
                    }
                    const unsigned hm = __ballot_sync(FULL, have);
                    const int sl = __ffs(hm) - 1;
                    Td = __shfl_sync(FULL, selD, sl);
                    Ti = __shfl_sync(FULL, selI, sl);
                } else {
                    /* pathological tie bin: m successive minimum searches (exact, slow) */
                    float curD = -1.f; int curI = -1;
                    for (int it = 0; it < m; ++it) {
                        float bestD = CUDART_INF_F; int bestI = INT_MAX;
                        vis.all([&](bool in, float dist, uint4 rec) {
                            const int idx = (int)rec.z;
                            if (in && dist <= r && dist < D0 && dist_bin(dist, bscale) == tb &&
                                key_greater(dist, idx, curD, curI) && key_greater(bestD, bestI, dist, idx)) {
                                bestD = dist; bestI = idx;
                            }
                        });
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const float od = __shfl_xor_sync(FULL, bestD, o);
                            const int oi = __shfl_xor_sync(FULL, bestI, o);
                            if (key_greater(bestD, bestI, od, oi)) { bestD = od; bestI = oi; }
                        }
                        curD = bestD; curI = bestI;
                    }
                    Td = curD; Ti = curI;
                }
            }
        }
        __syncwarp(); /* the selection scratch is dead from here on: its storage becomes the list */
    }

    /* final pass: membership, Sampson gate, append to the list (membership of a point does not depend on the
     * others once the threshold is known, so the list can be evaluated and reset at any time) */
    int nlist = 0;
    vis.all([&](bool in, float dist, uint4 rec) {
        const int idx = (int)rec.z;
        bool take = in && dist <= r && dist < D0;
        if (take && Tbin != INT_MAX) {
            const int bin = dist_bin(dist, bscale);
            take = bin < Tbin || (bin == Tbin && !key_greater(dist, idx, Td, Ti));
        }
        if (take && P.epipolar) {
            const double sd = sampson_dev(P.F, qx, qy, __uint_as_float(rec.x), __uint_as_float(rec.y));
            if (!isfinite(sd) || sd > P.sampson_thresh) take = false;
        }
        const unsigned tm = __ballot_sync(FULL, take);
        if (tm == 0) return;
        if (take) ws.list[nlist + __popc(tm & ((1u << lane) - 1))] = make_uint4(rec.z, __float_as_uint(dist), rec.w, 0u);
        nlist += __popc(tm);
        if (nlist > VISO_LIST_CAP - 32) {
            __syncwarp();
            eval_list(job.t.desc, ListAcc{ws}, nlist, lane, qa, qb, qsum, st);
            pairs += nlist;
            nlist = 0;
            __syncwarp();
        }
    });
    if (nlist > 0) {
        __syncwarp();
        eval_list(job.t.desc, ListAcc{ws}, nlist, lane, qa, qb, qsum, st);
        pairs += nlist;
        __syncwarp();
    }

    if (lane == 0) write_result(job, P, q, st);
    return pairs;
}

/*
 * sad_match_kernel: match_desc (viso.cpp:668-722) for a batch of jobs.  blockIdx.y = job, blockIdx.x = query tile
 * (VISO_TILE_W x VISO_TILE_H cells of the QUERY set's grid, 96 x 64 px).
 *
 * 1. Staging.  The candidate records (x, y, index, row sum) of every target cell that can hold a neighbour of any of
 *    the tile's queries -- the bounding box of the tile's query coordinates grown by radius + slack, clamped exactly
 *    like the per-query geometry -- are copied to shared memory, one contiguous span per grid row (coalesced).
 * 2. Candidate generation, LANE = QUERY.  Every warp walks a quarter of the staged points; the point is a shared
 *    memory broadcast and each lane tests it against its own query (radius and index-0 terminator), appending hits
 *    to that query's list (shared-memory counter).  ~9 instructions per 32 (query, point) tests and no ballots.
 * 3. Evaluation, WARP = QUERY.  The list is gathered into candidate records (Sampson gate for the stereo mode,
 *    lanes = candidates) and handed to eval_list().
 * Queries whose neighbourhood holds more than max_neighbors points (the top-K cut is needed) or overflows the list,
 * and tiles whose neighbourhood does not fit the staging buffer, are marked VISO_PENDING and counted; the generic
 * kernel sad_match_generic_kernel, launched right after, completes exactly those (and exits at once when there are
 * none).
 */
__global__ void __launch_bounds__(VISO_MATCH_WARPS * 32, VISO_MATCH_MINB)
sad_match_kernel(const MatchJob* __restrict__ jobs, MatchParamsPair mp, GridCfg g, int reg_cap, int ql_cap,
                 unsigned long long* sad_pairs, int* n_pending)
{
    extern __shared__ uint4 reg[];                                /* staged neighbourhood, reg_cap records */
    /* per-query candidate lists (region indices), 32 x (ql_cap + 2): the +2 makes the word stride odd so that
     * lanes = queries write without bank conflicts */
    unsigned short* const qlist = reinterpret_cast<unsigned short*>(reg + reg_cap);
    const int ql_stride = ql_cap + 2;
    __shared__ int qcnt[32];
    __shared__ int row_off[VISO_MAX_REG_ROWS + 1];
    __shared__ int tile_s[4];
    __shared__ uint4 qrec_s[32];

    const MatchJob job = jobs[blockIdx.y];
    const MatchParamsDev& P = mp.p[job.mode];
    const int nq = *job.q.n, nt = *job.t.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (g.gx + VISO_TILE_W - 1) / VISO_TILE_W;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    if (ty * VISO_TILE_H >= g.gy || nq <= 0) return;

    /* the tile's queries: one span of the cell-sorted query array per cell row */
    const int cx_lo = tx * VISO_TILE_W, cx_hi = min(cx_lo + VISO_TILE_W, g.gx);
    int qtot = 0;
    int qs[VISO_TILE_H], ql[VISO_TILE_H];
#pragma unroll
    for (int rr = 0; rr < VISO_TILE_H; ++rr) {
        const int cy = ty * VISO_TILE_H + rr;
        qs[rr] = 0; ql[rr] = 0;
        if (cy < g.gy) {
            qs[rr] = __ldg(job.q.cell_start + cy * g.gx + cx_lo);
            ql[rr] = __ldg(job.q.cell_start + cy * g.gx + cx_hi) - qs[rr];
        }
        qtot += ql[rr];
    }
    if (qtot == 0) return;

    auto query_rec = [&](int k) {
        int pos = 0;
#pragma unroll
        for (int rr = 0; rr < VISO_TILE_H; ++rr) {
            if (k >= 0 && k < ql[rr]) pos = qs[rr] + k;
            k -= ql[rr];
        }
        return __ldg(job.q.srec + pos);
    };

    if (nt <= 0) { /* no targets: every query is unmatched */
        for (int k = threadIdx.x; k < qtot; k += blockDim.x)
            job.out[query_rec(k).z] = make_int4(-1, INT_MAX, INT_MAX, 0);
        return;
    }

    /* Neighbourhood of the tile (warp 0): bounding box of the tile's query coordinates (queries outside the image
     * extent are clamped into border cells, so the cell rectangle is not a bound), grown by radius + the largest
     * per-query slack (make_geom) + a margin far above the float rounding of the sums; then the staged span of
     * every grid row. */
    const float r = P.radius;
    if (warp == 0) {
        float xmin = CUDART_INF_F, xmax = -CUDART_INF_F, ymin = CUDART_INF_F, ymax = -CUDART_INF_F, amax = 0.f;
        for (int k = lane; k < qtot; k += 32) {
            const uint4 qr = query_rec(k);
            const float x = __uint_as_float(qr.x), y = __uint_as_float(qr.y);
            xmin = fminf(xmin, x); xmax = fmaxf(xmax, x); ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
            amax = fmaxf(amax, fabsf(x) + fabsf(y));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xmin = fminf(xmin, __shfl_xor_sync(FULL, xmin, o)); xmax = fmaxf(xmax, __shfl_xor_sync(FULL, xmax, o));
            ymin = fminf(ymin, __shfl_xor_sync(FULL, ymin, o)); ymax = fmaxf(ymax, __shfl_xor_sync(FULL, ymax, o));
            amax = fmaxf(amax, __shfl_xor_sync(FULL, amax, o));
        }
        const float grow = r + (1.0f + 4e-6f * (amax + r)) + 4e-6f * (amax + r) + 1e-3f;
        const int cx0 = cell_coord(xmin - grow, g.gx), cx1 = cell_coord(xmax + grow, g.gx);
        const int cy0 = cell_coord(ymin - grow, g.gy), cy1 = cell_coord(ymax + grow, g.gy);
        const int nr = cy1 - cy0 + 1;
        int total = -1; /* -1: no staging */
        if (nr <= VISO_MAX_REG_ROWS && xmin == xmin && ymin == ymin && reg_cap > 0) {
            int run = 0;
            for (int b = 0; b < nr; b += 32) {
                const int rr = b + lane;
                int len = 0;
                if (rr < nr) {
                    const int cy = cy0 + rr;
                    len = __ldg(job.t.cell_start + cy * g.gx + cx1 + 1) - __ldg(job.t.cell_start + cy * g.gx + cx0);
                }
                const int incl = warp_incl_scan(len, lane);
                if (rr < nr) row_off[rr] = run + incl - len;
                run += __shfl_sync(FULL, incl, 31);
            }
            if (lane == 0) row_off[nr] = run;
            if (run <= reg_cap) total = run;
        }
        if (lane == 0) { tile_s[0] = cx0; tile_s[1] = cy0; tile_s[2] = nr; tile_s[3] = total; }
    }
    __syncthreads();
    const int rcx0 = tile_s[0], rcy0 = tile_s[1], nrows = tile_s[2];
    const bool tile_ok = tile_s[3] >= 0;

    unsigned pairs = 0;
    if (!tile_ok) { /* leave the whole tile to the generic kernel */
        for (int k = threadIdx.x; k < qtot; k += blockDim.x) job.out[query_rec(k).z] = make_int4(0, 0, 0, VISO_PENDING);
        if (threadIdx.x == 0) atomicAdd(n_pending, qtot);
    } else {
        const int R = tile_s[3];
        for (int rr = warp; rr < nrows; rr += VISO_MATCH_WARPS) {
            const int o = row_off[rr], len = row_off[rr + 1] - o;
            const uint4* src = job.t.srec + __ldg(job.t.cell_start + (rcy0 + rr) * g.gx + rcx0);
            for (int i = lane; i < len; i += 32) reg[o + i] = __ldg(src + i);
        }
        const float2 t0 = __ldg(job.t.xy);
        for (int g0 = 0; g0 < qtot; g0 += 32) { /* groups of 32 queries: lane = query */
            __syncthreads(); /* staging done (first round) / the previous group's lists are consumed */
            if (threadIdx.x < 32) qcnt[threadIdx.x] = 0;
            __syncthreads();
            {
                const int k = g0 + lane;
                const bool act = k < qtot;
                const uint4 qr = query_rec(act ? k : g0);
                if (warp == 0) qrec_s[lane] = qr;
                const float qx = __uint_as_float(qr.x), qy = __uint_as_float(qr.y);
                const float d0 = l1_dist(qx, qy, t0.x, t0.y);
                /* limit = min(radius, strictly below D0): candidates need dist <= r and dist < D0 (index-0 rule) */
                const float D0 = (d0 <= r) ? d0 : CUDART_INF_F;
                const float2* pts = reinterpret_cast<const float2*>(reg);
#pragma unroll 4
                for (int i = warp; i < R; i += VISO_MATCH_WARPS) {
                    const float2 p = pts[2 * i]; /* (x, y) of the uint4 record: broadcast */
                    const float dist = l1_dist(qx, qy, p.x, p.y);
                    if (act && dist <= r && dist < D0) {
                        const int j = atomicAdd(&qcnt[lane], 1);
                        if (j < ql_cap) qlist[lane * ql_stride + j] = (unsigned short)i;
                    }
                }
            }
            __syncthreads();
            const int gq = min(32, qtot - g0);
            for (int kk = warp; kk < gq; kk += VISO_MATCH_WARPS) {
                const uint4 qrec = qrec_s[kk];
                const int n = qcnt[kk];
                if (n > ql_cap || n > P.K) { /* top-K cut or list overflow: left to the generic kernel */
                    if (lane == 0) {
                        job.out[qrec.z] = make_int4(0, 0, 0, VISO_PENDING);
                        atomicAdd(n_pending, 1);
                    }
                    continue;
                }
                const int q = (int)qrec.z;
                const float qx = __uint_as_float(qrec.x), qy = __uint_as_float(qrec.y);
                const uint4* qp = reinterpret_cast<const uint4*>(job.q.desc + (size_t)q * VISO_DESC_U16) + (lane & 7);
                const uint4 qa = __ldg(qp), qb = __ldg(qp + 8);
                BestState st;
                st.b1 = 0xffffffffu; st.b2 = 0xffffffffu; st.bdist = 0; st.bidx = -1;
                unsigned short* ql = qlist + kk * ql_stride;
                int nlist = n;
                if (P.epipolar) { /* Sampson gate (viso.cpp:695-701), lanes = candidates, compacting the list in place */
                    nlist = 0;
                    for (int base = 0; base < n; base += 32) {
                        const int e = base + lane;
                        bool take = e < n;
                        const unsigned short ri = ql[take ? e : 0];
                        if (take) {
                            const uint4 rec = reg[ri];
                            const double sd = sampson_dev(P.F, qx, qy, __uint_as_float(rec.x), __uint_as_float(rec.y));
                            if (!isfinite(sd) || sd > P.sampson_thresh) take = false;
                        }
                        const unsigned tm = __ballot_sync(FULL, take);
                        __syncwarp(); /* every lane has read its entry before the slots are reused */
                        if (take) ql[nlist + __popc(tm & ((1u << lane) - 1))] = ri;
                        nlist += __popc(tm);
                    }
                    __syncwarp();
                }
                if (nlist > 0) eval_list(job.t.desc, TileAcc{ql, reg, qx, qy}, nlist, lane, qa, qb, qrec.w, st);
                pairs += nlist;
                if (lane == 0) write_result(job, P, q, st);
            }
        }
    }
    if (sad_pairs && lane == 0 && pairs) {
        atomicAdd(sad_pairs, (unsigned long long)pairs);
        atomicAdd(sad_pairs + 1, (unsigned long long)pairs);
    }
}

/*
 * Generic match kernel: CTA = VISO_STRIP_QPC consecutive cell-sorted queries, one warp per query, every query walks
 * its own grid-row spans in global memory and applies the exact top-K cut (match_query<GlobalVisitor>).  Handles any
 * radius, max_neighbors and density.  With only_pending it completes the queries the tile kernel marked
 * VISO_PENDING; VISO_MATCH_MODE=generic runs everything through it (A/B measurements and tests).
 */
__global__ void __launch_bounds__(VISO_MATCH_WARPS * 32)
sad_match_generic_kernel(const MatchJob* __restrict__ jobs, MatchParamsPair mp, GridCfg g, unsigned long long* sad_pairs,
                         const int* __restrict__ n_pending, int only_pending)
{
    __shared__ WarpScratch wscr[VISO_MATCH_WARPS];
    if (only_pending && *n_pending == 0) return;
    const MatchJob job = jobs[blockIdx.y];
    const MatchParamsDev& P = mp.p[job.mode];
    const int nq = *job.q.n, nt = *job.t.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch& ws = wscr[warp];
    unsigned pairs = 0;
    for (int chunk = blockIdx.x; chunk * VISO_STRIP_QPC < nq; chunk += gridDim.x)
    for (int qi = chunk * VISO_STRIP_QPC + warp; qi < min(nq, (chunk + 1) * VISO_STRIP_QPC); qi += VISO_MATCH_WARPS) {
        const uint4 qrec = __ldg(job.q.srec + qi);
        if (only_pending && job.out[qrec.z].w != VISO_PENDING) continue;
        if (nt <= 0) {
            if (lane == 0) job.out[qrec.z] = make_int4(-1, INT_MAX, INT_MAX, 0);
            continue;
        }
        GlobalVisitor vis{job.t, g, make_geom(g, __uint_as_float(qrec.x), __uint_as_float(qrec.y), P.radius), ws, lane, 0, true};
        pairs += match_query(vis, job, P, ws, lane, qrec);
    }
    if (sad_pairs && lane == 0 && pairs) {
        atomicAdd(sad_pairs, (unsigned long long)pairs);
        atomicAdd(sad_pairs + 1, (unsigned long long)pairs);
    }
}

/* ------------------------------------------------------------------------------------------------ sort */

/*
 * std::sort order, in parallel.  The reference sorts the Match vector with libstdc++'s UNSTABLE std::sort
 * (viso.cpp:724); which of several equal distances comes first is a property of that algorithm, so the device
 * reproduces the algorithm's data movement exactly instead of using its own sort (introsort.h is the sequential
 * restatement, checked against the real std::sort on the CPU).  Two observations make it parallel:
 *
 *  (1) __unguarded_partition(first+1, last, pivot) swaps the k-th element >= pivot from the left (position L_k) with
 *      the k-th element <= pivot from the right (position R_k) for k = 1..K, K = #{k : L_k < R_k}, and returns
 *      cut = min(L_{K+1}, R_K): both scans only ever read positions the swaps have not touched yet, so the pairing
 *      is a function of the ORIGINAL values.  A warp computes the two position lists with ballots, K with one
 *      monotone predicate, and does all swaps at once.
 *  (2) __final_insertion_sort never moves an element across the boundary of a final partition piece (everything to
 *      the left is <=), and inside a piece it is a stable insertion sort.  So once the <=16-element pieces are known
 *      every element computes its stable rank inside its piece, all in parallel.
 *
 * The median-of-3 pivot moves (3 reads, 1 swap) stay with lane 0; the depth-limit heapsort branch
 * (__partial_sort, taken only by adversarial inputs) is run sequentially by lane 0 with the restated code.
 * Segments are disjoint, so processing the right-hand pieces from a stack instead of by recursion does not change
 * the result.
 */
struct SortScratch {
    int stk_first[64], stk_last[64], stk_depth[64];
};

/* partition [first+1, last) around the pivot at p[first]; returns cut.  posL / posR: scratch, last-first entries */
__device__ __forceinline__ int warp_partition(viso_sort::KV* p, int first, int last, unsigned short* posL,
                                              unsigned short* posR, int lane)
{
    const int pv = p[first].d;
    int cl = 0, cr = 0;
    for (int base = first + 1; base < last; base += 32) {
        const int i = base + lane;
        const bool in = i < last;
        const int d = in ? p[i].d : 0;
        const bool fL = in && !(d < pv), fR = in && !(pv < d);
        const unsigned mL = __ballot_sync(FULL, fL), mR = __ballot_sync(FULL, fR);
        const unsigned lt = (1u << lane) - 1;
        if (fL) posL[cl + __popc(mL & lt)] = (unsigned short)(i - first);
        if (fR) posR[cr + __popc(mR & lt)] = (unsigned short)(i - first);
        cl += __popc(mL);
        cr += __popc(mR);
    }
    __syncwarp();
    /* R_k (k-th from the right) = posR[cr - k]; L_k = posL[k - 1] */
    const int kmax = min(cl, cr);
    int K = 0;
    for (int k0 = 0; k0 < kmax; k0 += 32) {
        const int k = k0 + lane;
        const bool ok = k < kmax && posL[k] < posR[cr - 1 - k];
        const unsigned m = __ballot_sync(FULL, ok);
        K += __popc(m);
        if (m != FULL) break; /* monotone: the first failure ends it */
    }
    for (int k = lane; k < K; k += 32) {
        const int a = first + posL[k], b = first + posR[cr - 1 - k];
        const viso_sort::KV t = p[a]; p[a] = p[b]; p[b] = t;
    }
    int cut = INT_MAX;
    if (K < cl) cut = first + posL[K];
    if (K > 0) cut = min(cut, first + (int)posR[cr - K]);
    __syncwarp();
    return cut;
}

/* __introsort_loop for p[0..n) by one warp; marks the first position of every final piece in leaf[] (pieces sorted
 * by the heapsort branch are marked element by element: they are already in order) */
__device__ void warp_introsort_loop(viso_sort::KV* p, int n, unsigned short* posL, unsigned short* posR,
                                    unsigned char* leaf, SortScratch& sc, int lane)
{
    if (n <= 0) return;
    int sp = 0;
    if (lane == 0) { sc.stk_first[0] = 0; sc.stk_last[0] = n; sc.stk_depth[0] = viso_sort::lg(n) * 2; }
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        --sp;
        int first = sc.stk_first[sp], last = sc.stk_last[sp], depth = sc.stk_depth[sp];
        __syncwarp();
        bool heap_done = false;
        while (last - first > 16) {
            if (depth == 0) {
                if (lane == 0) viso_sort::heap_sort_(p + first, last - first);
                for (int i = first + lane; i < last; i += 32) leaf[i] = 1;
                __syncwarp();
                heap_done = true;
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            if (lane == 0) viso_sort::median_to_first_(p, first, first + 1, mid, last - 1);
            __syncwarp();
            const int cut = warp_partition(p, first, last, posL + first, posR + first, lane);
            if (lane == 0) { sc.stk_first[sp] = cut; sc.stk_last[sp] = last; sc.stk_depth[sp] = depth; }
            ++sp;
            __syncwarp();
            last = cut;
        }
        if (!heap_done && last > first && lane == 0) leaf[first] = 1;
    }
    __syncwarp();
}

/*
 * Per frame: (1) compaction of the valid dense results in query order into Match(i, best_idx, best_d1)
 * (viso.cpp:711-722), (2) the reference's std::sort order (viso.cpp:724), (3) pos_of_query inverse table,
 * collect_matches (viso.cpp:501-514) and triangulate_rectified<double> (viso.cpp:1146-1152).
 *
 * When the frame's matches fit the CTA's shared memory (smem_cap of them, 13 bytes each) 8-byte (dist, query)
 * records are sorted there: warp 0 runs the parallel introsort loop, then every thread places one element with its
 * stable rank inside its final piece.  The algorithm only looks at dist, so the resulting permutation is the one
 * std::sort gives the Match vector.  Larger inputs are sorted in place in global memory by one thread with the
 * sequential restatement.
 */
__global__ void __launch_bounds__(128) compact_sort_kernel(const SortJob* __restrict__ jobs, ParamDev P, int smem_cap)
{
    extern __shared__ int sort_sm[];
    __shared__ int warp_tot[32];
    __shared__ SortScratch sc;
    viso_sort::KV* kv = reinterpret_cast<viso_sort::KV*>(sort_sm);                    /* [smem_cap] */
    unsigned short* posL = reinterpret_cast<unsigned short*>(sort_sm + 2 * smem_cap); /* [smem_cap] */
    unsigned short* posR = posL + smem_cap;                                            /* [smem_cap] */
    unsigned char* leaf = reinterpret_cast<unsigned char*>(posR + smem_cap);           /* [smem_cap] */
    const SortJob job = jobs[blockIdx.x];
    const int n = *job.n;
    int base = 0;
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        int4 r = make_int4(0, 0, 0, 0);
        if (i < n) {
            r = job.dense[i];
            if (job.pos_of_query) job.pos_of_query[i] = -1;
        }
        const bool flag = i < n && r.w != 0;
        const int slot = block_compact_slot(flag, base, warp_tot);
        if (flag) {
            if (slot < smem_cap) { kv[slot].d = r.y; kv[slot].pos = i; leaf[slot] = 0; }
            job.matches[3 * slot + 0] = i;
            job.matches[3 * slot + 1] = r.x;
            job.matches[3 * slot + 2] = r.y;
        }
    }
    const int M = base;
    const bool in_smem = M <= smem_cap;
    __syncthreads();
    if (in_smem) {
        if (threadIdx.x < 32) {
            if (M > 16) warp_introsort_loop(kv, M, posL, posR, leaf, sc, threadIdx.x);
            else if (M > 0 && threadIdx.x == 0) leaf[0] = 1;
        }
    } else if (threadIdx.x == 0) {
        viso_sort::sort(reinterpret_cast<viso_sort::M3*>(job.matches), M);
    }
    if (threadIdx.x == 0) *job.count = M;
    __syncthreads();
    for (int e = threadIdx.x; e < M; e += blockDim.x) {
        int p, i1, i2, d;
        if (in_smem) {
            /* __final_insertion_sort: stable rank inside the final piece [lo, hi) */
            int lo = e, hi = e + 1;
            while (!leaf[lo]) --lo;
            while (hi < M && !leaf[hi]) ++hi;
            const viso_sort::KV me = kv[e];
            int rank = 0;
            for (int j = lo; j < hi; ++j) {
                const int dj = kv[j].d;
                rank += (dj < me.d || (dj == me.d && j < e)) ? 1 : 0;
            }
            p = lo + rank;
            i1 = me.pos; d = me.d;
            i2 = job.dense[i1].x;
        } else {
            p = e;
            i1 = job.matches[3 * p]; i2 = job.matches[3 * p + 1]; d = job.matches[3 * p + 2];
        }
        if (in_smem) { job.matches[3 * p] = i1; job.matches[3 * p + 1] = i2; job.matches[3 * p + 2] = d; }
        if (job.pos_of_query) job.pos_of_query[i1] = p;
        if (job.x) {
            const float2 a = job.kp1[i1], b = job.kp2[i2];
            const double u1 = a.x, v1 = a.y, u2 = b.x, v2 = b.y;
            const int S = job.stride;
            job.x[0 * S + p] = u1; job.x[1 * S + p] = v1; job.x[2 * S + p] = u2; job.x[3 * S + p] = v2;
            if (job.X) {
                const double dd = u1 - u2;
                job.X[0 * S + p] = P.base * (u1 - P.cu) / dd;
                job.X[1 * S + p] = P.base * (v1 - P.cv) / dd;
                job.X[2 * S + p] = P.f * P.base / dd;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ circle */

/*
 * match_circle, viso.cpp:206-243, for match lists produced by match_desc (unique query index per list): the four
 * nested scans collapse to table lookups -- match11 and match22 are read from the dense per-query results, the
 * position k in match_lr_prev from pos_of_query of the previous frame.  Output order = ascending position i in
 * match_lr, as in the reference.  Also gathers x_c / Xp_c (viso.cpp:1291-1305).
 */
__global__ void __launch_bounds__(256) circle_kernel(const CircleJob* __restrict__ jobs)
{
    __shared__ int warp_tot[32];
    const CircleJob job = jobs[blockIdx.x];
    const int M = *job.lr_count, Mp = *job.lrp_count, npl = *job.n_prev_left;
    const int S = job.stride;
    int base = 0;
    for (int start = 0; start < M; start += blockDim.x) {
        const int i = start + threadIdx.x;
        bool flag = false;
        int il = 0, ir = 0, ilp = 0, irp = 0, k = 0;
        if (i < M) {
            il = job.lr[3 * i]; ir = job.lr[3 * i + 1];
            const int4 a = job.m11[il];
            if (a.w) {
                ilp = a.x;
                if (ilp >= 0 && ilp < npl) {
                    k = job.pos_prev[ilp];
                    if (k >= 0 && k < Mp) {
                        irp = job.lrp[3 * k + 1];
                        const int4 b = job.m22[ir];
                        flag = b.w && b.x == irp;
                    }
                }
            }
        }
        const int c = block_compact_slot(flag, base, warp_tot);
        if (flag) {
            job.circ4[4 * c] = il; job.circ4[4 * c + 1] = ir; job.circ4[4 * c + 2] = ilp; job.circ4[4 * c + 3] = irp;
            job.pcl2[2 * c] = i; job.pcl2[2 * c + 1] = k;
#pragma unroll
            for (int r = 0; r < 4; ++r) job.x_c[r * S + c] = job.x[r * S + i];
#pragma unroll
            for (int r = 0; r < 3; ++r) job.Xp_c[r * S + c] = job.Xp[r * S + k];
        }
    }
    if (threadIdx.x == 0) {
        *job.n_circ = base;
        viso_record_dev rec;
        for (int j = 0; j < 6; ++j) rec.tr[j] = 0;
        rec.ok = 0; rec.n_inliers = 0; rec.n_circ = base; rec.best_hyp = -1;
        *job.rec = rec;
    }
}

/* generic (standalone) variant working from explicit lookup tables, for viso_match_circle() */
__global__ void circle_tables_kernel(const int* __restrict__ m, int n, int* table, int table_n, int key_col, int val_mode,
                                     int* err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int key = m[3 * i + key_col];
    if (key < 0 || key >= table_n) return;
    const int val = val_mode ? i : m[3 * i + 1];
    const int old = atomicCAS(&table[key], -1, val);
    if (old != -1) atomicOr(err, 2);
}

__global__ void __launch_bounds__(256)
circle_generic_kernel(const int* __restrict__ lr, int nlr, const int* __restrict__ lrp, int nlrp,
                      const int* __restrict__ t11, int n_t11, const int* __restrict__ tlrp, int n_tlrp,
                      const int* __restrict__ t22, int n_t22, int* circ4, int* pcl3, int* n_out)
{
    __shared__ int warp_tot[32];
    int base = 0;
    for (int start = 0; start < nlr; start += blockDim.x) {
        const int i = start + threadIdx.x;
        bool flag = false;
        int il = 0, ir = 0, ilp = 0, irp = 0, k = 0;
        if (i < nlr) {
            il = lr[3 * i]; ir = lr[3 * i + 1];
            if (il >= 0 && il < n_t11 && (ilp = t11[il]) >= 0 && ilp < n_tlrp && (k = tlrp[ilp]) >= 0 && k < nlrp) {
                irp = lrp[3 * k + 1];
                flag = ir >= 0 && ir < n_t22 && t22[ir] == irp && irp >= 0;
            }
        }
        const int c = block_compact_slot(flag, base, warp_tot);
        if (flag) {
            circ4[4 * c] = il; circ4[4 * c + 1] = ir; circ4[4 * c + 2] = ilp; circ4[4 * c + 3] = irp;
            pcl3[3 * c] = i; pcl3[3 * c + 1] = k; pcl3[3 * c + 2] = 0;
        }
    }
    if (threadIdx.x == 0) *n_out = base;
}

/* ------------------------------------------------------------------------------------------------ estimation */

struct Rot {
    double r00, r01, r02, r10, r11, r12, r20, r21, r22;
    double rdrx10, rdrx11, rdrx12, rdrx20, rdrx21, rdrx22;
    double rdry00, rdry01, rdry02, rdry10, rdry11, rdry12, rdry20, rdry21, rdry22;
    double rdrz00, rdrz01, rdrz10, rdrz11, rdrz20, rdrz21;
    double tx, ty, tz;
};

/* viso.cpp:1406-1424 */
__device__ __forceinline__ void make_rot(const double* tr, Rot& R, bool derivs)
{
    const double rx = tr[0], ry = tr[1], rz = tr[2];
    R.tx = tr[3]; R.ty = tr[4]; R.tz = tr[5];
    const double sx = sin(rx), cx = cos(rx), sy = sin(ry);
    const double cy = cos(ry), sz = sin(rz), cz = cos(rz);
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

/* Jacobian rows + weighted residuals of one point, viso.cpp:1441-1495 (literal; no shortcuts for the constant
 * derivative columns so that non-finite inputs propagate exactly as in the reference).
 * out: 4 rows x 7 (6 Jacobian columns + residual). */
__device__ __forceinline__ void point_rows(const Rot& R, const ParamDev& P, double X1p, double Y1p, double Z1p,
                                           double weight, const double ob[4], double out[4][7])
{
    const double X1c = R.r00 * X1p + R.r01 * Y1p + R.r02 * Z1p + R.tx;
    const double Y1c = R.r10 * X1p + R.r11 * Y1p + R.r12 * Z1p + R.ty;
    const double Z1c = R.r20 * X1p + R.r21 * Y1p + R.r22 * Z1p + R.tz;
    const double X2c = X1c - P.base;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
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
        out[0][j] = weight * P.f * (X1cd * Z1c - X1c * Z1cd) / (Z1c * Z1c);
        out[1][j] = weight * P.f * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c);
        out[2][j] = weight * P.f * (X1cd * Z1c - X2c * Z1cd) / (Z1c * Z1c);
        out[3][j] = weight * P.f * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c);
    }
    double pred[4];
    pred[0] = P.f * X1c / Z1c + P.cu;
    pred[1] = P.f * Y1c / Z1c + P.cv;
    pred[2] = P.f * X2c / Z1c + P.cu;
    pred[3] = P.f * Y1c / Z1c + P.cv;
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r][6] = weight * (ob[r] - pred[r]);
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

#define VISO_HYP_PER_CTA 32 /* hypotheses per CTA of ransac_hyp_kernel: 4 lanes each */

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
__global__ void __launch_bounds__(VISO_HYP_PER_CTA * 4) ransac_hyp_kernel(const RansacProb* __restrict__ probs, ParamDev P)
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
    if (!bad_index) {
        for (int it = 0; it < 100; ++it) {
            /* make_rot, viso.cpp:1406-1424: lane 0 -> rx, lane 1 -> ry, lane 2 -> rz */
            const double ang = tr[q < 3 ? q : 0];
            const double sv = sin(ang), cv = cos(ang);
            Rot R;
            {
                const double sx = __shfl_sync(qmask, sv, q0), cx = __shfl_sync(qmask, cv, q0);
                const double sy = __shfl_sync(qmask, sv, q0 + 1), cy = __shfl_sync(qmask, cv, q0 + 1);
                const double sz = __shfl_sync(qmask, sv, q0 + 2), cz = __shfl_sync(qmask, cv, q0 + 2);
                R.tx = tr[3]; R.ty = tr[4]; R.tz = tr[5];
                R.r00 = +cy * cz;                R.r01 = -cy * sz;                R.r02 = +sy;
                R.r10 = +sx * sy * cz + cx * sz; R.r11 = -sx * sy * sz + cx * cz; R.r12 = -sx * cy;
                R.r20 = -cx * sy * cz + sx * sz; R.r21 = +cx * sy * sz + sx * cz; R.r22 = +cx * cy;
                R.rdrx10 = +cx * sy * cz - sx * sz; R.rdrx11 = -cx * sy * sz - sx * cz; R.rdrx12 = -cx * cy;
                R.rdrx20 = +sx * sy * cz + cx * sz; R.rdrx21 = -sx * sy * sz + cx * cz; R.rdrx22 = -sx * cy;
                R.rdry00 = -sy * cz;      R.rdry01 = +sy * sz;      R.rdry02 = +cy;
                R.rdry10 = +sx * cy * cz; R.rdry11 = -sx * cy * sz; R.rdry12 = +sx * sy;
                R.rdry20 = -cx * cy * cz; R.rdry21 = +cx * cy * sz; R.rdry22 = -cx * sy;
                R.rdrz00 = -cy * sz;                R.rdrz01 = -cy * cz;
                R.rdrz10 = -sx * sy * sz + cx * cz; R.rdrz11 = -sx * sy * cz - cx * sz;
                R.rdrz20 = +cx * sy * sz + sx * cz; R.rdrz21 = +cx * sy * cz - sx * sz;
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
            if (flag == 2) { ok = 0; break; }
            if (flag == 1) { ok = 1; break; }
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
    }
}

/* One warp per hypothesis: support-set size, viso.cpp:1563 (get_inliers, :1509-1537) */
__global__ void __launch_bounds__(256) ransac_score_kernel(const RansacProb* __restrict__ probs, ParamDev P)
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

/*
 * Block-cooperative minimize_reproj (viso.cpp:1583-1623) over an arbitrary active set.  Per iteration:
 * all threads write Jacobian rows + residuals to `scratch` ([4*na][7]); lanes 0..26 of warp 0 then form the 21
 * JtJ sums and 6 Jt*r sums SEQUENTIALLY in row order (bit-identical to cv::mulTransposed / the oracle's
 * J^T r); thread 0 solves and decides.  tr_s: shared double[6], in/out.  Returns 1 converged / 0 failed.
 */
__device__ int gn_block(const double* __restrict__ X, const double* __restrict__ obs, int stride, int n,
                        const int* __restrict__ active, int na, double* tr_s, const ParamDev& P,
                        double* __restrict__ scratch, double* sums_s /* smem[27] */, int* flag_s /* smem */)
{
    for (int it = 0; it < 100; ++it) {
        double tr[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) tr[j] = tr_s[j];
        Rot R;
        make_rot(tr, R, true);
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
        if (threadIdx.x < 27) {
            const int a = c_pair_a[threadIdx.x], b = c_pair_b[threadIdx.x];
            double s = 0;
            const int rowsN = 4 * na;
#pragma unroll 4
            for (int k = 0; k < rowsN; ++k) s += scratch[(size_t)k * 7 + a] * scratch[(size_t)k * 7 + b];
            sums_s[threadIdx.x] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double A[6][6], b[6];
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = a; c < 6; ++c) { A[a][c] = sums_s[t]; A[c][a] = sums_s[t]; ++t; }
#pragma unroll
            for (int a = 0; a < 6; ++a) b[a] = sums_s[21 + a];
            int flag;
            if (!lu_solve6(A, b)) flag = 2;
            else {
                bool conv = true;
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    if (b[j] > P.thresh) { conv = false; break; }
                if (conv) flag = 1;
                else {
                    flag = 0;
#pragma unroll
                    for (int j = 0; j < 6; ++j) tr_s[j] = tr_s[j] + b[j];
                }
            }
            *flag_s = flag;
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
__global__ void __launch_bounds__(256) ransac_final_kernel(const RansacProb* __restrict__ probs, ParamDev P)
{
    __shared__ int warp_tot[32];
    __shared__ int best_cnt_s[256], best_idx_s[256];
    __shared__ double tr_s[6];
    __shared__ double sums_s[27];
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
        ok = gn_block(pb.X, pb.obs, pb.stride, n, pb.active, n_act, tr_s, P, pb.scratch, sums_s, &flag_s);
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
    __shared__ int flag_s;
    if (threadIdx.x < 6) tr_s[threadIdx.x] = tr[threadIdx.x];
    __syncthreads();
    const int r = gn_block(X, obs, stride, stride, active, na, tr_s, P, scratch, sums_s, &flag_s);
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

cudaError_t viso_launch_pack(const PackJob* jobs, int n_jobs, int max_n, int dlen, int* err_flag, cudaStream_t s)
{
    if (n_jobs <= 0 || max_n <= 0) return cudaSuccess;
    dim3 grid((max_n + 7) / 8, n_jobs);
    pack_desc_kernel<<<grid, 256, 0, s>>>(jobs, dlen, err_flag);
    return cudaGetLastError();
}

cudaError_t viso_launch_extract(const ExtractJob* jobs, int n_jobs, int max_n, int width, int height, int pitch, int radius,
                                cudaStream_t s)
{
    if (n_jobs <= 0 || max_n <= 0) return cudaSuccess;
    dim3 grid((max_n + 8 * VISO_EXTRACT_KPW - 1) / (8 * VISO_EXTRACT_KPW), n_jobs);
    if (radius != 5) return cudaErrorInvalidValue; /* the pipeline's descriptor radius (viso.cpp:1174) */
    extract_desc_kernel<5><<<grid, 256, 0, s>>>(jobs, width, height, pitch);
    return cudaGetLastError();
}

cudaError_t viso_launch_grid(const GridJob* jobs, int n_jobs, GridCfg g, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    const size_t smem = (size_t)(2 * g.gx * g.gy + 1) * sizeof(int);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    grid_build_kernel<<<n_jobs, 512, smem, s>>>(jobs, g);
    return cudaGetLastError();
}

cudaError_t viso_launch_match(const MatchJob* jobs, int n_jobs, int max_nq, int max_nt, const MatchParamsPair& mp,
                              GridCfg g, unsigned long long* sad_pairs, int* n_pending, cudaStream_t s, int* launches)
{
    if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
    static int mode = -1;
    if (mode < 0) {
        const char* m = getenv("VISO_MATCH_MODE");
        mode = (m && m[0] == 'g') ? 1 : 0;
    }
    /* the generic kernel loops over query chunks: about 16 CTAs per SM in total, never more than one per chunk */
    const int gchunks = (max_nq + VISO_STRIP_QPC - 1) / VISO_STRIP_QPC;
    const dim3 ggrid(std::min(gchunks, std::max(1, (148 * 16 + n_jobs - 1) / n_jobs)), n_jobs);
    if (mode == 1) {
        sad_match_generic_kernel<<<ggrid, VISO_MATCH_WARPS * 32, 0, s>>>(jobs, mp, g, sad_pairs, n_pending, 0);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    /* staging capacity for a tile's neighbourhood: twice the expected point count of the grown tile box at the
     * densest target set, within [256, 6144] records of 16 bytes */
    const float r = fmaxf(mp.p[0].radius, mp.p[1].radius);
    const double ext_x = (double)g.gx * VISO_GRID_CS, ext_y = (double)g.gy * VISO_GRID_CS;
    const double bx = fmin(ext_x, VISO_TILE_W * VISO_GRID_CS + 2.0 * (r + 2) + 2 * VISO_GRID_CS);
    const double by = fmin(ext_y, VISO_TILE_H * VISO_GRID_CS + 2.0 * (r + 2) + 2 * VISO_GRID_CS);
    double expect = (double)max_nt * (bx * by) / (ext_x * ext_y);
    if (!(expect >= 0)) expect = 0;
    int cap = (int)fmin(6144.0, fmax(256.0, 2.0 * expect + 64.0));
    cap = (cap + 63) & ~63;
    /* per-query list capacity: twice the expected number of points in the L1 diamond (2 r^2), 64..256 */
    const double in_diamond = (double)max_nt * fmin(1.0, 2.0 * (double)r * r / (ext_x * ext_y));
    int ql_cap = (int)fmin(256.0, fmax(64.0, 2.0 * in_diamond + 16.0));
    ql_cap = (ql_cap + 31) & ~31;
    const size_t smem = (size_t)cap * sizeof(uint4) + (size_t)32 * (ql_cap + 2) * sizeof(unsigned short);
    static int attr_set = 0;
    if (smem > 24 * 1024 && !attr_set) {
        cudaError_t e = cudaFuncSetAttribute(sad_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             6144 * 16 + 32 * 258 * 2);
        if (e != cudaSuccess) return e;
        attr_set = 1;
    }
    cudaError_t e = cudaMemsetAsync(n_pending, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const int tiles = ((g.gx + VISO_TILE_W - 1) / VISO_TILE_W) * ((g.gy + VISO_TILE_H - 1) / VISO_TILE_H);
    sad_match_kernel<<<dim3(tiles, n_jobs), VISO_MATCH_WARPS * 32, smem, s>>>(jobs, mp, g, cap, ql_cap, sad_pairs, n_pending);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sad_match_generic_kernel<<<ggrid, VISO_MATCH_WARPS * 32, 0, s>>>(jobs, mp, g, sad_pairs, n_pending, 1);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t viso_launch_sort(const SortJob* jobs, int n_jobs, int max_n, ParamDev p, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    /* 13 bytes of shared memory per match: (dist, query) record, two u16 position lists, piece flags */
    int cap = max_n < 1 ? 1 : max_n;
    if (cap > 15000) cap = 15000; /* u16 positions and ~200 KB of shared memory */
    cap = (cap + 3) & ~3;
    const size_t smem = (size_t)cap * 13 + 16;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(compact_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    compact_sort_kernel<<<n_jobs, 128, smem, s>>>(jobs, p, cap);
    return cudaGetLastError();
}

cudaError_t viso_launch_circle(const CircleJob* jobs, int n_jobs, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    circle_kernel<<<n_jobs, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}

cudaError_t viso_launch_ransac(const RansacProb* probs, int n_probs, int max_H, int max_n, ParamDev p, cudaStream_t s,
                               int* launches)
{
    if (n_probs <= 0 || max_H <= 0) return cudaSuccess;
    ransac_hyp_kernel<<<dim3((max_H + VISO_HYP_PER_CTA - 1) / VISO_HYP_PER_CTA, n_probs), VISO_HYP_PER_CTA * 4, 0, s>>>(probs, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ransac_score_kernel<<<dim3((max_H + 7) / 8, n_probs), 256, 0, s>>>(probs, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ransac_final_kernel<<<n_probs, 256, 0, s>>>(probs, p);
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

cudaError_t viso_launch_circle_tables(const int* m, int n, int* table, int table_n, int key_col, int val_mode,
                                      int* err_flag, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    circle_tables_kernel<<<(n + 255) / 256, 256, 0, s>>>(m, n, table, table_n, key_col, val_mode, err_flag);
    return cudaGetLastError();
}

cudaError_t viso_launch_circle_generic(const int* lr, int nlr, const int* lrp, int nlrp, const int* t11, int n_t11,
                                       const int* tlrp, int n_tlrp, const int* t22, int n_t22,
                                       int* circ4, int* pcl3, int* n_out, cudaStream_t s)
{
    circle_generic_kernel<<<1, 256, 0, s>>>(lr, nlr, lrp, nlrp, t11, n_t11, tlrp, n_tlrp, t22, n_t22, circ4, pcl3, n_out);
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
