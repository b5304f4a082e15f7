/*
 * match.cu -- descriptor packing / extraction, candidate grid and the SAD matching kernels (match_desc,
 * viso.cpp:668-722; MyFeatureExtractor, viso.cpp:1004-1024) with their launch wrappers.
 */
#include "viso_dev.h"
#include "common.cuh"

#include <stdlib.h>
#include <algorithm>

/* ------------------------------------------------------------------------------------------------ pack */

/*
 * f32 cv::Mat descriptor rows (viso.cpp:999-1002) -> biased u16 rows (v + 1024, pad elements 0) in CELL-SORTED order
 * (row p is the descriptor of keypoint srec[p].z), plus the u32 sum of the packed row into srec[p].w.  Runs after
 * grid_build_kernel.  One warp per row, lane l owns elements 4l..4l+3.  Also the domain check (integer valued,
 * |v| <= 1023).
 */
__global__ void __launch_bounds__(256) pack_desc_kernel(const PackJob* __restrict__ jobs, int dlen, int* err)
{
    const PackJob job = jobs[blockIdx.y];
    if (job.from_image && *job.from_image) return;
    const int n = *job.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + warp; row < n; row += gridDim.x * 8) {
        const float* src = job.d + (size_t)job.srec[row].z * dlen;
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
        if (lane == 0) job.srec[row].w = tot;
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
 * Half a warp per keypoint, lane = column of the (2R+3)-pixel-wide source window: the lane loads its column (one
 * byte per window row; a row of the window is one contiguous run for the half warp), forms the vertical [1 2 1] sums
 * in registers, and the horizontal difference s(x+1) - s(x-1) comes from the lane two to the right by shuffle.
 * The 121 values go through a 256-byte shared-memory row to reach the packed layout (16 bytes per lane) and the row
 * sum.  In the pipeline the kernel runs after grid_build_kernel and walks the CELL-SORTED records: row p of the output
 * is the descriptor at srec[p].xy and its sum lands in srec[p].w (job.srec == null: rows follow job.kp, sums to rsum).  The image is touched once per frame, so the loads mostly miss to DRAM: a warp issues the loads of
 * VISO_EXTRACT_ROUNDS x 2 keypoints before using any.
 */
#ifndef VISO_EXTRACT_ITERS
#define VISO_EXTRACT_ITERS 16 /* iterations per warp: amortises the job set-up of a CTA (1: 0.79 ms, 8 and more: 0.66 ms) */
#endif
#ifndef VISO_EXTRACT_ROUNDS
#define VISO_EXTRACT_ROUNDS 2
#endif

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1); /* far outside: any in-range pixel, the sample is masked to 0 anyway */
}

template <int radius>
#ifndef VISO_EXTRACT_MINB
#define VISO_EXTRACT_MINB 8 /* 32 registers, 64 warps per SM: the kernel waits on image lines from DRAM */
#endif
__global__ void __launch_bounds__(256, VISO_EXTRACT_MINB) extract_desc_kernel(const ExtractJob* __restrict__ jobs, int width, int height, int pitch)
{
    constexpr int side = 2 * radius + 1, dlen = side * side, wside = side + 2;
    static_assert(wside <= 16 && dlen <= VISO_DESC_U16, "half a warp per keypoint: radius <= 6");
    /* a staging row: the 128 packed elements, then a trash area for the lanes beyond the window (columns side..15 store
     * there instead of branching around the store: element (r, col) of such a lane lands at 128 + (col - side) + r * side) */
    constexpr int SROW = VISO_DESC_U16 + (16 - side) + (side - 1) * side + 1 + 7 & ~7;
    __shared__ __align__(16) unsigned short out_s[8][VISO_EXTRACT_ROUNDS][2][SROW];
    const ExtractJob job = jobs[blockIdx.y];
    if (!*job.from_image) return;
    const int n = *job.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, col = lane & 15;
    constexpr int KPW = 2 * VISO_EXTRACT_ROUNDS;
    /* the pad elements dlen..127 of the staging rows are zero and stay zero (nothing else ever writes them) */
    for (int i = lane; i < VISO_EXTRACT_ROUNDS * 2 * (VISO_DESC_U16 - dlen); i += 32)
        (&out_s[warp][0][0][0])[(i / (VISO_DESC_U16 - dlen)) * SROW + dlen + i % (VISO_DESC_U16 - dlen)] = 0;
    __syncwarp();
    const int cx = min(col, wside - 1);                                   /* lanes beyond the window repeat its last column */
    const int ocol = col < side ? col : VISO_DESC_U16 + (col - side);      /* where this lane's elements go in a staging row */
    for (int row0 = (blockIdx.x * 8 + warp) * KPW; row0 < n; row0 += gridDim.x * 8 * KPW) {
        int px[VISO_EXTRACT_ROUNDS], py[VISO_EXTRACT_ROUNDS];
        bool inside = true; /* every window of this warp's keypoints lies inside the image */
#pragma unroll
        for (int q = 0; q < VISO_EXTRACT_ROUNDS; ++q) {
            const int kr = min(row0 + 2 * q + half, n - 1);
            float2 kp;
            if (job.srec) { const uint4 sr = job.srec[kr]; kp = make_float2(__uint_as_float(sr.x), __uint_as_float(sr.y)); }
            else kp = job.kp[kr];
            px[q] = __float2int_rn(kp.x); py[q] = __float2int_rn(kp.y);
            inside = inside && px[q] > radius && px[q] + radius + 1 < width && py[q] > radius && py[q] + radius + 1 < height;
        }
        unsigned char p[VISO_EXTRACT_ROUNDS][wside];
        if (__all_sync(FULL, inside)) {
            /* the common case, warp uniform: no reflection, no masked sample, no branch in the element loop */
#pragma unroll
            for (int q = 0; q < VISO_EXTRACT_ROUNDS; ++q) {
                const unsigned char* base = job.img + (size_t)(py[q] - radius - 1) * pitch + (px[q] - radius - 1 + cx);
#pragma unroll
                for (int r = 0; r < wside; ++r) p[q][r] = __ldg(base + r * pitch);
            }
#pragma unroll
            for (int q = 0; q < VISO_EXTRACT_ROUNDS; ++q) {
                unsigned short* orow = out_s[warp][q][half] + ocol;
#pragma unroll
                for (int r = 0; r < side; ++r) {
                    const int sv = (int)p[q][r] + 2 * (int)p[q][r + 1] + (int)p[q][r + 2]; /* column cx, image rows y-1..y+1 */
                    const int sob = __shfl_down_sync(FULL, sv, 2, 16) - sv;                 /* s(x+1) - s(x-1), x = px-R+col */
                    orow[r * side] = (unsigned short)(sob + 1024);
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < VISO_EXTRACT_ROUNDS; ++q) {
                const int ix = reflect101(px[q] - radius - 1 + cx, width);
#pragma unroll
                for (int r = 0; r < wside; ++r)
                    p[q][r] = __ldg(job.img + (size_t)reflect101(py[q] - radius - 1 + r, height) * pitch + ix);
            }
#pragma unroll
            for (int q = 0; q < VISO_EXTRACT_ROUNDS; ++q) {
                unsigned short* orow = out_s[warp][q][half] + ocol;
#pragma unroll
                for (int r = 0; r < side; ++r) {
                    const int sv = (int)p[q][r] + 2 * (int)p[q][r + 1] + (int)p[q][r + 2];
                    int sob = __shfl_down_sync(FULL, sv, 2, 16) - sv;
                    const int y = py[q] + r - radius, x = px[q] + col - radius;
                    if (!(y > 0 && y < height && x > 0 && x < width)) sob = 0;               /* viso.cpp:1018 */
                    orow[r * side] = (unsigned short)(sob + 1024);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < VISO_EXTRACT_ROUNDS; ++q) {
            const int row = row0 + 2 * q + half;
            const uint4 w = reinterpret_cast<const uint4*>(out_s[warp][q][half])[col];
            /* sum of the eight u16 of this lane: dp2a with ones adds the two halves of a word to the accumulator */
            unsigned sum = __dp2a_lo(w.x, 0x0101u, __dp2a_lo(w.y, 0x0101u, __dp2a_lo(w.z, 0x0101u, __dp2a_lo(w.w, 0x0101u, 0u))));
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o, 16);
            if (row < n) {
                reinterpret_cast<uint4*>(job.out + (size_t)row * VISO_DESC_U16)[col] = w;
                if (col == 0) {
                    if (job.srec) job.srec[row].w = sum;
                    else job.rsum[row] = sum;
                }
            }
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------------ grid */

/*
 * Counting sort of one keypoint set into 16-px cells (one CTA per set).  Emits, in cell order, the candidate
 * records the matcher streams: srec = (x, y, original index, row sum -- filled in by the pack / extract kernel that
 * follows), and the inverse permutation pos_of.
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
        job.srec[pos] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), (unsigned)i, 0u);
        job.pos_of[i] = pos;
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
            int p = 0;
            if (in) {
                while (fl >= ws.rowPre[row + 1]) ++row;
                p = ws.rowS0[row] + (fl - ws.rowPre[row]);
                rec = __ldg(t.srec + p);
                dist = l1_dist(q.qx, q.qy, __uint_as_float(rec.x), __uint_as_float(rec.y));
            }
            f(in, dist, rec, p);
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
struct ListAcc { /* the warp's uint4 list (generic kernel): (target index, L1 distance bits, row sum, sorted position) */
    const WarpScratch& ws;
    __device__ __forceinline__ uint4 entry(int e) const { return ws.list[e]; }
    __device__ __forceinline__ unsigned pos(const uint4& ent) const { return ent.w; }
};
struct TileAcc { /* region indices into the staged neighbourhood (tile kernel): the distance is recomputed */
    const unsigned short* ql;
    const uint4* reg;
    const int* regpos;   /* sorted position (= descriptor row) of every staged record */
    float qx, qy;
    __device__ __forceinline__ unsigned pos(const uint4& ent) const { return ent.w; }
    __device__ __forceinline__ uint4 entry(int e) const
    {
        const unsigned ri = ql[e];
        const uint4 r = reg[ri];
        return make_uint4(r.z, __float_as_uint(l1_dist(qx, qy, __uint_as_float(r.x), __uint_as_float(r.y))), r.w, (unsigned)regpos[ri]);
    }
};

/* a * one + c as one IMAD: inline PTX so that the front end cannot factor the multiplier out of a chain of them */
__device__ __forceinline__ unsigned imad1(unsigned a, unsigned one, unsigned c)
{
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
}

/* NB = butterfly width in steps (8: up to 32 candidates, 4: up to 16), NL <= NB = steps whose rows are actually
 * loaded and evaluated (the rest contribute 0 and are masked) */
template <int NB, int NL, class Acc>
__device__ __forceinline__ void eval_batch(const uint4* __restrict__ tbase, const Acc& acc_, int base, int n, int lane,
                                           const uint4& qa, const uint4& qb, unsigned qsum, BestState& st, unsigned one)
{
    const int sub = lane & 7, g = lane >> 3;
    const bool b0 = sub & 1, b1 = sub & 2, b2 = sub & 4;
    /* lane L fetches the record of candidate base + L once; the steps and the final fold get it by shuffle */
    const uint4 ent = acc_.entry(min(base + lane, n - 1));
    const unsigned epos = acc_.pos(ent); /* sorted position = row of the candidate's descriptor */
    unsigned part[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) part[s] = 0;
#pragma unroll
    for (int h = 0; h < NL; h += VISO_EVAL_DEPTH) { /* VISO_EVAL_DEPTH steps = 2*VISO_EVAL_DEPTH row loads in flight */
        uint4 ra[VISO_EVAL_DEPTH], rb[VISO_EVAL_DEPTH];
#pragma unroll
        for (int s = 0; s < VISO_EVAL_DEPTH; ++s) {
            if (h + s < NL) {
                const unsigned rowp = __shfl_sync(FULL, epos, 4 * (h + s) + g);
                const uint4* rp = tbase + (size_t)rowp * (VISO_DESC_U16 / 8);
                ra[s] = __ldg(rp);
                rb[s] = __ldg(rp + 8);
            }
        }
#pragma unroll
        for (int s = 0; s < VISO_EVAL_DEPTH; ++s) {
            if (h + s < NL) {
                /* min on the ALU pipe, the running sum on the FMA pipe (IMAD with the opaque multiplier `one`): with
                 * 3-input adds everything went down the ALU pipe, which was the kernel's limiter (IPC 2.45 of 4).
                 * Packed u16x2 sums: 16 elements x 2047 per half < 65536, no carry between halves */
                unsigned a0 = __vminu2(qa.x, ra[s].x), a1 = __vminu2(qb.x, rb[s].x);
                a0 = imad1(__vminu2(qa.y, ra[s].y), one, a0); a1 = imad1(__vminu2(qb.y, rb[s].y), one, a1);
                a0 = imad1(__vminu2(qa.z, ra[s].z), one, a0); a1 = imad1(__vminu2(qb.z, rb[s].z), one, a1);
                a0 = imad1(__vminu2(qa.w, ra[s].w), one, a0); a1 = imad1(__vminu2(qb.w, rb[s].w), one, a1);
                part[h + s] = imad1(a1, one, a0);
            }
        }
    }
    /* transposed reduction over the 8 lanes of a row group: lane (g, sub) ends with candidate 4*step + g where
     * step = sub (NB = 8) or sub & 3 (NB = 4; lanes sub and sub^4 then hold the same candidate) */
    /* the first stage adds two packed sums (2 x 32752 per half still fits); the halves are folded after it */
    unsigned r2[2];
    if (NB == 8) {
        unsigned r4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned lo = part[2 * j], hi = part[2 * j + 1];
            const unsigned t = imad1(__shfl_xor_sync(FULL, b0 ? lo : hi, 1), one, b0 ? hi : lo);
            r4[j] = imad1(t >> 16, one, t & 0xffffu);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned lo = r4[2 * j], hi = r4[2 * j + 1];
            r2[j] = imad1(__shfl_xor_sync(FULL, b1 ? lo : hi, 2), one, b1 ? hi : lo);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const unsigned lo = part[2 * j], hi = part[2 * j + 1];
            const unsigned t = imad1(__shfl_xor_sync(FULL, b0 ? lo : hi, 1), one, b0 ? hi : lo);
            r2[j] = imad1(t >> 16, one, t & 0xffffu);
        }
    }
    unsigned tot;
    int slot; /* position of this lane's candidate inside the batch */
    bool mine = true;
    if (NB == 8) {
        tot = imad1(__shfl_xor_sync(FULL, b2 ? r2[0] : r2[1], 4), one, b2 ? r2[1] : r2[0]);
        slot = 4 * sub + g;
    } else {
        const unsigned t2 = imad1(__shfl_xor_sync(FULL, b1 ? r2[0] : r2[1], 2), one, b1 ? r2[1] : r2[0]);
        tot = imad1(__shfl_xor_sync(FULL, t2, 4), one, t2);
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
                                          const uint4& qa, const uint4& qb, unsigned qsum, BestState& st, unsigned one)
{
    const uint4* tbase = reinterpret_cast<const uint4*>(tdesc) + (lane & 7);
    int base = 0;
    for (; n - base > 24; base += 32) eval_batch<8, 8>(tbase, acc, base, n, lane, qa, qb, qsum, st, one);
    const int rest = n - base; /* 0..24: rows are loaded in units of 8 candidates */
    if (rest > 16) eval_batch<8, 6>(tbase, acc, base, n, lane, qa, qb, qsum, st, one);
    else if (rest > 8) eval_batch<4, 4>(tbase, acc, base, n, lane, qa, qb, qsum, st, one);
    else if (rest > 0) eval_batch<4, 2>(tbase, acc, base, n, lane, qa, qb, qsum, st, one);
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
                                                int lane, const uint4 qrec, unsigned one)
{
    const int q = (int)qrec.z;
    const float qx = __uint_as_float(qrec.x), qy = __uint_as_float(qrec.y);
    const unsigned qsum = qrec.w;
    const uint4* qp = reinterpret_cast<const uint4*>(job.q.desc + (size_t)__ldg(job.q.pos_of + q) * VISO_DESC_U16) + (lane & 7);
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
        vis.all([&](bool in, float dist, uint4 rec, int) {
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
                    vis.all([&](bool in, float dist, uint4 rec, int) {
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
                        vis.all([&](bool in, float dist, uint4 rec, int) {
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
    vis.all([&](bool in, float dist, uint4 rec, int p) {
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
        if (take) ws.list[nlist + __popc(tm & ((1u << lane) - 1))] = make_uint4(rec.z, __float_as_uint(dist), rec.w, (unsigned)p);
        nlist += __popc(tm);
        if (nlist > VISO_LIST_CAP - 32) {
            __syncwarp();
            eval_list(job.t.desc, ListAcc{ws}, nlist, lane, qa, qb, qsum, st, one);
            pairs += nlist;
            nlist = 0;
            __syncwarp();
        }
    });
    if (nlist > 0) {
        __syncwarp();
        eval_list(job.t.desc, ListAcc{ws}, nlist, lane, qa, qb, qsum, st, one);
        pairs += nlist;
        __syncwarp();
    }

    if (lane == 0) write_result(job, P, q, st);
    return pairs;
}

__device__ __forceinline__ void pend_push(const PendingList& pend, const uint4& qrec, int job_index)
{
    const int slot = atomicAdd(pend.count, 1);
    if (slot < pend.cap) { pend.rec[slot] = qrec; pend.job[slot] = job_index; }
}

/*
 * sad_match_kernel: match_desc (viso.cpp:668-722) for a batch of jobs.  blockIdx.y = job, blockIdx.x = query tile
 * (TW x TH cells of the QUERY set's grid: 6 x 4 = 96 x 64 px).
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
template <int TW, int TH>
__global__ void __launch_bounds__(VISO_MATCH_WARPS * 32, VISO_MATCH_MINB)
sad_match_kernel(const MatchJob* __restrict__ jobs, MatchParamsPair mp, GridCfg g, int reg_cap, int ql_cap,
                 unsigned long long* sad_pairs, PendingList pend)
{
    extern __shared__ uint4 reg[];                                /* staged neighbourhood, reg_cap records */
    /* per-query candidate lists (region indices), 32 x (ql_cap + 2): the +2 makes the word stride odd so that
     * lanes = queries write without bank conflicts */
    int* const regpos = reinterpret_cast<int*>(reg + reg_cap);    /* sorted position of every staged record */
    unsigned short* const qlist = reinterpret_cast<unsigned short*>(regpos + reg_cap);
    const int ql_stride = ql_cap + 2;
    __shared__ int qcnt[32];
    __shared__ int row_off[VISO_MAX_REG_ROWS + 1];
    __shared__ int row_p0[VISO_MAX_REG_ROWS];
    __shared__ int tile_s[4];
    __shared__ uint4 qrec_s[32];

    const MatchJob job = jobs[blockIdx.y];
    const MatchParamsDev& P = mp.p[job.mode];
    const int nq = *job.q.n, nt = *job.t.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (g.gx + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    if (ty * TH >= g.gy || nq <= 0) return;

    /* the tile's queries: one span of the cell-sorted query array per cell row */
    const int cx_lo = tx * TW, cx_hi = min(cx_lo + TW, g.gx);
    int qtot = 0;
    int qs[TH], ql[TH];
#pragma unroll
    for (int rr = 0; rr < TH; ++rr) {
        const int cy = ty * TH + rr;
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
        for (int rr = 0; rr < TH; ++rr) {
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
        const int cy0 = cell_coord(ymin - grow, g.gy), cy1 = cell_coord(ymax + grow, g.gy);
        const int nr = cy1 - cy0 + 1;
        int total = -1; /* -1: no staging */
        if (nr <= VISO_MAX_REG_ROWS && xmin == xmin && ymin == ymin && reg_cap > 0) {
            int run = 0;
            for (int b = 0; b < nr; b += 32) {
                const int rr = b + lane;
                int len = 0, p0 = 0;
                if (rr < nr) {
                    /* the reach that is left at this grid row's vertical distance from the box: the corners of the
                     * bounding box are out of every query's L1 diamond and are not staged */
                    const int cy = cy0 + rr;
                    const float lo = (cy == 0) ? -CUDART_INF_F : (float)(cy * VISO_GRID_CS);
                    const float hi = (cy == g.gy - 1) ? CUDART_INF_F : (float)((cy + 1) * VISO_GRID_CS);
                    const float dymin = fmaxf(0.f, fmaxf(lo - ymax, ymin - hi));
                    const float rem = grow - dymin;
                    if (rem >= 0.f) {
                        const int cx0 = cell_coord(xmin - rem, g.gx), cx1 = cell_coord(xmax + rem, g.gx);
                        p0 = __ldg(job.t.cell_start + cy * g.gx + cx0);
                        len = __ldg(job.t.cell_start + cy * g.gx + cx1 + 1) - p0;
                    }
                }
                const int incl = warp_incl_scan(len, lane);
                if (rr < nr) { row_off[rr] = run + incl - len; row_p0[rr] = p0; }
                run += __shfl_sync(FULL, incl, 31);
            }
            if (lane == 0) row_off[nr] = run;
            if (run <= reg_cap) total = run;
        }
        if (lane == 0) { tile_s[0] = 0; tile_s[1] = cy0; tile_s[2] = nr; tile_s[3] = total; }
    }
    __syncthreads();
    const int nrows = tile_s[2];
    const bool tile_ok = tile_s[3] >= 0;

    unsigned pairs = 0;
    if (!tile_ok) { /* leave the whole tile to the generic kernel */
        for (int k = threadIdx.x; k < qtot; k += blockDim.x) {
            const uint4 qr = query_rec(k);
            job.out[qr.z] = make_int4(0, 0, 0, VISO_PENDING);
            pend_push(pend, qr, blockIdx.y);
        }
    } else {
        const int R = tile_s[3];
        for (int rr = warp; rr < nrows; rr += VISO_MATCH_WARPS) {
            const int o = row_off[rr], len = row_off[rr + 1] - o;
            const int p0 = row_p0[rr];
            const uint4* src = job.t.srec + p0;
            for (int i = lane; i < len; i += 32) { reg[o + i] = __ldg(src + i); regpos[o + i] = p0 + i; }
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
                /* candidates need dist <= r and dist < D0 (index-0 rule), D0 = d0 if d0 <= r: one strict comparison
                 * against lim = min(next float above r, D0); lanes without a query get lim = -inf */
                /* the next float above r (r >= 0; an infinite radius stays infinite; a negative one only moves
                 * further below every distance) */
                const float r_up = r < CUDART_INF_F ? __uint_as_float(__float_as_uint(r) + 1u) : r;
                const float lim = act ? ((d0 <= r) ? d0 : r_up) : -CUDART_INF_F;
                const float2* pts = reinterpret_cast<const float2*>(reg);
#pragma unroll 4
                for (int i = warp; i < R; i += VISO_MATCH_WARPS) {
                    const float2 p = pts[2 * i]; /* (x, y) of the uint4 record: broadcast */
                    const float dist = l1_dist(qx, qy, p.x, p.y);
                    if (dist < lim) {
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
                        pend_push(pend, qrec, blockIdx.y);
                    }
                    continue;
                }
                const int q = (int)qrec.z;
                const float qx = __uint_as_float(qrec.x), qy = __uint_as_float(qrec.y);
                const uint4* qp = reinterpret_cast<const uint4*>(job.q.desc + (size_t)__ldg(job.q.pos_of + q) * VISO_DESC_U16) + (lane & 7);
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
                if (nlist > 0) eval_list(job.t.desc, TileAcc{ql, reg, regpos, qx, qy}, nlist, lane, qa, qb, qrec.w, st, g.one);
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
#ifndef VISO_GENERIC_MINB
#define VISO_GENERIC_MINB 6 /* 80 registers: BASELINE configs[2] 0.212 / 0.309 ms (at the 96 the compiler picks unprompted) -> 0.199 / 0.271 ms */
#endif
__global__ void __launch_bounds__(VISO_MATCH_WARPS * 32, VISO_GENERIC_MINB)
sad_match_generic_kernel(const MatchJob* __restrict__ jobs, MatchParamsPair mp, GridCfg g, unsigned long long* sad_pairs,
                         PendingList pend, int only_pending)
{
    __shared__ WarpScratch wscr[VISO_MATCH_WARPS];
    const int np = only_pending ? *pend.count : 0;
    if (only_pending && np == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch& ws = wscr[warp];
    unsigned pairs = 0;
    auto one_query = [&](const MatchJob& job, const uint4& qrec) {
        const MatchParamsDev& P = mp.p[job.mode];
        if (*job.t.n <= 0) {
            if (lane == 0) job.out[qrec.z] = make_int4(-1, INT_MAX, INT_MAX, 0);
            return;
        }
        GlobalVisitor vis{job.t, g, make_geom(g, __uint_as_float(qrec.x), __uint_as_float(qrec.y), P.radius), ws, lane, 0, true};
        pairs += match_query(vis, job, P, ws, lane, qrec, g.one);
    };
    if (only_pending && pend.rec && np <= pend.cap) {
        /* the listed queries, spread over all CTAs of the grid (entry e goes to CTA e mod #CTAs) */
        const int cta = blockIdx.y * gridDim.x + blockIdx.x, ncta = gridDim.x * gridDim.y;
        for (int e = cta + warp * ncta; e < np; e += ncta * VISO_MATCH_WARPS) one_query(jobs[pend.job[e]], pend.rec[e]);
    } else {
        const MatchJob job = jobs[blockIdx.y];
        const int nq = *job.q.n;
        for (int chunk = blockIdx.x; chunk * VISO_STRIP_QPC < nq; chunk += gridDim.x)
            for (int qi = chunk * VISO_STRIP_QPC + warp; qi < min(nq, (chunk + 1) * VISO_STRIP_QPC); qi += VISO_MATCH_WARPS) {
                const uint4 qrec = __ldg(job.q.srec + qi);
                if (only_pending && job.out[qrec.z].w != VISO_PENDING) continue;
                one_query(job, qrec);
            }
    }
    if (sad_pairs && lane == 0 && pairs) {
        atomicAdd(sad_pairs, (unsigned long long)pairs);
        atomicAdd(sad_pairs + 1, (unsigned long long)pairs);
    }
}

/* ------------------------------------------------------------------------------------------------ staged tile kernel */

/* 1-D bulk async copies (TMA engine, SASS UBLKCP) completing on an mbarrier */
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    }
}

#define VISO_ST_WARPS 8
#define VISO_ST_SLOTS 64          /* queries of a tile handled per round */
#define VISO_ST_MAX_TH 8          /* tile height in cells at most */
#define VISO_ST_MAX_ROWS 32       /* grid rows a staged neighbourhood may span (one lane of warp 0 each) */

struct StagedCfg {
    int tw, th;       /* query tile in cells */
    int cap_rows;     /* staging capacity in descriptor rows / candidate records */
    int ql_cap;       /* per-query candidate list capacity */
};

/*
 * sad_match_staged_kernel: match_desc (viso.cpp:668-722) for a batch of jobs with the descriptor rows of a tile's
 * neighbourhood STAGED IN SHARED MEMORY.  blockIdx.y = job, blockIdx.x = query tile (cfg.tw x cfg.th cells).
 *
 * Descriptor rows live in HBM in cell-sorted order, so the target points that can be within `radius` of any query of
 * the tile are one contiguous span of rows (and of candidate records) per grid row.  Warp 0 works out those spans --
 * the bounding box of the tile's query coordinates grown by radius + slack, each grid row trimmed to the reach that
 * is left at its vertical distance -- and its lanes issue one cp.async.bulk per span for the 16-byte records and one
 * for the 256-byte rows, each set completing on its own mbarrier.  No thread touches the data on its way in.
 *
 * 1. (records arrived) Candidate generation, LANE = QUERY: every warp walks a share of the staged records; the point is
 *    a shared-memory broadcast and each lane tests it against its own query (radius and index-0 terminator, viso.cpp:
 *    693), appending hits to that query's list.
 * 2. (rows arrived) Evaluation, WARP = QUERY, LANE = CANDIDATE: each lane computes the complete 121-element SAD of its
 *    own candidate against the query -- no cross-lane reduction.  The query row sits in registers, 16 chunks of 16
 *    bytes; lane l walks the chunks in the order s ^ (l & 7), so the eight lanes of a quarter warp always read eight
 *    different 16-byte bank groups whatever rows they are on (rows are 256 bytes apart: without the rotation all
 *    lanes would hit the same four banks).  Per element pair one VIMNMX.U16x2 + one add: sum|a-b| = sum a + sum b -
 *    2 sum min(a,b), packed u16 partial sums of 32 words never carry.  Each lane keeps its own (best, second best,
 *    tie-break key); one REDUX fold per query at the end.  Ties on the SAD go to the largest (L1, index) key = the
 *    last in the reference's scan order (viso.cpp:703).
 * Queries that need the top-K cut or overflow their list, and tiles whose neighbourhood does not fit the staging
 * buffers, are left to sad_match_generic_kernel through the pending list, exactly like the gather tile kernel.
 */
__global__ void __launch_bounds__(VISO_ST_WARPS * 32, 2)
sad_match_staged_kernel(const MatchJob* __restrict__ jobs, MatchParamsPair mp, GridCfg g, StagedCfg cfg,
                        unsigned long long* sad_pairs, PendingList pend)
{
    extern __shared__ __align__(128) unsigned char st_smem[];
    unsigned char* const rows = st_smem;                                                /* cap_rows x 256 B */
    uint4* const reg = reinterpret_cast<uint4*>(st_smem + (size_t)cfg.cap_rows * 256);  /* cap_rows records */
    unsigned short* const qlist = reinterpret_cast<unsigned short*>(reg + cfg.cap_rows);
    const int ql_cap = cfg.ql_cap, ql_stride = ql_cap + 2; /* odd word stride: lanes = queries write conflict free */
    __shared__ int qcnt[VISO_ST_SLOTS];
    __shared__ uint4 qrec_s[VISO_ST_SLOTS];
    __shared__ int qpos_s[VISO_ST_SLOTS];
    __shared__ int tile_s[2];
    __shared__ __align__(8) unsigned long long mbar[2];

    const MatchJob job = jobs[blockIdx.y];
    const MatchParamsDev& P = mp.p[job.mode];
    const int nq = *job.q.n, nt = *job.t.n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_x = (g.gx + cfg.tw - 1) / cfg.tw;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    if (ty * cfg.th >= g.gy || nq <= 0) return;

    /* the tile's queries: one span of the cell-sorted query array per cell row */
    const int cx_lo = tx * cfg.tw, cx_hi = min(cx_lo + cfg.tw, g.gx);
    int qtot = 0;
    int qs[VISO_ST_MAX_TH], ql[VISO_ST_MAX_TH];
#pragma unroll
    for (int rr = 0; rr < VISO_ST_MAX_TH; ++rr) {
        const int cy = ty * cfg.th + rr;
        qs[rr] = 0; ql[rr] = 0;
        if (rr < cfg.th && cy < g.gy) {
            qs[rr] = __ldg(job.q.cell_start + cy * g.gx + cx_lo);
            ql[rr] = __ldg(job.q.cell_start + cy * g.gx + cx_hi) - qs[rr];
        }
        qtot += ql[rr];
    }
    if (qtot == 0) return;

    auto query_pos = [&](int k) {
        int pos = 0;
#pragma unroll
        for (int rr = 0; rr < VISO_ST_MAX_TH; ++rr) {
            if (k >= 0 && k < ql[rr]) pos = qs[rr] + k;
            k -= ql[rr];
        }
        return pos;
    };

    if (nt <= 0) { /* no targets: every query is unmatched */
        for (int k = threadIdx.x; k < qtot; k += blockDim.x)
            job.out[__ldg(job.q.srec + query_pos(k)).z] = make_int4(-1, INT_MAX, INT_MAX, 0);
        return;
    }

    const float r = P.radius;
    if (warp == 0) {
        if (lane == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            mbar_fence_init();
        }
        /* bounding box of the tile's query coordinates (queries outside the image extent are clamped into border
         * cells, so the cell rectangle is not a bound) */
        float xmin = CUDART_INF_F, xmax = -CUDART_INF_F, ymin = CUDART_INF_F, ymax = -CUDART_INF_F, amax = 0.f;
        for (int k = lane; k < qtot; k += 32) {
            const uint4 qr = __ldg(job.q.srec + query_pos(k));
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
        /* reach: radius + the largest per-query slack (make_geom) + a margin far above the float rounding of the sums */
        const float grow = r + (1.0f + 4e-6f * (amax + r)) + 4e-6f * (amax + r) + 1e-3f;
        const int cy0 = cell_coord(ymin - grow, g.gy), cy1 = cell_coord(ymax + grow, g.gy);
        const int nr = cy1 - cy0 + 1;
        int total = -1; /* -1: no staging */
        if (nr <= VISO_ST_MAX_ROWS && xmin == xmin && ymin == ymin) {
            /* lane = grid row of the neighbourhood: the reach left at the row's vertical distance from the box */
            int start = 0, len = 0;
            if (lane < nr) {
                const int cy = cy0 + lane;
                const float lo = (cy == 0) ? -CUDART_INF_F : (float)(cy * VISO_GRID_CS);
                const float hi = (cy == g.gy - 1) ? CUDART_INF_F : (float)((cy + 1) * VISO_GRID_CS);
                const float dymin = fmaxf(0.f, fmaxf(lo - ymax, ymin - hi));
                const float rem = grow - dymin;
                if (rem >= 0.f) {
                    const int cx0 = cell_coord(xmin - rem, g.gx), cx1 = cell_coord(xmax + rem, g.gx);
                    start = __ldg(job.t.cell_start + cy * g.gx + cx0);
                    len = __ldg(job.t.cell_start + cy * g.gx + cx1 + 1) - start;
                }
            }
            const int incl = warp_incl_scan(len, lane);
            const int run = __shfl_sync(FULL, incl, 31);
            if (run <= cfg.cap_rows) {
                total = run;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&mbar[0], (unsigned)run * 16u);
                    mbar_arrive_expect_tx(&mbar[1], (unsigned)run * 256u);
                }
                if (len > 0) {
                    const int off = incl - len;
                    bulk_g2s(reg + off, job.t.srec + start, (unsigned)len * 16u, &mbar[0]);
                    bulk_g2s(rows + (size_t)off * 256, job.t.desc + (size_t)start * VISO_DESC_U16, (unsigned)len * 256u, &mbar[1]);
                }
            }
        }
        if (lane == 0) tile_s[0] = total;
    }
    __syncthreads();
    const int R = tile_s[0];

    unsigned pairs = 0;
    if (R < 0) { /* leave the whole tile to the generic kernel */
        for (int k = threadIdx.x; k < qtot; k += blockDim.x) {
            const uint4 qr = __ldg(job.q.srec + query_pos(k));
            job.out[qr.z] = make_int4(0, 0, 0, VISO_PENDING);
            pend_push(pend, qr, blockIdx.y);
        }
        return;
    }
    const float2 t0 = __ldg(job.t.xy);
    const unsigned rows_u32 = smem_u32(rows);
    const int rot = lane & 7;
    for (int g0 = 0; g0 < qtot; g0 += VISO_ST_SLOTS) {
        const int nslot = min(VISO_ST_SLOTS, qtot - g0);
        if (g0 > 0) __syncthreads(); /* the previous round's lists are consumed */
        if (threadIdx.x < VISO_ST_SLOTS) qcnt[threadIdx.x] = 0;
        __syncthreads();
        if (g0 == 0) mbar_wait(&mbar[0], 0); /* candidate records have landed */
        {
            /* lane = query slot.  Up to 32 queries: all warps share slots 0..31; else warps 0-3 take slots 0..31 and
             * warps 4-7 slots 32..63; a warp walks every stride-th staged point */
            const bool two = nslot > 32;
            const int half = two ? (warp >> 2) : 0, part = two ? (warp & 3) : warp, stride = two ? 4 : 8;
            const int slot = half * 32 + lane;
            const bool act = slot < nslot;
            const int qpos = query_pos(g0 + (act ? slot : 0));
            const uint4 qr = __ldg(job.q.srec + qpos);
            if (part == 0 && act) { qrec_s[slot] = qr; qpos_s[slot] = qpos; }
            const float qx = __uint_as_float(qr.x), qy = __uint_as_float(qr.y);
            const float d0 = l1_dist(qx, qy, t0.x, t0.y);
            /* candidates need dist <= r and dist < D0 (index-0 rule), D0 = d0 if d0 <= r: one strict comparison
             * against lim = min(next float above r, D0); lanes without a query get lim = -inf */
            const float r_up = r < CUDART_INF_F ? __uint_as_float(__float_as_uint(r) + 1u) : r;
            const float lim = act ? ((d0 <= r) ? d0 : r_up) : -CUDART_INF_F;
            const float2* pts = reinterpret_cast<const float2*>(reg);
            unsigned short* mylist = qlist + slot * ql_stride;
#pragma unroll 4
            for (int i = part; i < R; i += stride) {
                const float2 p = pts[2 * i]; /* (x, y) of the uint4 record: broadcast */
                const float dist = l1_dist(qx, qy, p.x, p.y);
                if (dist < lim) {
                    const int j = atomicAdd(&qcnt[slot], 1);
                    if (j < ql_cap) mylist[j] = (unsigned short)i;
                }
            }
        }
        __syncthreads();
        if (g0 == 0) mbar_wait(&mbar[1], 0); /* descriptor rows have landed */
        for (int kk = warp; kk < nslot; kk += VISO_ST_WARPS) {
            const uint4 qrec = qrec_s[kk];
            const int n = qcnt[kk];
            if (n > ql_cap || n > P.K) { /* top-K cut or list overflow: left to the generic kernel */
                if (lane == 0) {
                    job.out[qrec.z] = make_int4(0, 0, 0, VISO_PENDING);
                    pend_push(pend, qrec, blockIdx.y);
                }
                continue;
            }
            const float qx = __uint_as_float(qrec.x), qy = __uint_as_float(qrec.y);
            unsigned short* qlst = qlist + kk * ql_stride;
            int nlist = n;
            if (P.epipolar) { /* Sampson gate (viso.cpp:695-701), lanes = candidates, compacting the list in place */
                nlist = 0;
                for (int base = 0; base < n; base += 32) {
                    const int e = base + lane;
                    bool take = e < n;
                    const unsigned short ri = qlst[take ? e : 0];
                    if (take) {
                        const uint4 rec = reg[ri];
                        const double sd = sampson_dev(P.F, qx, qy, __uint_as_float(rec.x), __uint_as_float(rec.y));
            if (!isfinite(sd) || sd > P.sampson_thresh) take = false;
                    }
                    const unsigned tm = __ballot_sync(FULL, take);
                    __syncwarp(); /* every lane has read its entry before the slots are reused */
                    if (take) qlst[nlist + __popc(tm & ((1u << lane) - 1))] = ri;
                    nlist += __popc(tm);
                }
                __syncwarp();
            }
            BestState st;
            st.b1 = 0xffffffffu; st.b2 = 0xffffffffu; st.bdist = 0; st.bidx = -1;
            if (nlist > 0) {
                /* the query row in this lane's chunk order */
                const uint4* qp = reinterpret_cast<const uint4*>(job.q.desc + (size_t)qpos_s[kk] * VISO_DESC_U16);
                uint4 qv[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) qv[s] = __ldg(qp + (s ^ rot));
                unsigned lb1 = 0xffffffffu, lb2 = 0xffffffffu, lkd = 0;
                int lki = -1;
                for (int base = 0; base < nlist; base += 32) {
                    const int e = base + lane;
                    const bool act = e < nlist;
                    const unsigned ri = qlst[act ? e : 0];
                    const uint4 rec = reg[ri];
                    const unsigned ra = rows_u32 + ri * 256u + (unsigned)rot * 16u;
                    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const unsigned ad = ra ^ ((unsigned)s << 4);
                        uint4 ta, tb;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(ta.x), "=r"(ta.y), "=r"(ta.z), "=r"(ta.w) : "r"(ad));
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+128];" : "=r"(tb.x), "=r"(tb.y), "=r"(tb.z), "=r"(tb.w) : "r"(ad));
                        a0 += __vminu2(qv[s].x, ta.x) + __vminu2(qv[s].y, ta.y);
                        a1 += __vminu2(qv[s].z, ta.z) + __vminu2(qv[s].w, ta.w);
                        a2 += __vminu2(qv[s + 8].x, tb.x) + __vminu2(qv[s + 8].y, tb.y);
                        a3 += __vminu2(qv[s + 8].z, tb.z) + __vminu2(qv[s + 8].w, tb.w);
                    }
                    /* every accumulator holds 16 words x 2 halves, each half <= 16 x 2047: no carry */
                    const unsigned tot = (a0 & 0xffffu) + (a0 >> 16) + (a1 & 0xffffu) + (a1 >> 16) + (a2 & 0xffffu) + (a2 >> 16) +
                                         (a3 & 0xffffu) + (a3 >> 16);
                    if (act) {
                        const unsigned sad = qrec.w + rec.w - 2u * tot;
                        const unsigned dbits = __float_as_uint(l1_dist(qx, qy, __uint_as_float(rec.x), __uint_as_float(rec.y)));
                        const int idx = (int)rec.z;
                        if (sad < lb1) {
                            lb2 = lb1; lb1 = sad; lkd = dbits; lki = idx;
                        } else if (sad == lb1) {
                            lb2 = lb1;
                            if (dbits > lkd || (dbits == lkd && idx > lki)) { lkd = dbits; lki = idx; }
                        } else if (sad < lb2) {
                            lb2 = sad;
                        }
                    }
                }
                /* fold the lanes: smallest SAD, ties to the largest (L1, index) key; second smallest with multiplicity */
                const unsigned m1 = __reduce_min_sync(FULL, lb1);
                const unsigned ties = __ballot_sync(FULL, lb1 == m1);
                unsigned m2 = m1;
                if (__popc(ties) < 2) m2 = __reduce_min_sync(FULL, lb1 == m1 ? lb2 : lb1);
                const unsigned kd = __reduce_max_sync(FULL, lb1 == m1 ? lkd : 0u);
                const int ki = __reduce_max_sync(FULL, (lb1 == m1 && lkd == kd) ? lki : -1);
                st.b1 = m1; st.b2 = m2; st.bdist = kd; st.bidx = ki;
                pairs += nlist;
            }
            if (lane == 0) write_result(job, P, (int)qrec.z, st);
        }
    }
    if (sad_pairs && lane == 0 && pairs) {
        atomicAdd(sad_pairs, (unsigned long long)pairs);
        atomicAdd(sad_pairs + 1, (unsigned long long)pairs);
    }
}

/* ------------------------------------------------------------------------------------------------ launchers */

cudaError_t viso_launch_pack(const PackJob* jobs, int n_jobs, int max_n, int dlen, int* err_flag, cudaStream_t s)
{
    if (n_jobs <= 0 || max_n <= 0) return cudaSuccess;
    dim3 grid((max_n + 7) / 8, n_jobs);
    pack_desc_kernel<<<grid, 256, 0, s>>>(jobs, dlen, err_flag);
    return cudaGetLastError();
}

__global__ void zero_words_kernel(int* p, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0;
}

cudaError_t viso_launch_zero(void* p, int n_words, cudaStream_t s)
{
    zero_words_kernel<<<1, 32, 0, s>>>(static_cast<int*>(p), n_words);
    return cudaGetLastError();
}

cudaError_t viso_launch_extract(const ExtractJob* jobs, int n_jobs, int max_n, int width, int height, int pitch, int radius,
                                cudaStream_t s)
{
    if (n_jobs <= 0 || max_n <= 0) return cudaSuccess;
    /* VISO_EXTRACT_ITERS iterations of 2 * VISO_EXTRACT_ROUNDS keypoints per warp */
    dim3 grid((max_n + 16 * VISO_EXTRACT_ROUNDS * VISO_EXTRACT_ITERS - 1) / (16 * VISO_EXTRACT_ROUNDS * VISO_EXTRACT_ITERS), n_jobs);
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

/* tile shape and staging capacities of sad_match_staged_kernel for the densest target set of the launch; false when
 * even the smallest tile's neighbourhood cannot be staged (dense sets, huge radii): the gather tile kernel runs then */
static bool staged_config(int max_nt, float r, GridCfg g, StagedCfg* out, size_t* smem_out)
{
    if (!(r >= 0.f) || !(r < 1e6f)) return false;
    const double ext_x = (double)g.gx * VISO_GRID_CS, ext_y = (double)g.gy * VISO_GRID_CS;
    const double dens = (double)max_nt / (ext_x * ext_y);
    /* per-query list capacity: 1.5 x the expected number of points in the L1 diamond (2 r^2) + 16, 64..256 */
    const double in_diamond = (double)max_nt * fmin(1.0, 2.0 * (double)r * r / (ext_x * ext_y));
    int ql_cap = (int)fmin(256.0, fmax(64.0, 1.5 * in_diamond + 16.0));
    ql_cap = (ql_cap + 31) & ~31;
    /* two CTAs per SM: 227 KB / 2 less the static shared memory and the 1 KB the system reserves per CTA */
    const size_t budget = 111 * 1024;
    const size_t lists = (size_t)VISO_ST_SLOTS * (ql_cap + 2) * sizeof(unsigned short);
    if (lists + 64 * 272 > budget) return false;
    int cap_rows = (int)((budget - lists) / 272) & ~7;
    static const int shapes[][2] = {{8, 4}, {6, 4}, {6, 3}, {4, 4}, {4, 3}, {4, 2}, {3, 2}, {2, 2}, {2, 1}, {1, 1}};
    for (const auto& sh : shapes) {
        const double bx = fmin(ext_x, sh[0] * VISO_GRID_CS + 2.0 * (r + 2) + VISO_GRID_CS);
        const double by = fmin(ext_y, sh[1] * VISO_GRID_CS + 2.0 * (r + 2) + VISO_GRID_CS);
        if (by / VISO_GRID_CS + 1 > VISO_ST_MAX_ROWS) continue;
        /* the four corners of the box are out of every query's reach and are not staged */
        const double area = fmax(0.25 * bx * by, bx * by - 1.4 * fmin((double)r * r, 0.25 * bx * by));
        if (1.3 * dens * area + 24.0 <= cap_rows) {
            out->tw = sh[0]; out->th = sh[1]; out->cap_rows = cap_rows; out->ql_cap = ql_cap;
            *smem_out = (size_t)cap_rows * 272 + lists;
            return true;
        }
    }
    return false;
}

cudaError_t viso_launch_match(const MatchJob* jobs, int n_jobs, int max_nq, int max_nt, const MatchParamsPair& mp,
                              GridCfg g, unsigned long long* sad_pairs, PendingList pend, int mode, int sm_count,
                              cudaStream_t s, int* launches)
{
    if (n_jobs <= 0 || max_nq <= 0) return cudaSuccess;
    if (sm_count <= 0) sm_count = 148;
    /* the generic kernel loops over query chunks: about 16 CTAs per SM in total, never more than one per chunk */
    const int gchunks = (max_nq + VISO_STRIP_QPC - 1) / VISO_STRIP_QPC;
    const dim3 ggrid(std::min(gchunks, std::max(1, (sm_count * 16 + n_jobs - 1) / n_jobs)), n_jobs);
    /* Dense target sets -- the expected number of points in a query's L1 diamond (2 r^2) exceeds half of max_neighbors,
     * so most queries need the exact (L1, index) top-K cut of viso.cpp:180-186 -- go to the generic kernel at once:
     * measured at 20 000 keypoints per image (BASELINE configs[2]) 0.21 / 0.31 ms per stereo / temporal match_desc call
     * against 0.33 / 0.42 ms for "tile kernel first, then every query pending" and 0.34 / 0.50 ms for running the
     * top-K selection inside the tile kernel on the staged records (2 x 2-cell tiles; removed again). */
    {
        const float rr = fmaxf(mp.p[0].radius, mp.p[1].radius);
        const double ext = (double)g.gx * VISO_GRID_CS * (double)g.gy * VISO_GRID_CS;
        const double in_diamond = rr >= 0.f ? (double)max_nt * fmin(1.0, 2.0 * (double)rr * rr / ext) : 0.0;
        if (mode == VISO_MATCH_AUTO && in_diamond > 0.5 * std::min(mp.p[0].K, mp.p[1].K)) mode = VISO_MATCH_GENERIC;
    }
    if (mode == VISO_MATCH_GENERIC) {
        if (pend.count) { /* the counter is reported by viso_seq_last_pending: nothing is pending on this path */
            cudaError_t ez = viso_launch_zero(pend.count, 1, s);
            if (ez != cudaSuccess) return ez;
            if (launches) *launches += 1;
        }
        sad_match_generic_kernel<<<ggrid, VISO_MATCH_WARPS * 32, 0, s>>>(jobs, mp, g, sad_pairs, pend, 0);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const float r = fmaxf(mp.p[0].radius, mp.p[1].radius);
    /* a kernel, not cudaMemsetAsync: memsets and copies on the compute stream can be scheduled on a copy engine and
     * then wait behind every upload queued there (see viso_seq_run_range) */
    cudaError_t e = viso_launch_zero(pend.count, 1, s);
    if (e != cudaSuccess) return e;
    StagedCfg sc;
    size_t st_smem = 0;
    /* AUTO runs the gather tile kernel: the staged kernel measured slower on B200 (10.6 vs 7.0 ms per 1000-frame launch,
     * DESIGN.md section 5); VISO_MATCH_STAGED selects it */
    bool staged = mode == VISO_MATCH_STAGED && staged_config(max_nt, r, g, &sc, &st_smem);
    if (!staged && mode == VISO_MATCH_STAGED) { /* forced: smallest tile, whatever does not fit goes to the generic kernel */
        sc.tw = 1; sc.th = 1; sc.ql_cap = 256;
        const size_t lists = (size_t)VISO_ST_SLOTS * (sc.ql_cap + 2) * sizeof(unsigned short);
        sc.cap_rows = (int)((111 * 1024 - lists) / 272) & ~7;
        st_smem = (size_t)sc.cap_rows * 272 + lists;
        staged = true;
    }
    if (staged) {
        e = cudaFuncSetAttribute(sad_match_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem);
        if (e != cudaSuccess) return e;
        const int tiles = ((g.gx + sc.tw - 1) / sc.tw) * ((g.gy + sc.th - 1) / sc.th);
        sad_match_staged_kernel<<<dim3(tiles, n_jobs), VISO_ST_WARPS * 32, st_smem, s>>>(jobs, mp, g, sc, sad_pairs, pend);
    } else {
        /* staging capacity for a tile's neighbourhood: the expected point count of the grown tile box less its four
         * corners (out of every query's reach, not staged) at the densest target set, x 1.35 + 48, within [128, 6144]
         * records.  Shared memory not used here is L1 for the descriptor rows. */
        const double ext_x = (double)g.gx * VISO_GRID_CS, ext_y = (double)g.gy * VISO_GRID_CS;
        /* per-query list capacity: 1.5 x the expected number of points in the L1 diamond (2 r^2) + 16, 64..256 */
        const double in_diamond = (double)max_nt * fmin(1.0, 2.0 * (double)r * r / (ext_x * ext_y));
        const int tw = VISO_TILE_W, th = VISO_TILE_H;
        const double bx = fmin(ext_x, tw * VISO_GRID_CS + 2.0 * (r + 2) + VISO_GRID_CS);
        const double by = fmin(ext_y, th * VISO_GRID_CS + 2.0 * (r + 2) + VISO_GRID_CS);
        const double area = fmax(0.25 * bx * by, bx * by - 1.4 * fmin((double)r * r, 0.25 * bx * by));
        double expect = (double)max_nt * area / (ext_x * ext_y);
        if (!(expect >= 0)) expect = 0;
        static const double cap_scale = getenv("VISO_GATHER_CAP_SCALE") ? atof(getenv("VISO_GATHER_CAP_SCALE")) : 1.35;
        int cap = (int)fmin(6144.0, fmax(128.0, cap_scale * expect + 48.0));
        cap = (cap + 31) & ~31;
        int ql_cap = (int)fmin(256.0, fmax(64.0, 1.5 * in_diamond + 16.0));
        ql_cap = (ql_cap + 31) & ~31;
        const size_t list_bytes = (size_t)32 * (ql_cap + 2) * sizeof(unsigned short);
        const size_t smem = (size_t)cap * (sizeof(uint4) + sizeof(int)) + list_bytes;
        const int tiles = ((g.gx + tw - 1) / tw) * ((g.gy + th - 1) / th);
        auto launch = [&](auto kern) -> cudaError_t {
            if (smem > 48 * 1024) { /* per device, so set whenever it is needed (a process may drive several GPUs) */
                cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e2 != cudaSuccess) return e2;
            }
            kern<<<dim3(tiles, n_jobs), VISO_MATCH_WARPS * 32, smem, s>>>(jobs, mp, g, cap, ql_cap, sad_pairs, pend);
            return cudaSuccess;
        };
        e = launch(sad_match_kernel<VISO_TILE_W, VISO_TILE_H>);
        if (e != cudaSuccess) return e;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    sad_match_generic_kernel<<<ggrid, VISO_MATCH_WARPS * 32, 0, s>>>(jobs, mp, g, sad_pairs, pend, 1);
    if (launches) *launches += 3; /* zero_words, tile kernel, generic kernel */
    return cudaGetLastError();
}
