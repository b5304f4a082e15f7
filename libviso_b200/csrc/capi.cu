/*
 * capi.cu -- the C-ABI of include/viso_b200.h: context, standalone entry points (host buffers in, host buffers out)
 * and host bookkeeping; the batched sequence pipeline is in capi_seq.cu.
 *
 * Host side only; every numeric result on the hot path is produced by the kernels in match.cu, sort_circle.cu, estimation.cu and geometry.cu.  There is no
 * CPU fallback: without a CUDA device viso_create() fails.  The only arithmetic done here is the once-per-sequence /
 * per-pose host bookkeeping the reference also keeps outside the per-frame loop (F_from_P, tr2mat, pose chaining)
 * and the RANSAC sample-table generators.
 */
#include "capi_internal.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace viso_capi {

/* bump allocator over one device block: first pass (base == null) measures, second pass hands out pointers */
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(char* b) : base(b) {}
    template <class T> T* take(size_t n)
    {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += std::max<size_t>(n, 1) * sizeof(T);
        return p;
    }
};

int ensure_scratch(viso_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->d_cap) return VISO_OK;
    if (ctx->d_scr) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(ctx->d_scr));
        ctx->d_scr = nullptr;
        ctx->d_cap = 0;
    }
    size_t cap = bytes + bytes / 4 + (1 << 20);
    cudaError_t e = cudaMalloc(&ctx->d_scr, cap);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? VISO_ERR_NOMEM : VISO_ERR_CUDA;
    }
    ctx->d_cap = cap;
    return VISO_OK;
}

ParamDev make_param_dev(const viso_param* p)
{
    ParamDev d;
    d.base = p->base; d.f = p->f; d.cu = p->cu; d.cv = p->cv;
    d.thr2 = p->inlier_threshold * p->inlier_threshold; /* viso.cpp:1532 */
    d.thresh = p->thresh;
    d.H = p->ransac_iter;
    d.pad = 0;
    return d;
}

MatchParamsDev make_match_dev(const viso_match_params* p)
{
    MatchParamsDev d;
    d.radius = (float)p->radius; /* radiusSearch takes float radius, viso.cpp:171-172 */
    d.K = p->max_neighbors;
    d.epipolar = p->enforce_epipolar ? 1 : 0;
    d.second_best = p->enforce_2nd_best ? 1 : 0;
    d.sampson_thresh = p->sampson_thresh;
    d.ratio = p->ratio_2nd_best;
    for (int i = 0; i < 9; ++i) d.F[i] = p->F[i];
    return d;
}

int status_from_flags(viso_ctx* ctx, int flags)
{
    if (flags & 1) return ctx->fail(VISO_ERR_DOMAIN, "descriptor outside the supported domain (integer valued, |v| <= 1023)");
    if (flags & 2) return ctx->fail(VISO_ERR_DUPLICATE, "match_circle: repeated query index in a Matches list");
    if (flags & 4) return ctx->fail(VISO_ERR_DIV0, "h2e: division by zero");
    if (flags & 8) return ctx->fail(VISO_ERR_ARG, "match index out of range");
    return VISO_OK;
}

/* OpenCV hal LU (partial pivoting, eps = 100*DBL_EPSILON) on a small host matrix: used for the once-per-sequence
 * F_from_P determinants (mvg.h:62-64) and the per-pose 4x4 inverse (viso.cpp:1319).  b: m x n right-hand sides. */
int host_lu(double* A, int m, double* b, int n)
{
    const double eps = DBL_EPSILON * 100;
    int p = 1;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::fabs(A[j * m + i]) > std::fabs(A[k * m + i])) k = j;
        if (std::fabs(A[k * m + i]) < eps) return 0;
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * m + j], A[k * m + j]);
            if (b) for (int j = 0; j < n; j++) std::swap(b[i * n + j], b[k * n + j]);
            p = -p;
        }
        const double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            const double alpha = A[j * m + i] * d;
            for (int c = i + 1; c < m; c++) A[j * m + c] += alpha * A[i * m + c];
            if (b) for (int c = 0; c < n; c++) b[j * n + c] += alpha * b[i * n + c];
        }
    }
    if (b)
        for (int i = m - 1; i >= 0; i--)
            for (int j = 0; j < n; j++) {
                double s = b[i * n + j];
                for (int k = i + 1; k < m; k++) s -= A[i * m + k] * b[k * n + j];
                b[i * n + j] = s / A[i * m + i];
            }
    return p;
}

double host_det4(const double M[16])
{
    double a[16];
    std::memcpy(a, M, sizeof(a));
    double r = host_lu(a, 4, nullptr, 0);
    if (r != 0) for (int i = 0; i < 4; i++) r *= a[i * 4 + i];
    return r;
}

int make_harris_cfg(viso_ctx* ctx, int width, int height, int pitch, int n_features, int nbinx, int nbiny, float k,
                    HarrisCfg* out)
{
    if (width < 8 || height < 8 || pitch < width) return ctx->fail(VISO_ERR_ARG, "detector: images must be at least 8 x 8");
    if (nbinx < 1 || nbiny < 1 || n_features < 0) return ctx->fail(VISO_ERR_ARG, "detector: bad bin counts"); /* viso.cpp:919 */
    HarrisCfg c{};
    c.w = width; c.h = height; c.pitch = pitch; c.nbinx = nbinx; c.nbiny = nbiny;
    c.sx = width / nbinx; c.sy = height / nbiny;                         /* viso.cpp:932-933 */
    if (c.sx < 1 || c.sy < 1) return ctx->fail(VISO_ERR_ARG, "detector: more bins than pixels (viso.cpp:934)");
    c.per = n_features / (nbinx * nbiny);                                /* viso.cpp:943 */
    c.k = k;
    const double sc = 1.0 / ((double)(1 << 4) * 3 * 255.0);              /* cornerHarris: aperture 5, block 3, 8-bit */
    c.f0 = (float)(6.0 * sc); c.f1 = (float)(4.0 * sc); c.f2 = (float)(1.0 * sc);
    if (viso_harris_cells(c) > 65535 || viso_harris_smem(c) > 200 * 1024)
        return ctx->fail(VISO_ERR_DOMAIN, "detector: a bin may hold at most ~33000 pixels (its responses are kept in shared memory)");
    *out = c;
    return VISO_OK;
}

} // namespace viso_capi
using namespace viso_capi;

extern "C" {

int viso_abi_version(void) { return VISO_ABI_VERSION; }

int viso_create(viso_ctx** out, int device)
{
    if (!out) return VISO_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return VISO_ERR_CUDA; /* no CPU fallback */
    if (cudaSetDevice(device) != cudaSuccess) return VISO_ERR_CUDA;
    viso_ctx* ctx = new viso_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return VISO_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return VISO_ERR_CUDA;
    }
    ctx->own_copy_stream = ctx->copy_stream;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) ctx->sm_count = sms;
    if (const char* m = getenv("VISO_MATCH_MODE")) /* generic | gather | staged: force one matching path (tests, A/B) */
        ctx->match_mode = m[0] == 'g' && m[1] == 'e' ? VISO_MATCH_GENERIC : m[0] == 'g' ? VISO_MATCH_GATHER
                          : m[0] == 's' ? VISO_MATCH_STAGED : VISO_MATCH_AUTO;
    if (const char* c = getenv("VISO_HYP_IT_CAP")) ctx->hyp_it_cap = std::max(1, std::min(100, atoi(c)));
    *out = ctx;
    return VISO_OK;
}

int viso_set_hyp_iteration_cap(viso_ctx* ctx, int cap)
{
    if (!ctx) return VISO_ERR_ARG;
    if (cap < 1 || cap > 100) return ctx->fail(VISO_ERR_ARG, "set_hyp_iteration_cap: 1..100");
    ctx->hyp_it_cap = cap;
    return VISO_OK;
}

int viso_set_match_mode(viso_ctx* ctx, int mode)
{
    if (!ctx) return VISO_ERR_ARG;
    if (mode < VISO_MATCH_AUTO || mode > VISO_MATCH_STAGED) return ctx->fail(VISO_ERR_ARG, "set_match_mode: unknown mode");
    ctx->match_mode = mode;
    return VISO_OK;
}

int viso_share_copy_stream(viso_ctx* ctx, viso_ctx* owner)
{
    if (!ctx) return VISO_ERR_ARG;
    if (owner && owner->device != ctx->device) return ctx->fail(VISO_ERR_ARG, "share_copy_stream: contexts on different devices");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->copy_stream = owner ? owner->own_copy_stream : ctx->own_copy_stream;
    return VISO_OK;
}

void viso_destroy(viso_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_scr) cudaFree(ctx->d_scr);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->own_copy_stream);
    cudaStreamDestroy(ctx->own_copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* viso_last_error(const viso_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* viso_stream(viso_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int viso_sync(viso_ctx* ctx)
{
    if (!ctx) return VISO_ERR_ARG;
    CK(cudaStreamSynchronize(ctx->copy_stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VISO_OK;
}

int viso_set_image_extent(viso_ctx* ctx, int width, int height)
{
    if (!ctx || width <= 0 || height <= 0) return VISO_ERR_ARG;
    int gx = (width + VISO_GRID_CS - 1) / VISO_GRID_CS, gy = (height + VISO_GRID_CS - 1) / VISO_GRID_CS;
    /* the grid histogram lives in shared memory: cap the cell count (coordinates beyond clamp into border cells) */
    while ((size_t)(2 * gx * gy + 1) * sizeof(int) > 200 * 1024) {
        if (gx >= gy) gx = (gx + 1) / 2; else gy = (gy + 1) / 2;
    }
    ctx->grid.gx = gx;
    ctx->grid.gy = gy;
    return VISO_OK;
}

int64_t viso_launch_count(const viso_ctx* ctx) { return ctx ? ctx->launches : 0; }

int viso_timer_begin(viso_ctx* ctx)
{
    if (!ctx) return VISO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->t0) {
        CK(cudaEventCreate(&ctx->t0));
        CK(cudaEventCreate(&ctx->t1));
    }
    CK(cudaEventRecord(ctx->t0, ctx->stream));
    return VISO_OK;
}

int viso_timer_end(viso_ctx* ctx, float* ms)
{
    if (!ctx || !ms || !ctx->t0) return VISO_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->t1, ctx->stream));
    CK(cudaEventSynchronize(ctx->t1));
    CK(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return VISO_OK;
}

void viso_match_params_stereo(viso_match_params* p, const double F[9])
{
    /* MatchParams(Mat F), viso.cpp:62-71 */
    std::memset(p, 0, sizeof(*p));
    p->enforce_epipolar = 1;
    p->sampson_thresh = 1;
    p->enforce_2nd_best = 0;
    p->ratio_2nd_best = .8;
    p->max_neighbors = 200;
    p->radius = 80;
    if (F) for (int i = 0; i < 9; i++) p->F[i] = F[i];
}

void viso_match_params_temporal(viso_match_params* p)
{
    /* MatchParams(), viso.cpp:72-74 */
    std::memset(p, 0, sizeof(*p));
    p->enforce_epipolar = 0;
    p->enforce_2nd_best = 1;
    p->ratio_2nd_best = .9;
    p->max_neighbors = 250;
    p->radius = 80;
}

void viso_param_default(viso_param* p)
{
    /* param(), viso.h:60; base / calib are left uninitialised by the reference, zero here */
    std::memset(p, 0, sizeof(*p));
    p->ransac_iter = 50;
    p->inlier_threshold = 2;
    p->thresh = 1e-4;
}

/* ------------------------------------------------------------------------------------------------ match_desc */

static int match_desc_impl(viso_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2, const float* d1,
                           const float* d2, int dlen, const viso_match_params* params, bool sorted,
                           int32_t* best_idx, int32_t* best_d1, int32_t* best_d2, int32_t* valid,
                           int32_t* matches, int32_t* n_matches)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n1 < 0 || n2 < 0 || !params) return ctx->fail(VISO_ERR_ARG, "match_desc: bad argument");
    if ((n1 > 0 && (!kp1 || !d1)) || (n2 > 0 && (!kp2 || !d2))) return ctx->fail(VISO_ERR_ARG, "match_desc: null input");
    if (dlen < 1 || dlen > VISO_DESC_U16 - 2) return ctx->fail(VISO_ERR_DOMAIN, "match_desc: desc_len must be 1..126");
    if (params->max_neighbors < 1) return ctx->fail(VISO_ERR_DOMAIN, "match_desc: max_neighbors must be >= 1");
    if (n_matches) *n_matches = 0;
    if (n1 == 0) return VISO_OK;
    CK(cudaSetDevice(ctx->device));
    const int ncell = ctx->ncell();

    struct Bufs {
        float2 *xy1, *xy2;
        uint4 *srec1, *srec2;
        int *po1, *po2;
        float *df1, *df2;
        uint16_t *du1, *du2;
        int *cell1, *cell2, *counts, *err, *matches, *mcount;
        int4* out;
        unsigned long long* pairs;
        int* pending;
        PackJob* pack;
        GridJob* grid;
        MatchJob* match;
        SortJob* sort;
    } b;
    auto carve = [&](Carver& c) {
        b.xy1 = c.take<float2>(n1); b.xy2 = c.take<float2>(n2);
        b.srec1 = c.take<uint4>(n1); b.srec2 = c.take<uint4>(n2);
        b.po1 = c.take<int>(n1); b.po2 = c.take<int>(n2);
        b.df1 = c.take<float>((size_t)n1 * dlen); b.df2 = c.take<float>((size_t)n2 * dlen);
        b.du1 = c.take<uint16_t>((size_t)n1 * VISO_DESC_U16); b.du2 = c.take<uint16_t>((size_t)n2 * VISO_DESC_U16);
        b.cell1 = c.take<int>(ncell + 1); b.cell2 = c.take<int>(ncell + 1);
        b.counts = c.take<int>(4); b.err = c.take<int>(1);
        b.matches = c.take<int>((size_t)n1 * 3); b.mcount = c.take<int>(1);
        b.out = c.take<int4>(n1);
        b.pairs = c.take<unsigned long long>(1);
        b.pending = c.take<int>(1);
        b.pack = c.take<PackJob>(2); b.grid = c.take<GridJob>(2); b.match = c.take<MatchJob>(1); b.sort = c.take<SortJob>(1);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);

    cudaStream_t s = ctx->stream;
    const int counts[4] = {n1, n2, 0, 0};
    CK(cudaMemcpyAsync(b.counts, counts, sizeof(counts), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(b.err, 0, sizeof(int), s));
    CK(cudaMemcpyAsync(b.xy1, kp1, (size_t)n1 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.df1, d1, (size_t)n1 * dlen * 4, cudaMemcpyHostToDevice, s));
    if (n2 > 0) {
        CK(cudaMemcpyAsync(b.xy2, kp2, (size_t)n2 * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b.df2, d2, (size_t)n2 * dlen * 4, cudaMemcpyHostToDevice, s));
    }
    PackJob pj[2] = {{b.df1, b.counts, b.du1, b.srec1, nullptr}, {b.df2, b.counts + 1, b.du2, b.srec2, nullptr}};
    GridJob gj[2] = {{b.xy1, b.counts, b.po1, b.srec1, b.cell1}, {b.xy2, b.counts + 1, b.po2, b.srec2, b.cell2}};
    MatchJob mj;
    mj.q = SetView{b.xy1, b.counts, b.du1, b.srec1, b.cell1, b.po1};
    mj.t = SetView{b.xy2, b.counts + 1, b.du2, b.srec2, b.cell2, b.po2};
    mj.out = b.out; mj.mode = 0; mj.pad = 0;
    SortJob sj;
    sj.dense = b.out; sj.n = b.counts; sj.kp1 = b.xy1; sj.kp2 = b.xy2; sj.matches = b.matches; sj.count = b.mcount;
    sj.pos_of_query = nullptr; sj.x = nullptr; sj.X = nullptr; sj.stride = n1; sj.pad = 0;
    CK(cudaMemcpyAsync(b.pack, pj, sizeof(pj), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.grid, gj, sizeof(gj), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.match, &mj, sizeof(mj), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.sort, &sj, sizeof(sj), cudaMemcpyHostToDevice, s));

    MatchParamsPair mp;
    mp.p[0] = make_match_dev(params);
    mp.p[1] = mp.p[0];
    CK(viso_launch_grid(b.grid, 2, ctx->grid, s));
    CK(viso_launch_pack(b.pack, 2, std::max(n1, n2), dlen, b.err, s)); /* rows in cell-sorted order: after the grid */
    int ml = 0;
    CK(viso_launch_match(b.match, 1, n1, n2, mp, ctx->grid, nullptr, PendingList{b.pending, nullptr, nullptr, 0},
                         ctx->match_mode, ctx->sm_count, s, &ml));
    ctx->launches += 2 + ml;
    std::vector<int4> host_out;
    std::vector<int> host_m;
    int flags = 0, mcount = 0;
    if (sorted) {
        ParamDev pd{};
        CK(viso_launch_sort(b.sort, 1, n1, pd, s));
        ctx->launches += 1;
        host_m.resize((size_t)n1 * 3);
        CK(cudaMemcpyAsync(host_m.data(), b.matches, (size_t)n1 * 12, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&mcount, b.mcount, sizeof(int), cudaMemcpyDeviceToHost, s));
    } else {
        host_out.resize(n1);
        CK(cudaMemcpyAsync(host_out.data(), b.out, (size_t)n1 * sizeof(int4), cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(&flags, b.err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    rc = status_from_flags(ctx, flags);
    if (rc) return rc;
    if (sorted) {
        if (matches) std::memcpy(matches, host_m.data(), (size_t)mcount * 12);
        if (n_matches) *n_matches = mcount;
    } else {
        for (int i = 0; i < n1; ++i) {
            if (best_idx) best_idx[i] = host_out[i].x;
            if (best_d1) best_d1[i] = host_out[i].y;
            if (best_d2) best_d2[i] = host_out[i].z;
            if (valid) valid[i] = host_out[i].w;
        }
    }
    return VISO_OK;
}

int viso_match_desc(viso_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2, const float* d1, const float* d2,
                    int desc_len, const viso_match_params* params, int32_t* best_idx, int32_t* best_d1, int32_t* best_d2,
                    int32_t* valid)
{
    return match_desc_impl(ctx, kp1, n1, kp2, n2, d1, d2, desc_len, params, false, best_idx, best_d1, best_d2, valid,
                           nullptr, nullptr);
}

int viso_match_desc_sorted(viso_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2, const float* d1,
                           const float* d2, int desc_len, const viso_match_params* params, int32_t* matches,
                           int32_t* n_matches)
{
    if (!n_matches) return ctx ? ctx->fail(VISO_ERR_ARG, "match_desc_sorted: null n_matches") : VISO_ERR_ARG;
    return match_desc_impl(ctx, kp1, n1, kp2, n2, d1, d2, desc_len, params, true, nullptr, nullptr, nullptr, nullptr,
                           matches, n_matches);
}

int viso_sort_matches(viso_ctx* ctx, int32_t* matches, int n)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 0 || (n > 0 && !matches)) return ctx->fail(VISO_ERR_ARG, "sort_matches: bad argument");
    if (n == 0) return VISO_OK;
    CK(cudaSetDevice(ctx->device));
    /* the sort kernel consumes dense per-query results: position i plays the query index */
    std::vector<int4> dense(n);
    for (int i = 0; i < n; ++i) dense[i] = make_int4(matches[3 * i + 1], matches[3 * i + 2], 0, 1);
    struct Bufs { int4* dense; int *n, *matches, *count; SortJob* job; } b;
    auto carve = [&](Carver& c) {
        b.dense = c.take<int4>(n); b.n = c.take<int>(1); b.matches = c.take<int>((size_t)n * 3); b.count = c.take<int>(1);
        b.job = c.take<SortJob>(1);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.dense, dense.data(), (size_t)n * sizeof(int4), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.n, &n, 4, cudaMemcpyHostToDevice, s));
    SortJob sj;
    std::memset(&sj, 0, sizeof(sj));
    sj.dense = b.dense; sj.n = b.n; sj.matches = b.matches; sj.count = b.count; sj.stride = n;
    CK(cudaMemcpyAsync(b.job, &sj, sizeof(sj), cudaMemcpyHostToDevice, s));
    ParamDev pd{};
    CK(viso_launch_sort(b.job, 1, n, pd, s));
    ctx->launches += 1;
    std::vector<int> out((size_t)n * 3);
    CK(cudaMemcpyAsync(out.data(), b.matches, (size_t)n * 12, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<int> first(n);
    for (int i = 0; i < n; ++i) first[i] = matches[3 * i];
    for (int p = 0; p < n; ++p) {
        matches[3 * p] = first[out[3 * p]];
        matches[3 * p + 1] = out[3 * p + 1];
        matches[3 * p + 2] = out[3 * p + 2];
    }
    return VISO_OK;
}

/* ------------------------------------------------------------------------------------------------ match_circle */

int viso_match_circle(viso_ctx* ctx, const int32_t* match_lr, int nlr, const int32_t* match_lr_prev, int nlrp,
                      const int32_t* match11, int n11, const int32_t* match22, int n22, int32_t* circ4, int32_t* pcl3,
                      int32_t* n_out)
{
    if (!ctx) return VISO_ERR_ARG;
    if (!n_out || nlr < 0 || nlrp < 0 || n11 < 0 || n22 < 0) return ctx->fail(VISO_ERR_ARG, "match_circle: bad argument");
    *n_out = 0;
    if (nlr == 0 || nlrp == 0 || n11 == 0 || n22 == 0) return VISO_OK;
    if (!match_lr || !match_lr_prev || !match11 || !match22) return ctx->fail(VISO_ERR_ARG, "match_circle: null input");
    CK(cudaSetDevice(ctx->device));
    /* lookup-table extents: largest query index per list (host has the lists anyway) */
    auto max_key = [](const int32_t* m, int n) { int mx = -1; for (int i = 0; i < n; ++i) mx = std::max(mx, m[3 * i]); return mx; };
    for (int i = 0; i < n11; ++i) if (match11[3 * i] < 0) return ctx->fail(VISO_ERR_ARG, "match_circle: negative index");
    for (int i = 0; i < nlrp; ++i) if (match_lr_prev[3 * i] < 0) return ctx->fail(VISO_ERR_ARG, "match_circle: negative index");
    for (int i = 0; i < n22; ++i) if (match22[3 * i] < 0) return ctx->fail(VISO_ERR_ARG, "match_circle: negative index");
    const int n_t11 = max_key(match11, n11) + 1, n_tlrp = max_key(match_lr_prev, nlrp) + 1, n_t22 = max_key(match22, n22) + 1;

    struct Bufs { int *lr, *lrp, *m11, *m22, *t11, *tlrp, *t22, *circ4, *pcl3, *n_out, *err; } b;
    auto carve = [&](Carver& c) {
        b.lr = c.take<int>((size_t)nlr * 3); b.lrp = c.take<int>((size_t)nlrp * 3);
        b.m11 = c.take<int>((size_t)n11 * 3); b.m22 = c.take<int>((size_t)n22 * 3);
        b.t11 = c.take<int>(n_t11); b.tlrp = c.take<int>(n_tlrp); b.t22 = c.take<int>(n_t22);
        b.circ4 = c.take<int>((size_t)nlr * 4); b.pcl3 = c.take<int>((size_t)nlr * 3);
        b.n_out = c.take<int>(1); b.err = c.take<int>(1);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.lr, match_lr, (size_t)nlr * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.lrp, match_lr_prev, (size_t)nlrp * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.m11, match11, (size_t)n11 * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.m22, match22, (size_t)n22 * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(b.t11, 0xff, (size_t)n_t11 * 4, s));
    CK(cudaMemsetAsync(b.tlrp, 0xff, (size_t)n_tlrp * 4, s));
    CK(cudaMemsetAsync(b.t22, 0xff, (size_t)n_t22 * 4, s));
    CK(cudaMemsetAsync(b.err, 0, 4, s));
    CK(viso_launch_circle_tables(b.m11, n11, b.t11, n_t11, 0, 0, b.err, s));
    CK(viso_launch_circle_tables(b.lrp, nlrp, b.tlrp, n_tlrp, 0, 1, b.err, s));
    CK(viso_launch_circle_tables(b.m22, n22, b.t22, n_t22, 0, 0, b.err, s));
    CK(viso_launch_circle_generic(b.lr, nlr, b.lrp, nlrp, b.t11, n_t11, b.tlrp, n_tlrp, b.t22, n_t22, b.circ4, b.pcl3,
                                  b.n_out, s));
    ctx->launches += 4;
    int flags = 0, c = 0;
    CK(cudaMemcpyAsync(&flags, b.err, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&c, b.n_out, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    rc = status_from_flags(ctx, flags);
    if (rc) return rc;
    if (c > 0) {
        if (circ4) CK(cudaMemcpyAsync(circ4, b.circ4, (size_t)c * 16, cudaMemcpyDeviceToHost, s));
        if (pcl3) CK(cudaMemcpyAsync(pcl3, b.pcl3, (size_t)c * 12, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    *n_out = c;
    return VISO_OK;
}

/* ------------------------------------------------------------------------------------------------ geometry */

int viso_collect_triangulate(viso_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2, const int32_t* matches,
                             int m, double f, double base, double cu, double cv, double* x, double* X)
{
    if (!ctx) return VISO_ERR_ARG;
    if (m < 0 || n1 < 0 || n2 < 0) return ctx->fail(VISO_ERR_ARG, "collect_triangulate: bad argument");
    if (m == 0) return VISO_OK;
    if (!kp1 || !kp2 || !matches || n1 == 0 || n2 == 0) return ctx->fail(VISO_ERR_ARG, "collect_triangulate: null input");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { float2 *k1, *k2; int *mt, *err; double *x, *X; } b;
    auto carve = [&](Carver& c) {
        b.k1 = c.take<float2>(n1); b.k2 = c.take<float2>(n2); b.mt = c.take<int>((size_t)m * 3); b.err = c.take<int>(1);
        b.x = c.take<double>((size_t)m * 4); b.X = c.take<double>((size_t)m * 3);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.k1, kp1, (size_t)n1 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.k2, kp2, (size_t)n2 * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.mt, matches, (size_t)m * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(b.err, 0, 4, s));
    ParamDev pd{};
    pd.base = base; pd.f = f; pd.cu = cu; pd.cv = cv;
    CK(viso_launch_collect_tri(b.k1, n1, b.k2, n2, b.mt, m, b.x, X ? b.X : nullptr, pd, b.err, s));
    ctx->launches += 1;
    int flags = 0;
    CK(cudaMemcpyAsync(&flags, b.err, 4, cudaMemcpyDeviceToHost, s));
    if (x) CK(cudaMemcpyAsync(x, b.x, (size_t)m * 32, cudaMemcpyDeviceToHost, s));
    if (X) CK(cudaMemcpyAsync(X, b.X, (size_t)m * 24, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return status_from_flags(ctx, flags);
}

int viso_triangulate_rectified_f64(viso_ctx* ctx, const double* x, int m, double f, double base, double cu, double cv,
                                   double* X)
{
    if (!ctx) return VISO_ERR_ARG;
    if (m < 0) return ctx->fail(VISO_ERR_ARG, "triangulate: bad argument");
    if (m == 0) return VISO_OK;
    if (!x || !X) return ctx->fail(VISO_ERR_ARG, "triangulate: null input");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { double *x, *X; } b;
    auto carve = [&](Carver& c) { b.x = c.take<double>((size_t)m * 4); b.X = c.take<double>((size_t)m * 3); };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.x, x, (size_t)m * 32, cudaMemcpyHostToDevice, s));
    ParamDev pd{};
    pd.base = base; pd.f = f; pd.cu = cu; pd.cv = cv;
    CK(viso_launch_triangulate_f64(b.x, m, m, b.X, pd, s));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(X, b.X, (size_t)m * 24, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VISO_OK;
}

int viso_triangulate_rectified_f32(viso_ctx* ctx, const float* x1, const float* x2, int m, double f, double base,
                                   double c1u, double c1v, float* X)
{
    if (!ctx) return VISO_ERR_ARG;
    if (m < 0) return ctx->fail(VISO_ERR_ARG, "triangulate: bad argument");
    if (m == 0) return VISO_OK;
    if (!x1 || !x2 || !X) return ctx->fail(VISO_ERR_ARG, "triangulate: null input");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { float *x1, *x2, *X; } b;
    auto carve = [&](Carver& c) { b.x1 = c.take<float>((size_t)m * 2); b.x2 = c.take<float>((size_t)m * 2); b.X = c.take<float>((size_t)m * 3); };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.x1, x1, (size_t)m * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.x2, x2, (size_t)m * 8, cudaMemcpyHostToDevice, s));
    CK(viso_launch_triangulate_f32(b.x1, b.x2, m, f, base, c1u, c1v, b.X, s));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(X, b.X, (size_t)m * 12, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VISO_OK;
}

int viso_project_points(viso_ctx* ctx, const double* X, int n, const double P[12], double* x)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 0) return ctx->fail(VISO_ERR_ARG, "project: bad argument");
    if (n == 0) return VISO_OK;
    if (!X || !P || !x) return ctx->fail(VISO_ERR_ARG, "project: null input");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { double *X, *P, *x; int* err; } b;
    auto carve = [&](Carver& c) { b.X = c.take<double>((size_t)n * 3); b.P = c.take<double>(12); b.x = c.take<double>((size_t)n * 2); b.err = c.take<int>(1); };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.X, X, (size_t)n * 24, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.P, P, 96, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(b.err, 0, 4, s));
    CK(viso_launch_project(b.X, n, b.P, b.x, b.err, s));
    ctx->launches += 1;
    int flags = 0;
    CK(cudaMemcpyAsync(&flags, b.err, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(x, b.x, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return status_from_flags(ctx, flags);
}

int viso_triangulate_dlt(viso_ctx* ctx, const float* x1, const float* x2, int m, const double P1[12], const double P2[12],
                         float* X)
{
    if (!ctx) return VISO_ERR_ARG;
    if (m < 0) return ctx->fail(VISO_ERR_ARG, "triangulate_dlt: bad argument");
    if (m == 0) return VISO_OK;
    if (!x1 || !x2 || !P1 || !P2 || !X) return ctx->fail(VISO_ERR_ARG, "triangulate_dlt: null input");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { float *x1, *x2, *X; double *P1, *P2; } b;
    auto carve = [&](Carver& c) {
        b.x1 = c.take<float>((size_t)m * 2); b.x2 = c.take<float>((size_t)m * 2); b.X = c.take<float>((size_t)m * 3);
        b.P1 = c.take<double>(12); b.P2 = c.take<double>(12);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.x1, x1, (size_t)m * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.x2, x2, (size_t)m * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.P1, P1, 96, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.P2, P2, 96, cudaMemcpyHostToDevice, s));
    CK(viso_launch_triangulate_dlt(b.x1, b.x2, m, b.P1, b.P2, b.X, s));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(X, b.X, (size_t)m * 12, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VISO_OK;
}

int viso_solve_rigid_motion(viso_ctx* ctx, const float* A, const float* B, int n, float T[16])
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 2 || !A || !B || !T) return ctx->fail(VISO_ERR_ARG, "solve_rigid_motion: needs at least 2 points (estimation.cpp:32)");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { float *A, *B, *T; } b;
    auto carve = [&](Carver& c) { b.A = c.take<float>((size_t)n * 3); b.B = c.take<float>((size_t)n * 3); b.T = c.take<float>(16); };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.A, A, (size_t)n * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.B, B, (size_t)n * 12, cudaMemcpyHostToDevice, s));
    CK(viso_launch_rigid_motion(b.A, b.B, n, b.T, s));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(T, b.T, 64, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return VISO_OK;
}

/* ------------------------------------------------------------------------------------------------ estimation */

int viso_get_inliers(viso_ctx* ctx, const double* X, const double* observe, int n, const double tr[6],
                     const viso_param* param, int32_t* inliers, int32_t* n_inliers)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 0 || !tr || !param || !n_inliers) return ctx->fail(VISO_ERR_ARG, "get_inliers: bad argument");
    *n_inliers = 0;
    if (n == 0) return VISO_OK;
    if (!X || !observe) return ctx->fail(VISO_ERR_ARG, "get_inliers: null input");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { double *X, *obs, *tr; int *inl, *cnt; } b;
    auto carve = [&](Carver& c) {
        b.X = c.take<double>((size_t)n * 3); b.obs = c.take<double>((size_t)n * 4); b.tr = c.take<double>(6);
        b.inl = c.take<int>(n); b.cnt = c.take<int>(1);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.X, X, (size_t)n * 24, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.obs, observe, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.tr, tr, 48, cudaMemcpyHostToDevice, s));
    CK(viso_launch_inliers(b.X, b.obs, n, n, b.tr, b.inl, b.cnt, make_param_dev(param), s));
    ctx->launches += 1;
    int c = 0;
    CK(cudaMemcpyAsync(&c, b.cnt, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (c > 0 && inliers) {
        CK(cudaMemcpyAsync(inliers, b.inl, (size_t)c * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    *n_inliers = c;
    return VISO_OK;
}

int viso_minimize_reproj(viso_ctx* ctx, const double* X, const double* observe, int n, double tr[6],
                         const viso_param* param, const int32_t* active, int n_active, int32_t* ok)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 1 || !X || !observe || !tr || !param || !ok || n_active < 0 || (n_active > 0 && !active))
        return ctx->fail(VISO_ERR_ARG, "minimize_reproj: bad argument");
    if (n_active > n) return ctx->fail(VISO_ERR_ARG, "minimize_reproj: more active points than columns (viso.cpp:1449 reads observe(0,i))");
    for (int i = 0; i < n_active; ++i)
        if (active[i] < 0 || active[i] >= n) return ctx->fail(VISO_ERR_ARG, "minimize_reproj: active index out of range");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { double *X, *obs, *tr, *scratch; int *active, *ok; } b;
    auto carve = [&](Carver& c) {
        b.X = c.take<double>((size_t)n * 3); b.obs = c.take<double>((size_t)n * 4); b.tr = c.take<double>(6);
        b.scratch = c.take<double>((size_t)n_active * 28); b.active = c.take<int>(n_active); b.ok = c.take<int>(1);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.X, X, (size_t)n * 24, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.obs, observe, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.tr, tr, 48, cudaMemcpyHostToDevice, s));
    if (n_active > 0) CK(cudaMemcpyAsync(b.active, active, (size_t)n_active * 4, cudaMemcpyHostToDevice, s));
    CK(viso_launch_gn(b.X, b.obs, n, b.active, n_active, b.tr, b.ok, b.scratch, make_param_dev(param), s));
    ctx->launches += 1;
    int okv = 0;
    CK(cudaMemcpyAsync(tr, b.tr, 48, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&okv, b.ok, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *ok = okv;
    return VISO_OK;
}

int viso_ransac_minimize_reproj(viso_ctx* ctx, const double* X, const double* observe, int n, const viso_param* param,
                                const int32_t* sample_table, double tr[6], int32_t* inliers, int32_t* n_inliers,
                                int32_t* ok, double* hyp_tr, int32_t* hyp_ok, int32_t* hyp_count, int32_t* best_hyp)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 0 || !param || !tr || !ok || param->ransac_iter < 0 || (param->ransac_iter > 0 && !sample_table))
        return ctx->fail(VISO_ERR_ARG, "ransac_minimize_reproj: bad argument");
    *ok = 0;
    if (n_inliers) *n_inliers = 0;
    if (best_hyp) *best_hyp = -1;
    if (n < 3) return VISO_OK; /* the reference's sampler cannot draw 3 of fewer than 3; report failure */
    /* no hypotheses: the loop of viso.cpp:1555 never runs, best_inliers stays empty, :1571 returns false with best_tr
     * untouched */
    if (param->ransac_iter == 0) return VISO_OK;
    if (!X || !observe) return ctx->fail(VISO_ERR_ARG, "ransac_minimize_reproj: null input");
    const int H = param->ransac_iter;
    for (int i = 0; i < 3 * H; ++i)
        if (sample_table[i] < 0 || sample_table[i] >= n) return ctx->fail(VISO_ERR_ARG, "ransac_minimize_reproj: sample index out of range");
    CK(cudaSetDevice(ctx->device));
    struct Bufs {
        double *X, *obs, *hyp_tr, *scratch;
        int *n, *table, *hyp_ok, *hyp_count, *inliers, *active;
        viso_record_dev* rec;
        RansacProb* prob;
        int* strag;
    } b;
    auto carve = [&](Carver& c) {
        b.X = c.take<double>((size_t)n * 3); b.obs = c.take<double>((size_t)n * 4);
        b.hyp_tr = c.take<double>((size_t)H * 6); b.scratch = c.take<double>((size_t)n * 28);
        b.n = c.take<int>(1); b.table = c.take<int>((size_t)H * 3); b.hyp_ok = c.take<int>(H); b.hyp_count = c.take<int>(H);
        b.inliers = c.take<int>(n); b.active = c.take<int>(n);
        b.rec = c.take<viso_record_dev>(1); b.prob = c.take<RansacProb>(1);
        b.strag = c.take<int>(2 + 2 * (size_t)H);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(b.X, X, (size_t)n * 24, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.obs, observe, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.n, &n, 4, cudaMemcpyHostToDevice, s));
    if (H > 0) CK(cudaMemcpyAsync(b.table, sample_table, (size_t)H * 12, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(b.hyp_ok, 0, (size_t)std::max(H, 1) * 4, s));
    RansacProb pb;
    std::memset(&pb, 0, sizeof(pb));
    pb.X = b.X; pb.obs = b.obs; pb.n = b.n; pb.stride = n; pb.H = H; pb.seeds = nullptr; pb.table = b.table;
    pb.hyp_tr = b.hyp_tr; pb.hyp_ok = b.hyp_ok; pb.hyp_count = b.hyp_count; pb.scratch = b.scratch;
    pb.inliers = b.inliers; pb.active = b.active; pb.rec = b.rec; pb.min_n = 3;
    for (int j = 0; j < 6; ++j) pb.tr_init[j] = tr[j];
    CK(cudaMemcpyAsync(b.prob, &pb, sizeof(pb), cudaMemcpyHostToDevice, s));
    int nl = 0;
    CK(viso_launch_ransac(b.prob, 1, H, n, make_param_dev(param), b.strag, ctx->hyp_it_cap, ctx->sm_count, s, &nl));
    ctx->launches += nl;
    viso_record_dev rec;
    CK(cudaMemcpyAsync(&rec, b.rec, sizeof(rec), cudaMemcpyDeviceToHost, s));
    if (hyp_tr && H > 0) CK(cudaMemcpyAsync(hyp_tr, b.hyp_tr, (size_t)H * 48, cudaMemcpyDeviceToHost, s));
    if (hyp_ok && H > 0) CK(cudaMemcpyAsync(hyp_ok, b.hyp_ok, (size_t)H * 4, cudaMemcpyDeviceToHost, s));
    if (hyp_count && H > 0) CK(cudaMemcpyAsync(hyp_count, b.hyp_count, (size_t)H * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (rec.n_inliers < 0 || rec.n_inliers > n) return ctx->fail(VISO_ERR_CUDA, "ransac_minimize_reproj: corrupt result record");
    if (inliers && rec.n_inliers > 0) {
        CK(cudaMemcpyAsync(inliers, b.inliers, (size_t)rec.n_inliers * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    for (int j = 0; j < 6; ++j) tr[j] = rec.tr[j];
    *ok = rec.ok;
    if (n_inliers) *n_inliers = rec.n_inliers;
    if (best_hyp) *best_hyp = rec.best_hyp;
    return VISO_OK;
}

void viso_randomsample_table(uint32_t seed, int H, int N, int32_t* table)
{
    /* randomsample(n=3, N, samples), viso.cpp:87-107 (Knuth Algorithm S), drawing from one std::mt19937(seed) stream
     * instead of a fresh random_device-seeded generator per call */
    std::mt19937 gen(seed);
    std::uniform_real_distribution<> dis(0, 1);
    for (int h = 0; h < H; ++h) {
        int t = 0, m = 0;
        const int n = 3;
        while (m < n) {
            const double u = dis(gen);
            if ((N - t) * u >= n - m) {
                t++;
            } else {
                table[3 * h + m] = t;
                t++;
                m++;
            }
        }
    }
}

void viso_samples_from_seeds(const uint32_t* seeds, int H, int N, int32_t* table)
{
    /* three uniform draws without replacement by fixed-point scaling, returned ascending (the reference's sampler
     * also yields ascending distinct triples); the device evaluates the same integer expression per hypothesis */
    for (int h = 0; h < H; ++h) {
        const uint32_t r0 = seeds[3 * h], r1 = seeds[3 * h + 1], r2 = seeds[3 * h + 2];
        int a = (int)(((uint64_t)r0 * (uint64_t)N) >> 32);
        int b = (int)(((uint64_t)r1 * (uint64_t)(N - 1)) >> 32);
        int c = (int)(((uint64_t)r2 * (uint64_t)(N - 2)) >> 32);
        if (b >= a) b++;
        const int lo = std::min(a, b), hi = std::max(a, b);
        if (c >= lo) c++;
        if (c >= hi) c++;
        int s[3] = {lo, hi, c};
        std::sort(s, s + 3);
        table[3 * h] = s[0]; table[3 * h + 1] = s[1]; table[3 * h + 2] = s[2];
    }
}

/* ------------------------------------------------------------------------------------------------ host bookkeeping */

void viso_tr2mat(const double tr[6], double T[16])
{
    /* viso.cpp:109-133 */
    const double rx = tr[0], ry = tr[1], rz = tr[2];
    const double sx = sin(rx), cx = cos(rx), sy = sin(ry), cy = cos(ry), sz = sin(rz), cz = cos(rz);
    T[0] = +cy * cz;                T[1] = -cy * sz;                T[2] = +sy;       T[3] = tr[3];
    T[4] = +sx * sy * cz + cx * sz; T[5] = -sx * sy * sz + cx * cz; T[6] = -sx * cy;  T[7] = tr[4];
    T[8] = -cx * sy * cz + sx * sz; T[9] = +cx * sy * sz + sx * cz; T[10] = +cx * cy; T[11] = tr[5];
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

void viso_F_from_P(const double P1[12], const double P2[12], int normalise, double F[9])
{
    /* F_from_P<double>, mvg.h:41-66: F(r,c) = det([P1 without row c ; P2 without row r]) in cyclic row order,
     * cv::determinant of a 4x4 = LU; then the normalisation of viso.cpp:1177-1180 */
    static const int rows[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double M[16];
            for (int k = 0; k < 4; k++) {
                M[0 * 4 + k] = P1[rows[c][0] * 4 + k];
                M[1 * 4 + k] = P1[rows[c][1] * 4 + k];
                M[2 * 4 + k] = P2[rows[r][0] * 4 + k];
                M[3 * 4 + k] = P2[rows[r][1] * 4 + k];
            }
            F[r * 3 + c] = host_det4(M);
        }
    if (normalise && F[8] > DBL_MIN) {
        /* viso.cpp:1177-1180, `F /= F.at<double>(2,2)`: cv::Mat's operator/=(Mat&, double) is a.convertTo(a, -1, 1./s)
         * (opencv2/core/mat.inl.hpp), i.e. a multiplication by the reciprocal, not a division */
        const double inv = 1. / F[8];
        for (int i = 0; i < 9; i++) F[i] = F[i] * inv;
    }
}

int viso_pose_update(const double pose[16], const double tr[6], double pose_out[16])
{
    /* pose = pose * tr_mat.inv(), viso.cpp:1315-1321 (cv::Mat::inv = LU) */
    double T[16], Ti[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    viso_tr2mat(tr, T);
    if (!host_lu(T, 4, Ti, 4)) return VISO_ERR_ARG;
    double out[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += pose[i * 4 + k] * Ti[k * 4 + j];
            out[i * 4 + j] = s;
        }
    std::memcpy(pose_out, out, sizeof(out));
    return VISO_OK;
}

int viso_chain_poses(const viso_record* records, int n_frames, double* poses)
{
    /* viso.cpp:1189-1190 (identity first) and :1313-1321 (append only when RANSAC succeeded) */
    double pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memcpy(poses, pose, sizeof(pose));
    int np = 1;
    for (int t = 1; t < n_frames; ++t) {
        if (records[t].n_circ < 3 || !records[t].ok) continue;
        double next[16];
        if (viso_pose_update(pose, records[t].tr, next) != VISO_OK) continue;
        std::memcpy(pose, next, sizeof(pose));
        std::memcpy(poses + 16 * np, pose, sizeof(pose));
        ++np;
    }
    return np;
}

/* ------------------------------------------------------------------------------------------------ detector */

int viso_detect_harris(viso_ctx* ctx, const uint8_t* img, int width, int height, int pitch, int n_features, int nbinx,
                       int nbiny, float k, float* kp_xy, float* kp_response, int32_t* n_out)
{
    if (!ctx) return VISO_ERR_ARG;
    if (!img || !kp_xy || !n_out) return ctx->fail(VISO_ERR_ARG, "detect_harris: null argument");
    HarrisCfg c;
    int rc = make_harris_cfg(ctx, width, height, pitch, n_features, nbinx, nbiny, k, &c);
    if (rc) return rc;
    *n_out = 0;
    if (c.per == 0) return VISO_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t nb = (size_t)nbinx * nbiny, slots = nb * c.per;
    struct Bufs { unsigned char* img; float2 *kp, *tmp; float *resp, *resp_tmp; int *n, *bin_count, *flag; DetectJob* job; } b;
    auto carve = [&](Carver& cv) {
        b.img = cv.take<unsigned char>((size_t)pitch * height);
        b.kp = cv.take<float2>(slots); b.tmp = cv.take<float2>(slots);
        b.resp = cv.take<float>(slots); b.resp_tmp = cv.take<float>(slots);
        b.n = cv.take<int>(1); b.bin_count = cv.take<int>(nb); b.flag = cv.take<int>(1);
        b.job = cv.take<DetectJob>(1);
    };
    Carver measure(nullptr);
    carve(measure);
    rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    const int one = 1;
    const DetectJob job{b.img, b.kp, b.n, b.tmp, b.resp_tmp, b.resp, b.bin_count, b.flag};
    CK(cudaMemcpyAsync(b.img, img, (size_t)pitch * height, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.flag, &one, 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.job, &job, sizeof(job), cudaMemcpyHostToDevice, s));
    CK(viso_launch_detect(b.job, 1, c, s));
    ctx->launches += 2;
    int n = 0;
    CK(cudaMemcpyAsync(&n, b.n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (n > 0) {
        CK(cudaMemcpyAsync(kp_xy, b.kp, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        if (kp_response) CK(cudaMemcpyAsync(kp_response, b.resp, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    *n_out = n;
    return VISO_OK;
}

int viso_debug_sincos(viso_ctx* ctx, const double* x, int n, double* s, double* c)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 0 || (n > 0 && (!x || !s || !c))) return ctx->fail(VISO_ERR_ARG, "debug_sincos: bad argument");
    if (n == 0) return VISO_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_scratch(ctx, (size_t)n * 24 + 256);
    if (rc) return rc;
    double* dx = reinterpret_cast<double*>(ctx->d_scr);
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(viso_launch_sincos_probe(dx, n, dx + n, dx + 2 * (size_t)n, st));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(s, dx + n, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c, dx + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return VISO_OK;
}

int viso_extract_descriptors(viso_ctx* ctx, const uint8_t* img, int width, int height, int pitch, const float* kp_xy, int n,
                             float* desc)
{
    if (!ctx) return VISO_ERR_ARG;
    if (n < 0 || width < 3 || height < 3 || pitch < width) return ctx->fail(VISO_ERR_ARG, "extract_descriptors: bad argument");
    if (n == 0) return VISO_OK;
    if (!img || !kp_xy || !desc) return ctx->fail(VISO_ERR_ARG, "extract_descriptors: null argument");
    CK(cudaSetDevice(ctx->device));
    struct Bufs { unsigned char* img; float2* kp; uint16_t* rows; unsigned* rsum; int *n, *flag; ExtractJob* job; } b;
    auto carve = [&](Carver& cv) {
        b.img = cv.take<unsigned char>((size_t)pitch * height);
        b.kp = cv.take<float2>(n); b.rows = cv.take<uint16_t>((size_t)n * VISO_DESC_U16); b.rsum = cv.take<unsigned>(n);
        b.n = cv.take<int>(1); b.flag = cv.take<int>(1); b.job = cv.take<ExtractJob>(1);
    };
    Carver measure(nullptr);
    carve(measure);
    int rc = ensure_scratch(ctx, measure.off);
    if (rc) return rc;
    Carver real(ctx->d_scr);
    carve(real);
    cudaStream_t s = ctx->stream;
    const int one = 1;
    const ExtractJob job{b.img, b.kp, b.n, b.rows, b.rsum, nullptr, b.flag};
    CK(cudaMemcpyAsync(b.img, img, (size_t)pitch * height, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.kp, kp_xy, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.n, &n, 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.flag, &one, 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b.job, &job, sizeof(job), cudaMemcpyHostToDevice, s));
    CK(viso_launch_extract(b.job, 1, n, width, height, pitch, 5, s));
    ctx->launches += 1;
    std::vector<uint16_t> rows((size_t)n * VISO_DESC_U16);
    CK(cudaMemcpyAsync(rows.data(), b.rows, rows.size() * 2, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    /* packed layout -> cv::Mat layout: elements are biased by 1024, 121 of the 128 are used */
    for (int k = 0; k < n; ++k)
        for (int c = 0; c < 121; ++c) desc[(size_t)k * 121 + c] = (float)((int)rows[(size_t)k * VISO_DESC_U16 + c] - 1024);
    return VISO_OK;
}

} /* extern "C" */
