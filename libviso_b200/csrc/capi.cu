/*
 * capi.cu -- the C-ABI of include/viso_b200.h: context, buffer management, job tables and launches.
 *
 * Host side only; every numeric result on the hot path is produced by the kernels in match.cu, sort_circle.cu, estimation.cu and geometry.cu.  There is no
 * CPU fallback: without a CUDA device viso_create() fails.  The only arithmetic done here is the once-per-sequence /
 * per-pose host bookkeeping the reference also keeps outside the per-frame loop (F_from_P, tr2mat, pose chaining)
 * and the RANSAC sample-table generators.
 */
#include "../../include/viso_b200.h"
#include "viso_dev.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

/* ------------------------------------------------------------------------------------------------ context */

struct viso_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    GridCfg grid{78, 24}; /* 1248 x 384 px in 16-px cells */
    char* d_scr = nullptr;
    size_t d_cap = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaStream_t copy_stream = nullptr;   /* host -> device uploads of sequence objects (overlap with compute) */

    int fail(int code, const std::string& msg)
    {
        err = msg;
        return code;
    }
    int fail_cuda(cudaError_t e, const char* what)
    {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return VISO_ERR_CUDA;
    }
    int ncell() const { return grid.gx * grid.gy; }
};

#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ctx->fail_cuda(e__, #call);      \
    } while (0)

namespace {

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

} // namespace

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
    *out = ctx;
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
    cudaStreamDestroy(ctx->copy_stream);
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
        unsigned *rs1, *rs2;
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
        b.rs1 = c.take<unsigned>(n1); b.rs2 = c.take<unsigned>(n2);
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
    PackJob pj[2] = {{b.df1, b.counts, b.du1, b.rs1, nullptr}, {b.df2, b.counts + 1, b.du2, b.rs2, nullptr}};
    GridJob gj[2] = {{b.xy1, b.counts, b.rs1, b.srec1, b.cell1}, {b.xy2, b.counts + 1, b.rs2, b.srec2, b.cell2}};
    MatchJob mj;
    mj.q = SetView{b.xy1, b.counts, b.du1, b.srec1, b.cell1};
    mj.t = SetView{b.xy2, b.counts + 1, b.du2, b.srec2, b.cell2};
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
    CK(viso_launch_pack(b.pack, 2, std::max(n1, n2), dlen, b.err, s));
    CK(viso_launch_grid(b.grid, 2, ctx->grid, s));
    int ml = 0;
    CK(viso_launch_match(b.match, 1, n1, n2, mp, ctx->grid, nullptr, b.pending, s, &ml));
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
    } b;
    auto carve = [&](Carver& c) {
        b.X = c.take<double>((size_t)n * 3); b.obs = c.take<double>((size_t)n * 4);
        b.hyp_tr = c.take<double>((size_t)H * 6); b.scratch = c.take<double>((size_t)n * 28);
        b.n = c.take<int>(1); b.table = c.take<int>((size_t)H * 3); b.hyp_ok = c.take<int>(H); b.hyp_count = c.take<int>(H);
        b.inliers = c.take<int>(n); b.active = c.take<int>(n);
        b.rec = c.take<viso_record_dev>(1); b.prob = c.take<RansacProb>(1);
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
    CK(viso_launch_ransac(b.prob, 1, H, n, make_param_dev(param), s, &nl));
    ctx->launches += nl;
    viso_record_dev rec;
    CK(cudaMemcpyAsync(&rec, b.rec, sizeof(rec), cudaMemcpyDeviceToHost, s));
    if (hyp_tr && H > 0) CK(cudaMemcpyAsync(hyp_tr, b.hyp_tr, (size_t)H * 48, cudaMemcpyDeviceToHost, s));
    if (hyp_ok && H > 0) CK(cudaMemcpyAsync(hyp_ok, b.hyp_ok, (size_t)H * 4, cudaMemcpyDeviceToHost, s));
    if (hyp_count && H > 0) CK(cudaMemcpyAsync(hyp_count, b.hyp_count, (size_t)H * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
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
        const double s = F[8];
        for (int i = 0; i < 9; i++) F[i] /= s;
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

/* ------------------------------------------------------------------------------------------------ sequence */

} /* extern "C" */

struct viso_seq {
    viso_ctx* ctx = nullptr;
    int F = 0, cap = 0, dlen = 0, maxH = 0, ncell = 0;
    GridCfg grid{};
    float2 *kpL = nullptr, *kpR = nullptr;
    uint4 *srecL = nullptr, *srecR = nullptr;
    unsigned *rsL = nullptr, *rsR = nullptr;
    float *dLf = nullptr, *dRf = nullptr;
    uint16_t *dLu = nullptr, *dRu = nullptr;
    int *nL = nullptr, *nR = nullptr, *cellL = nullptr, *cellR = nullptr;
    int4 *dense_lr = nullptr, *dense_11 = nullptr, *dense_22 = nullptr;
    int *lr = nullptr, *lr_count = nullptr, *pos = nullptr;
    double *x = nullptr, *X = nullptr, *x_c = nullptr, *Xp_c = nullptr;
    int *circ4 = nullptr, *pcl2 = nullptr, *n_circ = nullptr;
    double *hyp_tr = nullptr, *scratch = nullptr;
    int *hyp_ok = nullptr, *hyp_count = nullptr, *inliers = nullptr, *active = nullptr;
    viso_record_dev* rec = nullptr;
    uint32_t* seeds = nullptr;
    PackJob* pack_jobs = nullptr;
    GridJob* grid_jobs = nullptr;
    MatchJob* match_jobs = nullptr;
    SortJob* sort_jobs = nullptr;
    CircleJob* circ_jobs = nullptr;
    RansacProb* probs = nullptr;
    unsigned long long* pairs = nullptr;
    int* err = nullptr;
    int* pending = nullptr;
    int *h_nL = nullptr, *h_nR = nullptr, *h_from_image = nullptr; /* pinned: truly asynchronous count uploads */
    int* from_image = nullptr;            /* device [F]: frame t's descriptors come from its images */
    unsigned char *imgL = nullptr, *imgR = nullptr;
    int img_w = 0, img_h = 0;
    ExtractJob* extract_jobs = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_compute = nullptr;
    int run_hi = 0;                       /* frames [0, run_hi) may still be read by enqueued kernels */
    std::vector<RansacProb> h_probs;
    int H_cur = -1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_ms = false, calib_set = false, ran = false;
    double Fm[9]{}, base = 0, f = 0, cu = 0, cv = 0;
    std::vector<void*> allocs;
};

namespace {

template <class T> cudaError_t seq_alloc(viso_seq* s, T** p, size_t n)
{
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    s->allocs.push_back(v);
    *p = reinterpret_cast<T*>(v);
    return cudaSuccess;
}

void seq_free(viso_seq* s)
{
    for (void* p : s->allocs) cudaFree(p);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev_copy) cudaEventDestroy(s->ev_copy);
    if (s->ev_compute) cudaEventDestroy(s->ev_compute);
    if (s->h_nL) cudaFreeHost(s->h_nL);
    delete s;
}

} // namespace

extern "C" {

int viso_seq_create(viso_ctx* ctx, int n_frames, int max_kp, int desc_len, int max_ransac_iter, viso_seq** out)
{
    if (!ctx) return VISO_ERR_ARG;
    if (!out || n_frames < 1 || max_kp < 1 || max_ransac_iter < 0) return ctx->fail(VISO_ERR_ARG, "seq_create: bad argument");
    if (desc_len < 1 || desc_len > VISO_DESC_U16 - 2) return ctx->fail(VISO_ERR_DOMAIN, "seq_create: desc_len must be 1..126");
    if ((long long)3 * n_frames > 65535) return ctx->fail(VISO_ERR_ARG, "seq_create: at most 21845 frames per sequence object");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    viso_seq* s = new viso_seq();
    s->ctx = ctx; s->F = n_frames; s->cap = (max_kp + 31) & ~31; s->dlen = desc_len; s->maxH = max_ransac_iter;
    s->grid = ctx->grid; s->ncell = ctx->ncell();
    const size_t F = n_frames, cap = s->cap, H = std::max(max_ransac_iter, 1), nc = s->ncell + 1;
#define SA(ptr, count)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = seq_alloc(s, &s->ptr, (count));                                \
        if (e__ != cudaSuccess) {                                                        \
            seq_free(s);                                                                 \
            ctx->err = std::string("seq_create cudaMalloc: ") + cudaGetErrorString(e__); \
            return e__ == cudaErrorMemoryAllocation ? VISO_ERR_NOMEM : VISO_ERR_CUDA;    \
        }                                                                                \
    } while (0)
    SA(kpL, F * cap); SA(kpR, F * cap); SA(srecL, F * cap); SA(srecR, F * cap);
    SA(rsL, F * cap); SA(rsR, F * cap);
    SA(dLf, F * cap * desc_len); SA(dRf, F * cap * desc_len);
    SA(dLu, F * cap * VISO_DESC_U16); SA(dRu, F * cap * VISO_DESC_U16);
    SA(nL, F); SA(nR, F); SA(cellL, F * nc); SA(cellR, F * nc);
    SA(dense_lr, F * cap); SA(dense_11, F * cap); SA(dense_22, F * cap);
    SA(lr, F * cap * 3); SA(lr_count, F); SA(pos, F * cap);
    SA(x, F * cap * 4); SA(X, F * cap * 3); SA(x_c, F * cap * 4); SA(Xp_c, F * cap * 3);
    SA(circ4, F * cap * 4); SA(pcl2, F * cap * 2); SA(n_circ, F);
    SA(hyp_tr, F * H * 6); SA(scratch, F * cap * 28);
    SA(hyp_ok, F * H); SA(hyp_count, F * H); SA(inliers, F * cap); SA(active, F * cap);
    SA(rec, F); SA(seeds, F * H * 3);
    SA(pack_jobs, 2 * F); SA(grid_jobs, 2 * F); SA(match_jobs, 3 * F); SA(sort_jobs, F); SA(circ_jobs, F); SA(probs, F);
    SA(pairs, 2); SA(err, 1); SA(pending, 1); SA(from_image, F); SA(extract_jobs, 2 * F);
#undef SA
    if (cudaMallocHost(&s->h_nL, 3 * F * sizeof(int)) != cudaSuccess) {
        seq_free(s);
        return ctx->fail(VISO_ERR_NOMEM, "seq_create: cudaMallocHost failed");
    }
    s->h_nR = s->h_nL + F;
    s->h_from_image = s->h_nL + 2 * F;
    std::memset(s->h_nL, 0, 3 * F * sizeof(int));
    cudaStream_t st = ctx->stream;
    auto bail = [&](cudaError_t e, const char* what) { seq_free(s); return ctx->fail_cuda(e, what); };
    cudaError_t e;
    if ((e = cudaEventCreate(&s->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&s->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&s->ev_compute, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMemsetAsync(s->from_image, 0, F * 4, st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(s->nL, 0, F * 4, st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(s->nR, 0, F * 4, st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(s->lr_count, 0, F * 4, st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(s->n_circ, 0, F * 4, st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(s->rec, 0, F * sizeof(viso_record_dev), st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(s->pairs, 0, 16, st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");
    if ((e = cudaMemsetAsync(s->err, 0, 4, st)) != cudaSuccess) return bail(e, "cudaMemsetAsync");

    /* job tables: all pointers are fixed for the life of the object */
    std::vector<PackJob> pj(2 * F);
    std::vector<GridJob> gj(2 * F);
    std::vector<MatchJob> mj;
    std::vector<SortJob> sj(F);
    std::vector<CircleJob> cj(F);
    s->h_probs.resize(F);
    auto viewL = [&](size_t t) {
        return SetView{s->kpL + t * cap, s->nL + t, s->dLu + t * cap * VISO_DESC_U16, s->srecL + t * cap,
                       s->cellL + t * nc};
    };
    auto viewR = [&](size_t t) {
        return SetView{s->kpR + t * cap, s->nR + t, s->dRu + t * cap * VISO_DESC_U16, s->srecR + t * cap,
                       s->cellR + t * nc};
    };
    for (size_t t = 0; t < F; ++t) {
        pj[2 * t] = PackJob{s->dLf + t * cap * desc_len, s->nL + t, s->dLu + t * cap * VISO_DESC_U16, s->rsL + t * cap,
                            s->from_image + t};
        pj[2 * t + 1] = PackJob{s->dRf + t * cap * desc_len, s->nR + t, s->dRu + t * cap * VISO_DESC_U16, s->rsR + t * cap,
                                s->from_image + t};
        gj[2 * t] = GridJob{s->kpL + t * cap, s->nL + t, s->rsL + t * cap, s->srecL + t * cap, s->cellL + t * nc};
        gj[2 * t + 1] = GridJob{s->kpR + t * cap, s->nR + t, s->rsR + t * cap, s->srecR + t * cap, s->cellR + t * nc};
        MatchJob m;
        m.pad = 0;
        m.q = viewL(t); m.t = viewR(t); m.out = s->dense_lr + t * cap; m.mode = 0; /* stereo, viso.cpp:1240 */
        mj.push_back(m);
        if (t > 0) {
            m.q = viewL(t); m.t = viewL(t - 1); m.out = s->dense_11 + t * cap; m.mode = 1; /* viso.cpp:1264 */
            mj.push_back(m);
            m.q = viewR(t); m.t = viewR(t - 1); m.out = s->dense_22 + t * cap; m.mode = 1; /* viso.cpp:1275 */
            mj.push_back(m);
        }
        SortJob& so = sj[t];
        so.dense = s->dense_lr + t * cap; so.n = s->nL + t; so.kp1 = s->kpL + t * cap; so.kp2 = s->kpR + t * cap;
        so.matches = s->lr + t * cap * 3; so.count = s->lr_count + t; so.pos_of_query = s->pos + t * cap;
        so.x = s->x + t * cap * 4; so.X = s->X + t * cap * 3; so.stride = (int)cap; so.pad = 0;
        const size_t tp = t > 0 ? t - 1 : 0;
        CircleJob& c = cj[t];
        c.lr = s->lr + t * cap * 3; c.lr_count = s->lr_count + t;
        c.lrp = s->lr + tp * cap * 3; c.lrp_count = s->lr_count + tp;
        c.pos_prev = s->pos + tp * cap; c.n_prev_left = s->nL + tp;
        c.m11 = s->dense_11 + t * cap; c.m22 = s->dense_22 + t * cap;
        c.x = s->x + t * cap * 4; c.Xp = s->X + tp * cap * 3;
        c.circ4 = s->circ4 + t * cap * 4; c.pcl2 = s->pcl2 + t * cap * 2; c.n_circ = s->n_circ + t;
        c.x_c = s->x_c + t * cap * 4; c.Xp_c = s->Xp_c + t * cap * 3;
        c.rec = s->rec + t; c.stride = (int)cap; c.pad = 0;
        RansacProb& p = s->h_probs[t];
        std::memset(&p, 0, sizeof(p));
        p.X = s->Xp_c + t * cap * 3; p.obs = s->x_c + t * cap * 4; p.n = s->n_circ + t; p.stride = (int)cap;
        p.H = 0; p.seeds = nullptr; p.table = nullptr;
        p.hyp_tr = s->hyp_tr + t * H * 6; p.hyp_ok = s->hyp_ok + t * H; p.hyp_count = s->hyp_count + t * H;
        p.scratch = s->scratch + t * cap * 28; p.inliers = s->inliers + t * cap; p.active = s->active + t * cap;
        p.rec = s->rec + t; p.min_n = 3; /* viso.cpp:1283: fewer than 3 circular matches => frame skipped */
    }
    if ((e = cudaMemcpyAsync(s->pack_jobs, pj.data(), pj.size() * sizeof(PackJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "cudaMemcpyAsync");
    if ((e = cudaMemcpyAsync(s->grid_jobs, gj.data(), gj.size() * sizeof(GridJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "cudaMemcpyAsync");
    if ((e = cudaMemcpyAsync(s->match_jobs, mj.data(), mj.size() * sizeof(MatchJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "cudaMemcpyAsync");
    if ((e = cudaMemcpyAsync(s->sort_jobs, sj.data(), sj.size() * sizeof(SortJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "cudaMemcpyAsync");
    if ((e = cudaMemcpyAsync(s->circ_jobs, cj.data(), cj.size() * sizeof(CircleJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "cudaMemcpyAsync");
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail(e, "cudaStreamSynchronize");
    *out = s;
    return VISO_OK;
}

void viso_seq_destroy(viso_seq* s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    seq_free(s);
}

int viso_seq_set_calib(viso_seq* s, const double P1[12], const double P2[12])
{
    if (!s || !P1 || !P2) return VISO_ERR_ARG;
    viso_F_from_P(P1, P2, 1, s->Fm);      /* viso.cpp:1176-1180 */
    s->base = std::fabs(P2[3] / P2[0]);   /* :1184 */
    s->f = P1[0]; s->cu = P1[2]; s->cv = P1[6]; /* :1185-1187 */
    s->calib_set = true;
    return VISO_OK;
}

/* an upload that overwrites a frame enqueued kernels may still read has to wait for them (not for the others: that
 * is what lets the uploads of one chunk overlap the kernels of the previous one) */
static int upload_guard(viso_seq* s, int t)
{
    viso_ctx* ctx = s->ctx;
    if (t < s->run_hi) {
        CK(cudaStreamWaitEvent(ctx->copy_stream, s->ev_compute, 0));
        s->run_hi = 0; /* everything enqueued so far is now ordered before later uploads */
    }
    return VISO_OK;
}

int viso_seq_upload_frame(viso_seq* s, int t, const float* kpL, int nL, const float* kpR, int nR, const float* dL,
                          const float* dR)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || nL < 0 || nR < 0 || nL > s->cap || nR > s->cap) return ctx->fail(VISO_ERR_ARG, "seq_upload_frame: bad frame index or keypoint count");
    if ((nL > 0 && (!kpL || !dL)) || (nR > 0 && (!kpR || !dR))) return ctx->fail(VISO_ERR_ARG, "seq_upload_frame: null input");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_guard(s, t);
    if (rc) return rc;
    cudaStream_t st = ctx->copy_stream;
    const size_t cap = s->cap, dl = s->dlen;
    if (nL > 0) {
        CK(cudaMemcpyAsync(s->kpL + t * cap, kpL, (size_t)nL * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->dLf + t * cap * dl, dL, (size_t)nL * dl * 4, cudaMemcpyHostToDevice, st));
    }
    if (nR > 0) {
        CK(cudaMemcpyAsync(s->kpR + t * cap, kpR, (size_t)nR * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->dRf + t * cap * dl, dR, (size_t)nR * dl * 4, cudaMemcpyHostToDevice, st));
    }
    s->h_nL[t] = nL;
    s->h_nR[t] = nR;
    s->h_from_image[t] = 0;
    return VISO_OK;
}

int viso_seq_set_image_size(viso_seq* s, int width, int height)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (width < 3 || height < 3) return ctx->fail(VISO_ERR_ARG, "seq_set_image_size: images must be at least 3 x 3");
    if (s->dlen != 121) return ctx->fail(VISO_ERR_DOMAIN, "seq_set_image_size: the device extractor produces 11 x 11 descriptors (desc_len 121)");
    if (s->imgL) return (width == s->img_w && height == s->img_h) ? VISO_OK : ctx->fail(VISO_ERR_ARG, "seq_set_image_size: size already set");
    CK(cudaSetDevice(ctx->device));
    const size_t F = s->F, cap = s->cap, bytes = (size_t)width * height;
    cudaError_t e;
    /* one allocation, [frame][left, right][height][width]: a run of frames is one contiguous block */
    if ((e = seq_alloc(s, &s->imgL, 2 * F * bytes)) != cudaSuccess) {
        ctx->err = std::string("seq_set_image_size cudaMalloc: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? VISO_ERR_NOMEM : VISO_ERR_CUDA;
    }
    s->imgR = s->imgL + bytes;
    s->img_w = width; s->img_h = height;
    std::vector<ExtractJob> ej(2 * F);
    for (size_t t = 0; t < F; ++t) {
        ej[2 * t] = ExtractJob{s->imgL + 2 * t * bytes, s->kpL + t * cap, s->nL + t, s->dLu + t * cap * VISO_DESC_U16,
                               s->rsL + t * cap, s->from_image + t};
        ej[2 * t + 1] = ExtractJob{s->imgR + 2 * t * bytes, s->kpR + t * cap, s->nR + t, s->dRu + t * cap * VISO_DESC_U16,
                                   s->rsR + t * cap, s->from_image + t};
    }
    CK(cudaMemcpyAsync(s->extract_jobs, ej.data(), ej.size() * sizeof(ExtractJob), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VISO_OK;
}

int viso_seq_upload_frame_images(viso_seq* s, int t, const uint8_t* imgL, const uint8_t* imgR, const float* kpL, int nL,
                                 const float* kpR, int nR)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!s->imgL) return ctx->fail(VISO_ERR_ARG, "seq_upload_frame_images: viso_seq_set_image_size has not been called");
    if (t < 0 || t >= s->F || nL < 0 || nR < 0 || nL > s->cap || nR > s->cap) return ctx->fail(VISO_ERR_ARG, "seq_upload_frame_images: bad frame index or keypoint count");
    if (!imgL || !imgR || (nL > 0 && !kpL) || (nR > 0 && !kpR)) return ctx->fail(VISO_ERR_ARG, "seq_upload_frame_images: null input");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_guard(s, t);
    if (rc) return rc;
    cudaStream_t st = ctx->copy_stream;
    const size_t cap = s->cap, bytes = (size_t)s->img_w * s->img_h;
    if (imgR == imgL + bytes) {
        CK(cudaMemcpyAsync(s->imgL + 2 * t * bytes, imgL, 2 * bytes, cudaMemcpyHostToDevice, st));
    } else {
        CK(cudaMemcpyAsync(s->imgL + 2 * t * bytes, imgL, bytes, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->imgR + 2 * t * bytes, imgR, bytes, cudaMemcpyHostToDevice, st));
    }
    if (nL > 0) CK(cudaMemcpyAsync(s->kpL + t * cap, kpL, (size_t)nL * 8, cudaMemcpyHostToDevice, st));
    if (nR > 0) CK(cudaMemcpyAsync(s->kpR + t * cap, kpR, (size_t)nR * 8, cudaMemcpyHostToDevice, st));
    s->h_nL[t] = nL;
    s->h_nR[t] = nR;
    s->h_from_image[t] = 1;
    return VISO_OK;
}

int viso_seq_capacity(const viso_seq* s) { return s ? s->cap : 0; }

int viso_seq_upload_chunk_images(viso_seq* s, int t0, int count, const uint8_t* images, const float* kpL, const int32_t* nL,
                                 const float* kpR, const int32_t* nR)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!s->imgL) return ctx->fail(VISO_ERR_ARG, "seq_upload_chunk_images: viso_seq_set_image_size has not been called");
    if (t0 < 0 || count < 1 || t0 + count > s->F || !images || !kpL || !kpR || !nL || !nR)
        return ctx->fail(VISO_ERR_ARG, "seq_upload_chunk_images: bad argument");
    for (int i = 0; i < count; ++i)
        if (nL[i] < 0 || nR[i] < 0 || nL[i] > s->cap || nR[i] > s->cap) return ctx->fail(VISO_ERR_ARG, "seq_upload_chunk_images: bad keypoint count");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_guard(s, t0);
    if (rc) return rc;
    cudaStream_t st = ctx->copy_stream;
    const size_t cap = s->cap, bytes = (size_t)s->img_w * s->img_h;
    CK(cudaMemcpyAsync(s->imgL + 2 * (size_t)t0 * bytes, images, 2 * (size_t)count * bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->kpL + (size_t)t0 * cap, kpL, (size_t)count * cap * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->kpR + (size_t)t0 * cap, kpR, (size_t)count * cap * 8, cudaMemcpyHostToDevice, st));
    for (int i = 0; i < count; ++i) {
        s->h_nL[t0 + i] = nL[i];
        s->h_nR[t0 + i] = nR[i];
        s->h_from_image[t0 + i] = 1;
    }
    return VISO_OK;
}

int viso_seq_set_seeds(viso_seq* s, const uint32_t* seeds, int ransac_iter)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (ransac_iter < 0 || ransac_iter > s->maxH || (ransac_iter > 0 && !seeds)) return ctx->fail(VISO_ERR_ARG, "seq_set_seeds: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_guard(s, 0);
    if (rc) return rc;
    cudaStream_t st = ctx->copy_stream;
    const size_t H = ransac_iter;
    if (H > 0) CK(cudaMemcpyAsync(s->seeds, seeds, (size_t)s->F * H * 12, cudaMemcpyHostToDevice, st));
    if (s->H_cur != ransac_iter) {
        for (int t = 0; t < s->F; ++t) {
            s->h_probs[t].H = ransac_iter;
            s->h_probs[t].seeds = s->seeds + (size_t)t * H * 3;
        }
        CK(cudaStreamSynchronize(st)); /* h_probs is pageable: keep the copy ordered with later edits */
        CK(cudaMemcpyAsync(s->probs, s->h_probs.data(), (size_t)s->F * sizeof(RansacProb), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        s->H_cur = ransac_iter;
    }
    return VISO_OK;
}

int viso_seq_run_range(viso_seq* s, const viso_param* param, int t0, int t1)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!param) return ctx->fail(VISO_ERR_ARG, "seq_run: null param");
    if (t0 < 0 || t1 > s->F || t0 >= t1) return ctx->fail(VISO_ERR_ARG, "seq_run_range: bad frame range");
    if (!s->calib_set) return ctx->fail(VISO_ERR_ARG, "seq_run: viso_seq_set_calib has not been called");
    if (s->H_cur != param->ransac_iter) return ctx->fail(VISO_ERR_ARG, "seq_run: seeds were set for a different ransac_iter");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    /* everything uploaded so far (frames, seeds) is visible to the kernels below */
    CK(cudaEventRecord(s->ev_copy, ctx->copy_stream));
    CK(cudaStreamWaitEvent(st, s->ev_copy, 0));
    const int nf = t1 - t0;
    CK(cudaMemcpyAsync(s->nL + t0, s->h_nL + t0, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->nR + t0, s->h_nR + t0, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->from_image + t0, s->h_from_image + t0, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    if (t0 == 0) {
        CK(cudaMemsetAsync(s->pairs, 0, 16, st));
        CK(cudaMemsetAsync(s->err, 0, 4, st));
    }
    int max_n = 0, max_nL = 0, any_img = 0, any_f32 = 0;
    for (int t = t0; t < t1; ++t) {
        max_n = std::max(max_n, std::max(s->h_nL[t], s->h_nR[t]));
        max_nL = std::max(max_nL, s->h_nL[t]);
        if (s->h_from_image[t]) any_img = 1; else any_f32 = 1;
    }
    int max_nt = max_n; /* the temporal targets of frame t0 live in frame t0 - 1 */
    if (t0 > 0) max_nt = std::max(max_nt, std::max(s->h_nL[t0 - 1], s->h_nR[t0 - 1]));
    viso_param pp = *param;
    pp.base = s->base; pp.f = s->f; pp.cu = s->cu; pp.cv = s->cv;
    const ParamDev pd = make_param_dev(&pp);
    viso_match_params ms, mt;
    viso_match_params_stereo(&ms, s->Fm);
    viso_match_params_temporal(&mt);
    MatchParamsPair mp;
    mp.p[0] = make_match_dev(&ms);
    mp.p[1] = make_match_dev(&mt);

    int nl = 0;
    if (any_f32 && max_n > 0) { CK(viso_launch_pack(s->pack_jobs + 2 * t0, 2 * nf, max_n, s->dlen, s->err, st)); ++nl; }
    if (any_img && max_n > 0) {
        CK(viso_launch_extract(s->extract_jobs + 2 * t0, 2 * nf, max_n, s->img_w, s->img_h, s->img_w, 5, st));
        ++nl;
    }
    CK(viso_launch_grid(s->grid_jobs + 2 * t0, 2 * nf, s->grid, st));
    ++nl;
    /* match jobs: frame 0 has one (stereo), frame t >= 1 has three (stereo, temporal L, temporal R) */
    const int mj0 = t0 == 0 ? 0 : 3 * t0 - 2, mj1 = 3 * t1 - 2;
    CK(cudaEventRecord(s->ev0, st));
    CK(viso_launch_match(s->match_jobs + mj0, mj1 - mj0, max_n, max_nt, mp, s->grid, s->pairs, s->pending, st, &nl));
    CK(cudaEventRecord(s->ev1, st));
    CK(viso_launch_sort(s->sort_jobs + t0, nf, max_nL, pd, st));
    ++nl;
    const int c0 = std::max(t0, 1);
    if (t1 > c0) {
        CK(viso_launch_circle(s->circ_jobs + c0, t1 - c0, st));
        ++nl;
        CK(viso_launch_ransac(s->probs + c0, t1 - c0, param->ransac_iter, max_nL, pd, st, &nl));
    }
    ctx->launches += nl;
    CK(cudaEventRecord(s->ev_compute, st));
    s->run_hi = std::max(s->run_hi, t1);
    s->have_ms = max_n > 0;
    s->ran = true;
    return VISO_OK;
}

int viso_seq_run_resident(viso_seq* s, const viso_param* param)
{
    if (!s) return VISO_ERR_ARG;
    return viso_seq_run_range(s, param, 0, s->F);
}

int viso_seq_run(viso_seq* s, const viso_param* param, const uint32_t* seeds)
{
    if (!s) return VISO_ERR_ARG;
    if (!param) return s->ctx->fail(VISO_ERR_ARG, "seq_run: null param");
    int rc = viso_seq_set_seeds(s, seeds, param->ransac_iter);
    if (rc) return rc;
    return viso_seq_run_resident(s, param);
}

int viso_seq_download(viso_seq* s, viso_record* records)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!records) return ctx->fail(VISO_ERR_ARG, "seq_download: null output");
    if (!s->ran) return ctx->fail(VISO_ERR_ARG, "seq_download: nothing has been run");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    static_assert(sizeof(viso_record) == sizeof(viso_record_dev), "record layout");
    int flags = 0;
    CK(cudaMemcpyAsync(records, s->rec, (size_t)s->F * sizeof(viso_record), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&flags, s->err, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    s->run_hi = 0;
    std::memset(&records[0], 0, sizeof(viso_record)); /* first frame: no pose (viso.cpp:1256-1260) */
    records[0].best_hyp = -1;
    return status_from_flags(ctx, flags);
}

int viso_seq_stats(viso_seq* s, int64_t* match_bytes, int64_t* sad_pairs, int64_t* sad_evaluated)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    if (match_bytes) {
        /* SURVEY 8d, per frame pair, fused, u16 layout: the four descriptor sets (rows of 256 B + 8 B of
         * coordinates) read once and three dense int4 outputs written */
        int64_t b = 0;
        for (int t = 1; t < s->F; ++t) {
            const int64_t nl = s->h_nL[t], nr = s->h_nR[t], nlp = s->h_nL[t - 1], nrp = s->h_nR[t - 1];
            b += (nl + nr + nlp + nrp) * (VISO_DESC_U16 * 2 + 8) + 16 * (2 * nl + nr);
        }
        *match_bytes = b;
    }
    if (sad_pairs || sad_evaluated) {
        unsigned long long p[2] = {0, 0};
        CK(cudaMemcpyAsync(p, s->pairs, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (sad_pairs) *sad_pairs = (int64_t)p[0];
        if (sad_evaluated) *sad_evaluated = (int64_t)p[1];
    }
    return VISO_OK;
}

int viso_seq_match_ms(viso_seq* s, float* ms)
{
    if (!s || !ms) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!s->have_ms) return ctx->fail(VISO_ERR_ARG, "seq_match_ms: no timed run");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(s->ev1));
    CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return VISO_OK;
}

/* ---- parity-test getters ---- */

int viso_seq_get_dense(viso_seq* s, int which, int t, int32_t* out4, int32_t* n)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || which < 0 || which > 2 || !out4 || !n) return ctx->fail(VISO_ERR_ARG, "seq_get_dense: bad argument");
    CK(cudaSetDevice(ctx->device));
    const int cnt = which == 2 ? s->h_nR[t] : s->h_nL[t];
    const int4* src = (which == 0 ? s->dense_lr : which == 1 ? s->dense_11 : s->dense_22) + (size_t)t * s->cap;
    *n = cnt;
    if (cnt > 0) CK(cudaMemcpyAsync(out4, src, (size_t)cnt * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VISO_OK;
}

int viso_seq_get_packed(viso_seq* s, int t, int side, uint16_t* rows, int32_t* n)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || side < 0 || side > 1 || !rows || !n) return ctx->fail(VISO_ERR_ARG, "seq_get_packed: bad argument");
    CK(cudaSetDevice(ctx->device));
    const int cnt = side ? s->h_nR[t] : s->h_nL[t];
    const uint16_t* src = (side ? s->dRu : s->dLu) + (size_t)t * s->cap * VISO_DESC_U16;
    *n = cnt;
    if (cnt > 0) CK(cudaMemcpyAsync(rows, src, (size_t)cnt * VISO_DESC_U16 * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VISO_OK;
}

int viso_seq_get_lr_matches(viso_seq* s, int t, int32_t* matches3, int32_t* n)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || !matches3 || !n) return ctx->fail(VISO_ERR_ARG, "seq_get_lr_matches: bad argument");
    CK(cudaSetDevice(ctx->device));
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, s->lr_count + t, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (cnt > 0) CK(cudaMemcpyAsync(matches3, s->lr + (size_t)t * s->cap * 3, (size_t)cnt * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    return VISO_OK;
}

int viso_seq_get_circ(viso_seq* s, int t, int32_t* circ4, int32_t* pcl2, int32_t* n)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || !n) return ctx->fail(VISO_ERR_ARG, "seq_get_circ: bad argument");
    CK(cudaSetDevice(ctx->device));
    int cnt = 0;
    if (t > 0) {
        CK(cudaMemcpyAsync(&cnt, s->n_circ + t, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (cnt > 0 && circ4) CK(cudaMemcpyAsync(circ4, s->circ4 + (size_t)t * s->cap * 4, (size_t)cnt * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (cnt > 0 && pcl2) CK(cudaMemcpyAsync(pcl2, s->pcl2 + (size_t)t * s->cap * 2, (size_t)cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    return VISO_OK;
}

int viso_seq_get_inliers(viso_seq* s, int t, int32_t* inliers, int32_t* n)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || !n) return ctx->fail(VISO_ERR_ARG, "seq_get_inliers: bad argument");
    CK(cudaSetDevice(ctx->device));
    int cnt = 0;
    if (t > 0) {
        viso_record_dev r;
        CK(cudaMemcpyAsync(&r, s->rec + t, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        cnt = r.n_inliers;
    }
    if (cnt > 0 && inliers) CK(cudaMemcpyAsync(inliers, s->inliers + (size_t)t * s->cap, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    return VISO_OK;
}

int viso_seq_get_hyp(viso_seq* s, int t, double* hyp_tr, int32_t* hyp_ok, int32_t* hyp_count)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 1 || t >= s->F || s->H_cur < 1) return ctx->fail(VISO_ERR_ARG, "seq_get_hyp: bad argument");
    CK(cudaSetDevice(ctx->device));
    const size_t H = s->H_cur, Hs = std::max(s->maxH, 1);
    if (hyp_tr) CK(cudaMemcpyAsync(hyp_tr, s->hyp_tr + (size_t)t * Hs * 6, H * 48, cudaMemcpyDeviceToHost, ctx->stream));
    if (hyp_ok) CK(cudaMemcpyAsync(hyp_ok, s->hyp_ok + (size_t)t * Hs, H * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (hyp_count) CK(cudaMemcpyAsync(hyp_count, s->hyp_count + (size_t)t * Hs, H * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VISO_OK;
}

} /* extern "C" */
