/*
 * capi_seq.cu -- the batched sequence pipeline of include/viso_b200.h (viso_seq_*): the per-frame loop of
 * sequence_odometry (reference src/viso.cpp:1205-1327) for all frames of a sequence, job tables built once, one
 * launch of each kernel per submission, uploads on a copy stream.  Host side only.
 */
#include "capi_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

using namespace viso_capi;

#define VISO_SEQ_PENDING_CAP 65536

struct viso_seq {
    viso_ctx* ctx = nullptr;
    int F = 0, cap = 0, dlen = 0, maxH = 0, ncell = 0;
    GridCfg grid{};
    float2 *kpL = nullptr, *kpR = nullptr;
    uint4 *srecL = nullptr, *srecR = nullptr;
    int *posL = nullptr, *posR = nullptr;   /* original index -> cell-sorted position */
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
    int* strag = nullptr;   /* hypotheses handed from the quad kernel to the warp-per-hypothesis kernel (estimation.cu) */
    PackJob* pack_jobs = nullptr;
    GridJob* grid_jobs = nullptr;
    MatchJob* match_jobs = nullptr;
    MatchJob* match_jobs_stereo = nullptr;   /* [F] the stereo jobs alone, [2(F-1)] the temporal jobs alone: viso_seq_time_match */
    MatchJob* match_jobs_temporal = nullptr;
    SortJob* sort_jobs = nullptr;
    CircleJob* circ_jobs = nullptr;
    RansacProb* probs = nullptr;
    unsigned long long* pairs = nullptr;
    int* err = nullptr;
    int* pending = nullptr;
    uint4* pend_rec = nullptr;            /* the queries the tile kernel left to the generic kernel (first 64 Ki) */
    int* pend_job = nullptr;
    int *h_nL = nullptr, *h_nR = nullptr, *h_from_image = nullptr; /* pinned: truly asynchronous count uploads */
    int* from_image = nullptr;            /* device [F]: frame t's descriptors come from its images */
    unsigned char *imgL = nullptr, *imgR = nullptr;
    int img_w = 0, img_h = 0;
    ExtractJob* extract_jobs = nullptr;
    /* device detector (viso_seq_set_detector) */
    DetectJob* detect_jobs = nullptr;
    float2* kp_tmp = nullptr;
    int *bin_count = nullptr, *detect = nullptr, *h_detect = nullptr;
    HarrisCfg hc{};
    bool det_set = false;
    int det_n = 0;                        /* keypoints per image at most: per * nbins */
    cudaEvent_t ev_copy = nullptr, ev_compute = nullptr, ev_counts = nullptr;
    int counts_hi = 0;                    /* frames [0, counts_hi): their host-side counts may still be read by an enqueued copy */
    int run_hi = 0;                       /* frames [0, run_hi) may still be read by enqueued kernels */
    int run_hi_total = 0;                 /* frames [0, run_hi_total) have been through the pipeline at least once */
    std::vector<RansacProb> h_probs;
    int H_cur = -1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool have_ms = false, calib_set = false, ran = false;
    double Fm[9]{}, base = 0, f = 0, cu = 0, cv = 0;
    std::vector<void*> allocs;
    size_t bytes = 0;                     /* device memory held by this object */
};

namespace {

template <class T> cudaError_t seq_alloc(viso_seq* s, T** p, size_t n)
{
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    s->allocs.push_back(v);
    s->bytes += std::max<size_t>(n, 1) * sizeof(T);
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
    if (s->ev_counts) cudaEventDestroy(s->ev_counts);
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
    SA(posL, F * cap); SA(posR, F * cap);
    SA(dLu, F * cap * VISO_DESC_U16); SA(dRu, F * cap * VISO_DESC_U16);
    SA(nL, F); SA(nR, F); SA(cellL, F * nc); SA(cellR, F * nc);
    SA(dense_lr, F * cap); SA(dense_11, F * cap); SA(dense_22, F * cap);
    SA(lr, F * cap * 3); SA(lr_count, F); SA(pos, F * cap);
    SA(x, F * cap * 4); SA(X, F * cap * 3); SA(x_c, F * cap * 4); SA(Xp_c, F * cap * 3);
    SA(circ4, F * cap * 4); SA(pcl2, F * cap * 2); SA(n_circ, F);
    SA(hyp_tr, F * H * 6); SA(scratch, F * cap * 28);
    SA(hyp_ok, F * H); SA(hyp_count, F * H); SA(inliers, F * cap); SA(active, F * cap);
    SA(rec, F); SA(seeds, F * H * 3);
    SA(pack_jobs, 2 * F); SA(grid_jobs, 2 * F); SA(match_jobs, 3 * F); SA(match_jobs_stereo, F); SA(match_jobs_temporal, 2 * F); SA(sort_jobs, F); SA(circ_jobs, F); SA(probs, F);
    SA(strag, 2 + 2 * F * (size_t)std::max(max_ransac_iter, 1));
    SA(pairs, 2); SA(err, 1); SA(pending, 1); SA(pend_rec, VISO_SEQ_PENDING_CAP); SA(pend_job, VISO_SEQ_PENDING_CAP); SA(from_image, F); SA(extract_jobs, 2 * F);
#undef SA
    if (cudaMallocHost(&s->h_nL, 4 * F * sizeof(int)) != cudaSuccess) {
        seq_free(s);
        return ctx->fail(VISO_ERR_NOMEM, "seq_create: cudaMallocHost failed");
    }
    s->h_nR = s->h_nL + F;
    s->h_from_image = s->h_nL + 2 * F;
    s->h_detect = s->h_nL + 3 * F;
    std::memset(s->h_nL, 0, 4 * F * sizeof(int));
    cudaStream_t st = ctx->stream;
    auto bail = [&](cudaError_t e, const char* what) { seq_free(s); return ctx->fail_cuda(e, what); };
    cudaError_t e;
    if ((e = cudaEventCreate(&s->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&s->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&s->ev_compute, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&s->ev_counts, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
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
    std::vector<MatchJob> mj, mjs, mjt;
    std::vector<SortJob> sj(F);
    std::vector<CircleJob> cj(F);
    s->h_probs.resize(F);
    auto viewL = [&](size_t t) {
        return SetView{s->kpL + t * cap, s->nL + t, s->dLu + t * cap * VISO_DESC_U16, s->srecL + t * cap,
                       s->cellL + t * nc, s->posL + t * cap};
    };
    auto viewR = [&](size_t t) {
        return SetView{s->kpR + t * cap, s->nR + t, s->dRu + t * cap * VISO_DESC_U16, s->srecR + t * cap,
                       s->cellR + t * nc, s->posR + t * cap};
    };
    for (size_t t = 0; t < F; ++t) {
        /* the f32 descriptor staging (2 x F x cap x desc_len floats: 2 GB for a 1000-frame sequence) is only allocated when
         * descriptors are uploaded as cv::Mat rows (viso_seq_upload_frame); d is filled in then */
        pj[2 * t] = PackJob{nullptr, s->nL + t, s->dLu + t * cap * VISO_DESC_U16, s->srecL + t * cap,
                            s->from_image + t};
        pj[2 * t + 1] = PackJob{nullptr, s->nR + t, s->dRu + t * cap * VISO_DESC_U16, s->srecR + t * cap,
                                s->from_image + t};
        gj[2 * t] = GridJob{s->kpL + t * cap, s->nL + t, s->posL + t * cap, s->srecL + t * cap, s->cellL + t * nc};
        gj[2 * t + 1] = GridJob{s->kpR + t * cap, s->nR + t, s->posR + t * cap, s->srecR + t * cap, s->cellR + t * nc};
        MatchJob m;
        m.pad = 0;
        m.q = viewL(t); m.t = viewR(t); m.out = s->dense_lr + t * cap; m.mode = 0; /* stereo, viso.cpp:1240 */
        mj.push_back(m); mjs.push_back(m);
        if (t > 0) {
            m.q = viewL(t); m.t = viewL(t - 1); m.out = s->dense_11 + t * cap; m.mode = 1; /* viso.cpp:1264 */
            mj.push_back(m); mjt.push_back(m);
            m.q = viewR(t); m.t = viewR(t - 1); m.out = s->dense_22 + t * cap; m.mode = 1; /* viso.cpp:1275 */
            mj.push_back(m); mjt.push_back(m);
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
    if ((e = cudaMemcpyAsync(s->match_jobs_stereo, mjs.data(), mjs.size() * sizeof(MatchJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "cudaMemcpyAsync");
    if (!mjt.empty() && (e = cudaMemcpyAsync(s->match_jobs_temporal, mjt.data(), mjt.size() * sizeof(MatchJob), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "cudaMemcpyAsync");
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
static int upload_guard(viso_seq* s, int t, bool frame_data = true)
{
    viso_ctx* ctx = s->ctx;
    if (t < s->counts_hi) {
        /* the per-frame counts live in pinned host words that run_range copies asynchronously: the copy reads them
         * when it executes, so they may not be overwritten before it has */
        CK(cudaEventSynchronize(s->ev_counts));
        s->counts_hi = 0;
    }
    if (t < s->run_hi) {
        CK(cudaStreamWaitEvent(ctx->copy_stream, s->ev_compute, 0));
        s->run_hi = 0; /* everything enqueued so far is now ordered before later uploads */
    }
    if (frame_data && t < s->run_hi_total) s->run_hi_total = t; /* frames from t on have to be run again before a later range may start */
    return VISO_OK;
}

/* first f32-descriptor upload: allocate the staging arrays and point the pack jobs at them */
static int ensure_f32_staging(viso_seq* s)
{
    viso_ctx* ctx = s->ctx;
    if (s->dLf) return VISO_OK;
    const size_t F = s->F, cap = s->cap, dl = s->dlen;
    cudaError_t e;
    if ((e = seq_alloc(s, &s->dLf, F * cap * dl)) != cudaSuccess || (e = seq_alloc(s, &s->dRf, F * cap * dl)) != cudaSuccess) {
        s->dLf = nullptr;
        ctx->err = std::string("seq_upload_frame cudaMalloc: ") + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? VISO_ERR_NOMEM : VISO_ERR_CUDA;
    }
    std::vector<PackJob> pj(2 * F);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(pj.data(), s->pack_jobs, pj.size() * sizeof(PackJob), cudaMemcpyDeviceToHost));
    for (size_t t = 0; t < F; ++t) {
        pj[2 * t].d = s->dLf + t * cap * dl;
        pj[2 * t + 1].d = s->dRf + t * cap * dl;
    }
    CK(cudaMemcpy(s->pack_jobs, pj.data(), pj.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
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
    int rc = ensure_f32_staging(s);
    if (rc) return rc;
    rc = upload_guard(s, t);
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
    s->h_detect[t] = 0;
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
                               nullptr, s->srecL + t * cap, s->from_image + t};
        ej[2 * t + 1] = ExtractJob{s->imgR + 2 * t * bytes, s->kpR + t * cap, s->nR + t, s->dRu + t * cap * VISO_DESC_U16,
                                   nullptr, s->srecR + t * cap, s->from_image + t};
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
    s->h_detect[t] = 0;
    return VISO_OK;
}

int viso_seq_capacity(const viso_seq* s) { return s ? s->cap : 0; }

void* viso_seq_records_device(viso_seq* s) { return s ? (void*)s->rec : nullptr; }

int64_t viso_seq_device_bytes(const viso_seq* s) { return s ? (int64_t)s->bytes : 0; }

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
    /* pieces of ~16 MB: measured 54 GB/s against 36 GB/s for one 1 GB copy (tools/h2d_probe.py) */
    const size_t total = 2 * (size_t)count * bytes, piece = (size_t)16 << 20;
    for (size_t off = 0; off < total; off += piece)
        CK(cudaMemcpyAsync(s->imgL + 2 * (size_t)t0 * bytes + off, images + off, std::min(piece, total - off),
                           cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->kpL + (size_t)t0 * cap, kpL, (size_t)count * cap * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->kpR + (size_t)t0 * cap, kpR, (size_t)count * cap * 8, cudaMemcpyHostToDevice, st));
    for (int i = 0; i < count; ++i) {
        s->h_nL[t0 + i] = nL[i];
        s->h_nR[t0 + i] = nR[i];
        s->h_from_image[t0 + i] = 1;
        s->h_detect[t0 + i] = 0;
    }
    return VISO_OK;
}

int viso_seq_set_detector(viso_seq* s, int n_features, int nbinx, int nbiny, float k)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!s->imgL) return ctx->fail(VISO_ERR_ARG, "seq_set_detector: viso_seq_set_image_size has not been called");
    HarrisCfg c;
    int rc = make_harris_cfg(ctx, s->img_w, s->img_h, s->img_w, n_features, nbinx, nbiny, k, &c);
    if (rc) return rc;
    const size_t nb = (size_t)nbinx * nbiny, slots = nb * c.per, F = s->F, cap = s->cap, bytes = (size_t)s->img_w * s->img_h;
    if (slots < 1 || slots > cap) return ctx->fail(VISO_ERR_ARG, "seq_set_detector: n_features must be between the bin count and the sequence's max_kp");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (!s->det_set || nb * c.per != (size_t)s->det_n || c.nbinx != s->hc.nbinx || c.nbiny != s->hc.nbiny) {
        if (s->det_set) return ctx->fail(VISO_ERR_ARG, "seq_set_detector: the bin layout of a sequence object is fixed once set");
        cudaError_t e;
        if ((e = seq_alloc(s, &s->kp_tmp, 2 * F * slots)) != cudaSuccess || (e = seq_alloc(s, &s->bin_count, 2 * F * nb)) != cudaSuccess ||
            (e = seq_alloc(s, &s->detect, F)) != cudaSuccess || (e = seq_alloc(s, &s->detect_jobs, 2 * F)) != cudaSuccess) {
            ctx->err = std::string("seq_set_detector cudaMalloc: ") + cudaGetErrorString(e);
            return e == cudaErrorMemoryAllocation ? VISO_ERR_NOMEM : VISO_ERR_CUDA;
        }
        std::vector<DetectJob> dj(2 * F);
        for (size_t t = 0; t < F; ++t) {
            dj[2 * t] = DetectJob{s->imgL + 2 * t * bytes, s->kpL + t * cap, s->nL + t, s->kp_tmp + 2 * t * slots, nullptr, nullptr,
                                  s->bin_count + 2 * t * nb, s->detect + t};
            dj[2 * t + 1] = DetectJob{s->imgR + 2 * t * bytes, s->kpR + t * cap, s->nR + t, s->kp_tmp + (2 * t + 1) * slots, nullptr,
                                      nullptr, s->bin_count + (2 * t + 1) * nb, s->detect + t};
        }
        CK(cudaMemcpyAsync(s->detect_jobs, dj.data(), dj.size() * sizeof(DetectJob), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(s->detect, 0, F * 4, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    s->hc = c;
    s->det_n = (int)slots;
    s->det_set = true;
    return VISO_OK;
}

int viso_seq_upload_frame_raw(viso_seq* s, int t, const uint8_t* imgL, const uint8_t* imgR)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!s->det_set) return ctx->fail(VISO_ERR_ARG, "seq_upload_frame_raw: viso_seq_set_detector has not been called");
    if (t < 0 || t >= s->F || !imgL || !imgR) return ctx->fail(VISO_ERR_ARG, "seq_upload_frame_raw: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_guard(s, t);
    if (rc) return rc;
    cudaStream_t st = ctx->copy_stream;
    const size_t bytes = (size_t)s->img_w * s->img_h;
    CK(cudaMemcpyAsync(s->imgL + 2 * t * bytes, imgL, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->imgR + 2 * t * bytes, imgR, bytes, cudaMemcpyHostToDevice, st));
    s->h_nL[t] = s->h_nR[t] = s->det_n;   /* upper bound for launch sizing; the kernels read the device-side counts */
    s->h_from_image[t] = 1;
    s->h_detect[t] = 1;
    return VISO_OK;
}

int viso_seq_upload_chunk_raw(viso_seq* s, int t0, int count, const uint8_t* images)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!s->det_set) return ctx->fail(VISO_ERR_ARG, "seq_upload_chunk_raw: viso_seq_set_detector has not been called");
    if (t0 < 0 || count < 1 || t0 + count > s->F || !images) return ctx->fail(VISO_ERR_ARG, "seq_upload_chunk_raw: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_guard(s, t0);
    if (rc) return rc;
    cudaStream_t st = ctx->copy_stream;
    const size_t bytes = (size_t)s->img_w * s->img_h;
    const size_t total = 2 * (size_t)count * bytes, piece = (size_t)16 << 20;
    for (size_t off = 0; off < total; off += piece)
        CK(cudaMemcpyAsync(s->imgL + 2 * (size_t)t0 * bytes + off, images + off, std::min(piece, total - off),
                           cudaMemcpyHostToDevice, st));
    for (int i = 0; i < count; ++i) {
        s->h_nL[t0 + i] = s->h_nR[t0 + i] = s->det_n;
        s->h_from_image[t0 + i] = 1;
        s->h_detect[t0 + i] = 1;
    }
    return VISO_OK;
}

int viso_seq_set_seeds(viso_seq* s, const uint32_t* seeds, int ransac_iter)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (ransac_iter < 0 || ransac_iter > s->maxH || (ransac_iter > 0 && !seeds)) return ctx->fail(VISO_ERR_ARG, "seq_set_seeds: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_guard(s, 0, false);
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
    if (t0 > s->run_hi_total)
        return ctx->fail(VISO_ERR_ARG, "seq_run_range: frames before t0 have not been run since they were uploaded (ranges must be submitted in order)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int nf = t1 - t0;
    /* The per-frame counts and flags go up on the COPY stream, behind the frames they describe: a host-to-device
     * copy on the compute stream would queue on the copy engine behind every upload already enqueued by other
     * sequence objects (measured: it serialises double-buffered pipelines completely).  Kernels of an earlier
     * submission may still read these words, hence the guard. */
    {
        int rc = upload_guard(s, t0, false);
        if (rc) return rc;
        cudaStream_t cs = ctx->copy_stream;
        CK(cudaMemcpyAsync(s->nL + t0, s->h_nL + t0, (size_t)nf * 4, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(s->nR + t0, s->h_nR + t0, (size_t)nf * 4, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(s->from_image + t0, s->h_from_image + t0, (size_t)nf * 4, cudaMemcpyHostToDevice, cs));
        if (s->det_set) CK(cudaMemcpyAsync(s->detect + t0, s->h_detect + t0, (size_t)nf * 4, cudaMemcpyHostToDevice, cs));
        CK(cudaEventRecord(s->ev_counts, cs));
        s->counts_hi = std::max(s->counts_hi, t1);
    }
    /* everything uploaded so far (frames, seeds, counts) is visible to the kernels below */
    CK(cudaEventRecord(s->ev_copy, ctx->copy_stream));
    CK(cudaStreamWaitEvent(st, s->ev_copy, 0));
    int nl = 0;
    if (t0 == 0) {
        CK(viso_launch_zero(s->pairs, 4, st));
        CK(viso_launch_zero(s->err, 1, st));
        nl += 2;
    }
    int max_n = 0, max_nL = 0, any_img = 0, any_f32 = 0, any_det = 0;
    for (int t = t0; t < t1; ++t) {
        if (s->h_detect[t]) any_det = 1;
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

    if (any_det) {
        /* overwrites the counts copied above with the detector's own */
        CK(viso_launch_detect(s->detect_jobs + 2 * t0, 2 * nf, s->hc, st));
        nl += 2;
    }
    CK(viso_launch_grid(s->grid_jobs + 2 * t0, 2 * nf, s->grid, st));
    ++nl;
    /* descriptor rows are written in cell-sorted order, so they follow the grid */
    if (any_f32 && max_n > 0) { CK(viso_launch_pack(s->pack_jobs + 2 * t0, 2 * nf, max_n, s->dlen, s->err, st)); ++nl; }
    if (any_img && max_n > 0) {
        CK(viso_launch_extract(s->extract_jobs + 2 * t0, 2 * nf, max_n, s->img_w, s->img_h, s->img_w, 5, st));
        ++nl;
    }
    /* match jobs: frame 0 has one (stereo), frame t >= 1 has three (stereo, temporal L, temporal R) */
    const int mj0 = t0 == 0 ? 0 : 3 * t0 - 2, mj1 = 3 * t1 - 2;
    CK(cudaEventRecord(s->ev0, st));
    CK(viso_launch_match(s->match_jobs + mj0, mj1 - mj0, max_n, max_nt, mp, s->grid, s->pairs,
                         PendingList{s->pending, s->pend_rec, s->pend_job, VISO_SEQ_PENDING_CAP}, ctx->match_mode,
                         ctx->sm_count, st, &nl));
    CK(cudaEventRecord(s->ev1, st));
    CK(viso_launch_sort(s->sort_jobs + t0, nf, max_nL, pd, st));
    ++nl;
    const int c0 = std::max(t0, 1);
    if (t1 > c0) {
        CK(viso_launch_circle(s->circ_jobs + c0, t1 - c0, st));
        ++nl;
        CK(viso_launch_ransac(s->probs + c0, t1 - c0, param->ransac_iter, max_nL, pd, s->strag, ctx->hyp_it_cap, ctx->sm_count, st, &nl));
    }
    ctx->launches += nl;
    CK(cudaEventRecord(s->ev_compute, st));
    s->run_hi = std::max(s->run_hi, t1);
    s->run_hi_total = std::max(s->run_hi_total, t1);
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

/* frames whose keypoints were detected on the device: bring their counts to the host (getters, statistics) */
static int refresh_counts(viso_seq* s)
{
    viso_ctx* ctx = s->ctx;
    if (!s->det_set || !s->ran) return VISO_OK;
    std::vector<int> n(2 * (size_t)s->F);
    CK(cudaMemcpyAsync(n.data(), s->nL, (size_t)s->F * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(n.data() + s->F, s->nR, (size_t)s->F * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < s->F; ++t)
        if (s->h_detect[t] && t < s->run_hi_total) { s->h_nL[t] = n[t]; s->h_nR[t] = n[s->F + t]; }
    return VISO_OK;
}

int viso_seq_stats(viso_seq* s, int64_t* match_bytes, int64_t* sad_pairs, int64_t* sad_evaluated)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    int rcc = refresh_counts(s);
    if (rcc) return rcc;
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

int viso_seq_time_match(viso_seq* s, int which, float* ms, int64_t* sad_pairs, int32_t* n_pending)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (!ms || which < 0 || which > 1) return ctx->fail(VISO_ERR_ARG, "seq_time_match: bad argument");
    if (!s->ran || s->run_hi_total < s->F) return ctx->fail(VISO_ERR_ARG, "seq_time_match: run the whole sequence first");
    if (which == 1 && s->F < 2) return ctx->fail(VISO_ERR_ARG, "seq_time_match: temporal jobs need two frames");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int max_n = 0;
    for (int t = 0; t < s->F; ++t) max_n = std::max(max_n, std::max(s->h_nL[t], s->h_nR[t]));
    viso_match_params ms_, mt_;
    viso_match_params_stereo(&ms_, s->Fm);
    viso_match_params_temporal(&mt_);
    MatchParamsPair mp;
    mp.p[0] = make_match_dev(&ms_);
    mp.p[1] = make_match_dev(&mt_);
    const MatchJob* jobs = which == 0 ? s->match_jobs_stereo : s->match_jobs_temporal;
    const int nj = which == 0 ? s->F : 2 * (s->F - 1);
    int nl = 0;
    CK(viso_launch_zero(s->pairs, 4, st));
    CK(cudaEventRecord(s->ev0, st));
    CK(viso_launch_match(jobs, nj, max_n, max_n, mp, s->grid, s->pairs,
                         PendingList{s->pending, s->pend_rec, s->pend_job, VISO_SEQ_PENDING_CAP}, ctx->match_mode, ctx->sm_count, st, &nl));
    CK(cudaEventRecord(s->ev1, st));
    ctx->launches += nl + 1;
    CK(cudaEventSynchronize(s->ev1));
    CK(cudaEventElapsedTime(ms, s->ev0, s->ev1));
    unsigned long long p[2] = {0, 0};
    CK(cudaMemcpyAsync(p, s->pairs, 16, cudaMemcpyDeviceToHost, st));
    if (n_pending) CK(cudaMemcpyAsync(n_pending, s->pending, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (sad_pairs) *sad_pairs = (int64_t)p[0];
    return VISO_OK;
}

int viso_seq_last_pending(viso_seq* s, int32_t* n_pending)
{
    if (!s || !n_pending) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(n_pending, s->pending, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return VISO_OK;
}

/* ---- parity-test getters ---- */

int viso_seq_get_dense(viso_seq* s, int which, int t, int32_t* out4, int32_t* n)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || which < 0 || which > 2 || !out4 || !n) return ctx->fail(VISO_ERR_ARG, "seq_get_dense: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rcc = refresh_counts(s);
    if (rcc) return rcc;
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
    int rcc = refresh_counts(s);
    if (rcc) return rcc;
    const int cnt = side ? s->h_nR[t] : s->h_nL[t];
    const uint16_t* src = (side ? s->dRu : s->dLu) + (size_t)t * s->cap * VISO_DESC_U16;
    const int* pos = (side ? s->posR : s->posL) + (size_t)t * s->cap;
    *n = cnt;
    if (cnt <= 0) return VISO_OK;
    /* the device keeps the rows in cell-sorted order; hand them back in the caller's keypoint order */
    std::vector<uint16_t> sorted((size_t)cnt * VISO_DESC_U16);
    std::vector<int> po(cnt);
    CK(cudaMemcpyAsync(sorted.data(), src, sorted.size() * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(po.data(), pos, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < cnt; ++i) {
        if (po[i] < 0 || po[i] >= cnt) return ctx->fail(VISO_ERR_CUDA, "seq_get_packed: corrupt position table");
        std::memcpy(rows + (size_t)i * VISO_DESC_U16, sorted.data() + (size_t)po[i] * VISO_DESC_U16, VISO_DESC_U16 * 2);
    }
    return VISO_OK;
}

int viso_seq_get_keypoints(viso_seq* s, int t, int side, float* kp_xy, int32_t* n)
{
    if (!s) return VISO_ERR_ARG;
    viso_ctx* ctx = s->ctx;
    if (t < 0 || t >= s->F || side < 0 || side > 1 || !kp_xy || !n) return ctx->fail(VISO_ERR_ARG, "seq_get_keypoints: bad argument");
    CK(cudaSetDevice(ctx->device));
    int rcc = refresh_counts(s);
    if (rcc) return rcc;
    const int cnt = side ? s->h_nR[t] : s->h_nL[t];
    const float2* src = (side ? s->kpR : s->kpL) + (size_t)t * s->cap;
    *n = cnt;
    if (cnt > 0) CK(cudaMemcpyAsync(kp_xy, src, (size_t)cnt * 8, cudaMemcpyDeviceToHost, ctx->stream));
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
