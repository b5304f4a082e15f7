/*
 * viso_dev.h -- device-side data layout shared by the kernel translation units (match.cu, sort_circle.cu, estimation.cu, geometry.cu) and capi.cu (not part of the public ABI).
 *
 * HBM layout (see DESIGN.md "Data layout"):
 *   keypoints        float2[n]                      original order (index = the reference's keypoint index)
 *   descriptors f32  float[n][desc_len]             the reference's cv::Mat layout (input only)
 *   descriptors u16  uint16[n][128]                 biased (v+1024) Sobel values, pad elements 0; 256 B per row;
 *                                                   rows in CELL-SORTED order (row p belongs to srec[p]), so the
 *                                                   neighbourhood of a query tile is one contiguous span per grid row
 *   candidate grid   cell_start int[ncell+1], srec uint4[n] = (x, y, original index, row sum), pos_of int[n]
 *                                                   counting sort by 16-px cell; row sum = sum of the biased row
 *                                                   (SAD = sum a + sum b - 2 sum min(a,b)); pos_of[i] = sorted position
 *   dense match out  int4[n]                        (best_idx, best_d1, best_d2, valid) per query
 */
#ifndef VISO_DEV_H_
#define VISO_DEV_H_

#include <cuda_runtime.h>
#include <stdint.h>

#define VISO_DESC_U16 128
#define VISO_GRID_CS 16          /* candidate grid cell size in pixels (power of two: exact float scaling) */
#define VISO_HIST_BINS 128
#define VISO_TIE_CAP 64
#define VISO_LIST_CAP 128        /* per-warp scanned-candidate list (evaluated and reset when it fills) */
#ifndef VISO_MATCH_WARPS
#define VISO_MATCH_WARPS 4
#endif
#ifndef VISO_EVAL_DEPTH
#define VISO_EVAL_DEPTH 1          /* SAD steps whose row loads are issued together (4 rows each); with 36 warps per SM one
                                      step in flight per warp measures best (6.22 ms; 2: 6.27, 3 / 4: 6.27) */
#endif
#ifndef VISO_MATCH_MINB
#define VISO_MATCH_MINB 9          /* resident CTAs per SM the tile kernel is compiled for: 56 registers, 36 warps per SM.  Since
                                      the running sums moved to the FMA pipe the kernel fits: 6.48 -> 6.27 ms (7: 6.64, 8: 6.48,
                                      10 = 48 registers: 6.80) */
#endif
#ifndef VISO_TILE_W
#define VISO_TILE_W 6            /* query tile of sad_match: 6 x 4 cells = 96 x 64 px */
#endif
#ifndef VISO_TILE_H
#define VISO_TILE_H 4
#endif
#define VISO_STRIP_QPC 16         /* queries per CTA of the generic match kernel */
#define VISO_PENDING (-2)         /* dense result .w: left by the tile kernel for the generic kernel */
#define VISO_MAX_REG_ROWS 64     /* grid rows a staged tile neighbourhood may span */

/* `one` is always 1: a multiplier the compiler cannot see through, so that `a * one + b` stays an IMAD (FMA pipe) where the
 * ALU pipe is the busy one (match.cu, eval_batch) */
struct GridCfg { int gx, gy; unsigned one = 1; };

struct SetView {
    const float2* xy;        /* original order */
    const int* n;            /* device pointer to the keypoint count */
    const uint16_t* desc;    /* packed rows, CELL-SORTED order (row p <-> srec[p]) */
    const uint4* srec;       /* cell-sorted candidate records: (x, y, original index, row sum) */
    const int* cell_start;   /* ncell+1 */
    const int* pos_of;       /* original index -> sorted position */
};

struct MatchParamsDev {
    float radius;
    int K;
    int epipolar;
    int second_best;
    double sampson_thresh;
    double ratio;
    double F[9];
};

struct MatchParamsPair { MatchParamsDev p[2]; };

struct MatchJob {
    SetView q, t;
    int4* out;               /* dense per query, indexed by original query index */
    int mode;                /* index into MatchParamsPair */
    int pad;
};

struct PackJob {
    const float* d;          /* n x dlen float, original order */
    const int* n;
    uint16_t* out;           /* n x 128 u16, cell-sorted order */
    uint4* srec;             /* cell-sorted records: .z = source row, .w receives the row sum */
    const int* from_image;   /* != 0: this set's descriptors come from extract_desc_kernel, skip */
};

/* MyFeatureExtractor::computeImpl on the device (viso.cpp:1004-1024): 8-bit image -> packed descriptor rows */
struct ExtractJob {
    const unsigned char* img; /* rows x pitch */
    const float2* kp;         /* srec == null: keypoints, row k of `out` belongs to kp[k] */
    const int* n;
    uint16_t* out;            /* n x 128 u16 */
    unsigned* rsum;           /* srec == null: row sums */
    uint4* srec;              /* != null: cell-sorted records; row p is extracted at srec[p].xy, its sum goes to srec[p].w */
    const int* from_image;    /* == 0: skip (descriptors were uploaded as f32 rows) */
};

/* HarrisBinnedFeatureDetector::detectImpl on the device (viso.cpp:925-976): 8-bit image -> keypoints */
struct DetectJob {
    const unsigned char* img; /* rows x pitch */
    float2* kp;               /* out: keypoints, bins concatenated in the reference's order */
    int* n;                   /* out: keypoint count */
    float2* tmp;              /* [nbins][per] per-bin slots */
    float* resp_tmp;          /* [nbins][per] |response| of the slots, or null */
    float* resp;              /* out: |response| per keypoint (KeyPoint::response, viso.cpp:968), or null */
    int* bin_count;           /* [nbins] */
    const int* detect;        /* == 0: skip (keypoints were uploaded) */
};

struct HarrisCfg {
    int w, h, pitch;
    int nbinx, nbiny, sx, sy; /* bins and their size in pixels (w / nbinx, h / nbiny) */
    int per;                  /* keypoints kept per bin */
    float k, f0, f1, f2;      /* Harris k and the scaled 5-tap smoothing kernel (include/viso_b200.h) */
};

struct GridJob {
    const float2* xy;
    const int* n;
    int* pos_of;             /* out: original index -> sorted position */
    uint4* srec;             /* out: (x, y, original index, 0); the pack / extract kernels fill in .w */
    int* cell_start;
};

struct ParamDev {
    double base, f, cu, cv, thr2, thresh;
    int H;
    int pad;
};

/* one RANSAC problem (a frame pair, or a standalone call): 3 x n points, 4 x n observations, row stride `stride` */
struct RansacProb {
    const double* X;         /* rows at X + r*stride */
    const double* obs;
    const int* n;            /* device pointer to the correspondence count */
    int stride;
    int H;
    const uint32_t* seeds;   /* [H][3] or null */
    const int* table;        /* [H][3] explicit sample table or null (then derived from seeds) */
    double* hyp_tr;          /* [H][6] */
    int* hyp_ok;             /* [H] */
    int* hyp_count;          /* [H] */
    double* scratch;         /* [stride*4][7] doubles: Jacobian rows + residual for the refine */
    int* inliers;            /* [stride] */
    int* active;             /* [stride] scratch: RANSAC support set */
    struct viso_record_dev* rec;
    double tr_init[6];       /* caller's best_tr (kept when no hypothesis succeeds) */
    int min_n;               /* problems with n < min_n are skipped (3 in the pipeline, viso.cpp:1283) */
    int pad;
};

struct viso_record_dev {
    double tr[6];
    int ok, n_inliers, n_circ, best_hyp;
};

struct SortJob {               /* per frame: compaction + reference sort order + collect + triangulate */
    const int4* dense;         /* stereo dense result */
    const int* n;              /* queries (left keypoints) */
    const float2* kp1;
    const float2* kp2;
    int* matches;              /* [cap][3] */
    int* count;                /* out */
    int* pos_of_query;         /* [cap] position in sorted order or -1 */
    double* x;                 /* 4 x stride */
    double* X;                 /* 3 x stride */
    int stride;
    int pad;
};

struct CircleJob {             /* per frame pair */
    const int* lr; const int* lr_count;          /* match_lr (cur) */
    const int* lrp; const int* lrp_count;        /* match_lr_prev */
    const int* pos_prev;                         /* pos_of_query of previous frame */
    const int* n_prev_left;                      /* bound for pos_prev lookups */
    const int4* m11; const int4* m22;            /* temporal dense results (cur queries) */
    const double* x; const double* Xp;           /* x of cur (4 x stride), X of prev (3 x stride) */
    int* circ4; int* pcl2; int* n_circ;
    double* x_c; double* Xp_c;                   /* 4 x stride, 3 x stride */
    viso_record_dev* rec;
    int stride;
    int pad;
};

/* launch wrappers (defined next to their kernels).  All enqueue on `s` and return cudaGetLastError(). */
cudaError_t viso_launch_pack(const PackJob* jobs, int n_jobs, int max_n, int dlen, int* err_flag, cudaStream_t s);
cudaError_t viso_launch_extract(const ExtractJob* jobs, int n_jobs, int max_n, int width, int height, int pitch, int radius,
                                cudaStream_t s);
size_t viso_harris_smem(const HarrisCfg& c);
size_t viso_harris_cells(const HarrisCfg& c);   /* response slots a bin occupies in shared memory */
cudaError_t viso_launch_detect(const DetectJob* jobs, int n_jobs, const HarrisCfg& c, cudaStream_t s);
cudaError_t viso_launch_zero(void* p, int n_words, cudaStream_t s);   /* a few words, by a kernel */
cudaError_t viso_launch_grid(const GridJob* jobs, int n_jobs, GridCfg g, cudaStream_t s);
/* queries the tile kernel leaves to the generic kernel: a counter and (optionally) the list of the first `cap` of them,
 * (job index, cell-sorted query record); without a list, or when it overflows, the generic kernel scans every query */
struct PendingList {
    int* count;
    uint4* rec;
    int* job;
    int cap;
};
/* mode: VISO_MATCH_AUTO picks the staged tile kernel when a tile's neighbourhood fits shared memory, else the
 * gather tile kernel; the others force one path (tests, A/B measurements).  sm_count sizes the generic kernel's grid. */
#define VISO_MATCH_AUTO 0
#define VISO_MATCH_GENERIC 1
#define VISO_MATCH_GATHER 2
#define VISO_MATCH_STAGED 3
cudaError_t viso_launch_match(const MatchJob* jobs, int n_jobs, int max_nq, int max_nt, const MatchParamsPair& mp,
                              GridCfg g, unsigned long long* sad_pairs, PendingList pend, int mode, int sm_count,
                              cudaStream_t s, int* launches);
cudaError_t viso_launch_sort(const SortJob* jobs, int n_jobs, int max_n, ParamDev p, cudaStream_t s);
cudaError_t viso_launch_circle(const CircleJob* jobs, int n_jobs, cudaStream_t s);
/* strag: device scratch of 2 + 2 * n_probs * max_H ints (count, pad, then (problem, hypothesis) pairs) for the hypotheses
 * the quad kernel hands over to the warp-per-hypothesis kernel after hyp_it_cap iterations; null or a cap of 100 =
 * everything in the quad kernel */
#define VISO_HYP_IT_CAP 8 /* iterations in the quad kernel before a hypothesis is handed to the warp-per-hypothesis one:
                             93 % of the samples have converged by then; 6, 8 and 12 measure the same */
cudaError_t viso_launch_ransac(const RansacProb* probs, int n_probs, int max_H, int max_n, ParamDev p, int* strag,
                               int hyp_it_cap, int sm_count, cudaStream_t s, int* launches);
cudaError_t viso_launch_gn(const double* X, const double* obs, int stride, const int* active, int na, double* tr,
                           int* ok, double* scratch, ParamDev p, cudaStream_t s);
cudaError_t viso_launch_inliers(const double* X, const double* obs, int n, int stride, const double* tr, int* inliers,
                                int* count, ParamDev p, cudaStream_t s);
cudaError_t viso_launch_triangulate_f64(const double* x, int m, int stride, double* X, ParamDev p, cudaStream_t s);
cudaError_t viso_launch_triangulate_f32(const float* x1, const float* x2, int m, double f, double base, double c1u,
                                        double c1v, float* X, cudaStream_t s);
cudaError_t viso_launch_project(const double* X, int n, const double* P, double* x, int* err_flag, cudaStream_t s);
cudaError_t viso_launch_circle_tables(const int* m, int n, int* table, int table_n, int key_col, int val_mode,
                                      int* err_flag, cudaStream_t s);
cudaError_t viso_launch_circle_generic(const int* lr, int nlr, const int* lrp, int nlrp, const int* t11, int n_t11,
                                       const int* tlrp, int n_tlrp, const int* t22, int n_t22,
                                       int* circ4, int* pcl3, int* n_out, cudaStream_t s);
cudaError_t viso_launch_collect_tri(const float2* kp1, int n1, const float2* kp2, int n2, const int* matches, int m,
                                    double* x, double* X, ParamDev p, int* err_flag, cudaStream_t s);

cudaError_t viso_launch_sincos_probe(const double* x, int n, double* s, double* c, cudaStream_t st);
cudaError_t viso_launch_triangulate_dlt(const float* x1, const float* x2, int m, const double* P1, const double* P2, float* X,
                                        cudaStream_t s);
cudaError_t viso_launch_rigid_motion(const float* A, const float* B, int n, float* T, cudaStream_t s);

#endif
