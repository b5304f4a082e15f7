/*
 * viso_b200.h -- C-ABI of the B200-native libviso hot path (libviso_b200/libviso_b200.so).
 *
 * The reference (alexkreimer/libviso) has no FFI: its boundary is the set of free C++ functions declared in
 * src/viso.h, src/mvg.h and src/estimation.h and linked statically into `kitti` and `tester`
 * (reference src/CMakeLists.txt:17-20).  Every entry point below replaces the body of one of those functions
 * (file:line given per function); the C++ mirror of the reference headers that calls them lives in
 * libviso_b200/host/ and INTEGRATION.md shows the binding a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes, no C++ / torch types; all matrices row-major like cv::Mat;
 * every function returns VISO_OK (0) or a negative viso_status and never throws; text for the last error of a
 * context is available through viso_last_error().  All host-buffer functions are synchronous.  There is NO CPU
 * fallback: without a CUDA device viso_create() fails and nothing else can be called.
 *
 * Domain restrictions (checked on the device, VISO_ERR_DOMAIN when violated):
 *   - descriptors must be integer valued with |v| <= 1023 and desc_len <= 126 (the reference's descriptors are
 *     3x3 Sobel-x responses of 8-bit images, |v| <= 1020, 121 per keypoint: viso.cpp:999-1002,1010);
 *   - max_neighbors >= 1.
 */
#ifndef VISO_B200_H_
#define VISO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VISO_ABI_VERSION 1
#define VISO_DESC_PAD 128 /* packed descriptor row: 128 x u16 = 256 B */

typedef enum {
    VISO_OK = 0,
    VISO_ERR_CUDA = -1,        /* a CUDA runtime call failed; see viso_last_error */
    VISO_ERR_ARG = -2,         /* bad argument (null pointer, negative size, ...) */
    VISO_ERR_DOMAIN = -3,      /* input outside the supported domain (see above) */
    VISO_ERR_DIV0 = -4,        /* h2e: homogeneous w == 0 (the reference throws std::overflow_error, misc.h:118-119) */
    VISO_ERR_DUPLICATE = -5,   /* match_circle: a Matches list has a repeated query index (never produced by match_desc) */
    VISO_ERR_NOMEM = -6
} viso_status;

/* MatchParams, reference src/viso.cpp:48-75 */
typedef struct {
    int32_t enforce_epipolar;
    int32_t enforce_2nd_best;
    int32_t max_neighbors;
    int32_t _pad;
    double radius;
    double sampson_thresh;
    double ratio_2nd_best;
    double F[9];
} viso_match_params;

/* struct param, reference src/viso.h:58-72 (calib.f/cu/cv flattened) */
typedef struct {
    double base;
    double f, cu, cv;
    double inlier_threshold;
    double thresh;
    int32_t ransac_iter;
    int32_t _pad;
} viso_param;

/* One frame pair's result: what sequence_odometry (viso.cpp:1313-1324) needs to chain poses. 64 bytes. */
typedef struct {
    double tr[6];
    int32_t ok;         /* ransac_minimize_reproj returned true */
    int32_t n_inliers;
    int32_t n_circ;     /* circular matches; < 3 => frame skipped (viso.cpp:1283-1288) */
    int32_t best_hyp;   /* winning hypothesis or -1 */
} viso_record;

typedef struct viso_ctx viso_ctx;
typedef struct viso_seq viso_seq;

int viso_abi_version(void);

/* ---- context ---- */
int viso_create(viso_ctx** ctx, int device);
void viso_destroy(viso_ctx* ctx);
const char* viso_last_error(const viso_ctx* ctx);
void* viso_stream(viso_ctx* ctx);                  /* the cudaStream_t every kernel of this context is launched on */
int viso_sync(viso_ctx* ctx);
/* Uploads of a context's sequence objects run on the context's own copy stream.  Several contexts that take turns on
 * one PCIe link (double buffering) should share ONE copy stream so that their uploads are served first-in first-out
 * instead of interleaved piece by piece: `ctx` uses `owner`'s copy stream from now on (owner == NULL: its own again).
 * `owner` must outlive `ctx`'s use of it. */
int viso_share_copy_stream(viso_ctx* ctx, viso_ctx* owner);
/* extent of the uniform candidate grid (pixels); coordinates outside are clamped into border cells, so any
 * value is correct, a matching one is fast.  Default 1248 x 384. */
int viso_set_image_extent(viso_ctx* ctx, int width, int height);
/* Which kernel path match_desc runs through (results are identical; tests and A/B measurements force one):
 * 0 auto (descriptor rows of a query tile's neighbourhood staged in shared memory when they fit, else gathered
 * through L1), 1 generic per-query kernel only, 2 gather tile kernel, 3 staged tile kernel.  The environment
 * variable VISO_MATCH_MODE = generic | gather | staged sets the initial value at viso_create. */
int viso_set_match_mode(viso_ctx* ctx, int mode);
/* number of kernel launches issued by this context since creation (bench.py's gpu_launches) */
int64_t viso_launch_count(const viso_ctx* ctx);
/* CUDA-event stopwatch on the context stream: begin records an event; end records a second one, waits for it and
 * returns the elapsed milliseconds of everything enqueued in between */
int viso_timer_begin(viso_ctx* ctx);
int viso_timer_end(viso_ctx* ctx, float* ms);

void viso_match_params_stereo(viso_match_params* p, const double F[9]); /* MatchParams(Mat F), viso.cpp:62-71 */
void viso_match_params_temporal(viso_match_params* p);                  /* MatchParams(), viso.cpp:72-74 */
void viso_param_default(viso_param* p);                                 /* param(), viso.h:60 */

/* ---- match_desc, reference src/viso.cpp:668-726 (+ radiusSearch :170-203, sampsonDistance :652-666) ----
 * kp: n x 2 float (x,y); d: n x desc_len float.  Dense per-query outputs (n1 each, host):
 *   best_idx (-1: none), best_d1, best_d2 (INT32_MAX: none), valid (1 iff the reference pushes a Match).
 * The caller compacts valid rows in query order into Match(i,best_idx,best_d1) and applies std::sort by dist
 * (viso.cpp:724), or calls viso_match_desc_sorted / viso_sort_matches, which do both on the device. */
int viso_match_desc(viso_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2,
                    const float* d1, const float* d2, int desc_len, const viso_match_params* params,
                    int32_t* best_idx, int32_t* best_d1, int32_t* best_d2, int32_t* valid);
/* same, plus the compaction and the reference's std::sort order done on the device (restated libstdc++
 * introsort); matches: n1 x 3 ints capacity, n_matches out */
int viso_match_desc_sorted(viso_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2,
                           const float* d1, const float* d2, int desc_len, const viso_match_params* params,
                           int32_t* matches, int32_t* n_matches);

/* ---- std::sort(match.begin(), match.end(), by dist), reference src/viso.cpp:724: the order libstdc++'s introsort
 * gives (the reference's sort is unstable; the order of equal distances is observable downstream).  In place. */
int viso_sort_matches(viso_ctx* ctx, int32_t* matches, int n);

/* ---- match_circle, reference src/viso.cpp:206-243.  Matches are m x 3 ints. circ4: cap nlr x 4, pcl3: cap nlr x 3 */
int viso_match_circle(viso_ctx* ctx, const int32_t* match_lr, int nlr, const int32_t* match_lr_prev, int nlrp,
                      const int32_t* match11, int n11, const int32_t* match22, int n22,
                      int32_t* circ4, int32_t* pcl3, int32_t* n_out);

/* ---- collect_matches (Mat x, 4 x m), reference src/viso.cpp:501-514, fused with
 *      triangulate_rectified<double>, reference src/viso.cpp:1137-1162.  x: 4 x m, X: 3 x m (either may be NULL) */
int viso_collect_triangulate(viso_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2,
                             const int32_t* matches, int m, double f, double base, double cu, double cv,
                             double* x, double* X);
/* triangulate_rectified<double>(x, ...), reference src/viso.cpp:1137-1154 */
int viso_triangulate_rectified_f64(viso_ctx* ctx, const double* x, int m, double f, double base, double cu, double cv,
                                   double* X);
/* triangulate_rectified (float), reference src/mvg.cpp:172-192; x1,x2: 2 x m, X: 3 x m */
int viso_triangulate_rectified_f32(viso_ctx* ctx, const float* x1, const float* x2, int m, double f, double base,
                                   double c1u, double c1v, float* X);
/* triangulate_dlt, reference src/mvg.cpp:124-169; x1, x2: 2 x m float, P1, P2: 3 x 4 double, X: 3 x m float.  The
 * null vector of a 4 x 4 SVD is unique only up to rounding: agreement with cv::SVD is ~1e-6 relative, not bitwise. */
int viso_triangulate_dlt(viso_ctx* ctx, const float* x1, const float* x2, int m, const double P1[12], const double P2[12],
                         float* X);
/* solveRigidMotion, reference src/estimation.cpp:29-51 (Kabsch); A, B: 3 x n float; T: 4 x 4 float, T * B ~ A */
int viso_solve_rigid_motion(viso_ctx* ctx, const float* A, const float* B, int n, float T[16]);
/* projectPoints(X,P), reference src/viso.cpp:326-333 (e2h/h2e misc.h:90-124); X 3 x n, P 3x4, x 2 x n */
int viso_project_points(viso_ctx* ctx, const double* X, int n, const double P[12], double* x);

/* ---- get_inliers, reference src/viso.cpp:1509-1537.  X 3 x n, observe 4 x n ---- */
int viso_get_inliers(viso_ctx* ctx, const double* X, const double* observe, int n, const double tr[6],
                     const viso_param* param, int32_t* inliers, int32_t* n_inliers);
/* ---- minimize_reproj, reference src/viso.cpp:1583-1623 (compute_J :1401-1497).  tr in/out, ok out ---- */
int viso_minimize_reproj(viso_ctx* ctx, const double* X, const double* observe, int n, double tr[6],
                         const viso_param* param, const int32_t* active, int n_active, int32_t* ok);
/* ---- ransac_minimize_reproj, reference src/viso.cpp:1543-1580, with randomsample (:87-107) replaced by the
 * host-supplied sample_table[param->ransac_iter][3].  tr in/out.  Optional diagnostics may be NULL:
 * hyp_tr[H][6], hyp_ok[H], hyp_count[H], best_hyp. */
int viso_ransac_minimize_reproj(viso_ctx* ctx, const double* X, const double* observe, int n,
                                const viso_param* param, const int32_t* sample_table,
                                double tr[6], int32_t* inliers, int32_t* n_inliers, int32_t* ok,
                                double* hyp_tr, int32_t* hyp_ok, int32_t* hyp_count, int32_t* best_hyp);
/* default sample table: Knuth Algorithm S as in randomsample (viso.cpp:87-107), one std::mt19937(seed) stream */
void viso_randomsample_table(uint32_t seed, int H, int N, int32_t* table);
/* the pipeline's seeds -> ascending distinct triple mapping (see DESIGN.md "sample seeds"), N >= 3 */
void viso_samples_from_seeds(const uint32_t* seeds, int H, int N, int32_t* table);

/* ---- host bookkeeping kept outside the device: tr2mat (viso.cpp:109-133), F_from_P<double> (mvg.h:41-66 plus
 * the normalisation at viso.cpp:1176-1180), pose = pose * inv(tr2mat(tr)) (viso.cpp:1315-1321) ---- */
void viso_tr2mat(const double tr[6], double T[16]);
void viso_F_from_P(const double P1[12], const double P2[12], int normalise, double F[9]);
int viso_pose_update(const double pose[16], const double tr[6], double pose_out[16]);

/* ==== batched sequence pipeline: the per-frame loop of sequence_odometry, reference src/viso.cpp:1205-1327,
 * for all frames of a sequence at once (frame pairs are independent given features: the pose never feeds back,
 * viso.cpp:1208-1222 vs :1317-1321).  Stereo match + sort + triangulate per frame; two temporal matches, circle
 * closure, gather, RANSAC + Gauss-Newton per frame pair.  ==== */
int viso_seq_create(viso_ctx* ctx, int n_frames, int max_kp, int desc_len, int max_ransac_iter, viso_seq** seq);
void viso_seq_destroy(viso_seq* seq);
/* F, base, f, cu, cv from P1,P2 exactly as viso.cpp:1176-1187 */
int viso_seq_set_calib(viso_seq* seq, const double P1[12], const double P2[12]);
/* host -> device copy of one frame's features.  All viso_seq_upload_* functions enqueue cudaMemcpyAsync straight from
 * the caller's buffers on the context's COPY stream and return at once: with pinned buffers the copy is truly
 * asynchronous, so the source (this includes the seeds of viso_seq_set_seeds) must stay valid and unmodified until
 * viso_sync(), viso_seq_download() or a later synchronisation point; pageable buffers are staged by the driver before
 * the call returns.  The f32 descriptor staging on the device (2 x n_frames x max_kp x desc_len floats) is allocated
 * by the first call of this function; image-mode sequences never pay for it. */
int viso_seq_upload_frame(viso_seq* seq, int t, const float* kpL, int nL, const float* kpR, int nR,
                          const float* dL, const float* dR);
/* Front-end on the device: MyFeatureExtractor::computeImpl, reference src/viso.cpp:1004-1024 (cv::Sobel x 3x3,
 * BORDER_REFLECT_101, 11x11 patch around Point2i(kp.pt), samples with row/col <= 0 or >= size are 0).  Instead of
 * the n x 121 f32 descriptor matrix (484 B per keypoint) the caller uploads the two 8-bit images of the frame
 * (width*height bytes each, rows contiguous) and the keypoints; the descriptors are computed on the device straight
 * into the packed layout.  Results are identical to uploading the reference's descriptors.  desc_len must be 121. */
int viso_seq_set_image_size(viso_seq* seq, int width, int height);
int viso_seq_upload_frame_images(viso_seq* seq, int t, const uint8_t* imgL, const uint8_t* imgR,
                                 const float* kpL, int nL, const float* kpR, int nR);
/* The same for `count` consecutive frames in three copies (large copies reach PCIe peak, per-frame ones do not):
 * images = count x [left, right] x height x width bytes; kpL / kpR = count x viso_seq_capacity() x 2 floats (row i
 * holds nL[i] / nR[i] keypoints, the rest is ignored). */
int viso_seq_capacity(const viso_seq* seq);
int viso_seq_upload_chunk_images(viso_seq* seq, int t0, int count, const uint8_t* images, const float* kpL,
                                 const int32_t* nL, const float* kpR, const int32_t* nR);
/* Device address of the F 64-byte result records (viso_record layout; record t is valid once the kernels of a run
 * that covered frame t have finished on viso_stream(ctx)).  For consumers that stay on the GPU -- e.g. an NCCL gather
 * of the records of several ranks -- without a round trip through host memory. */
void* viso_seq_records_device(viso_seq* seq);
/* device memory held by the sequence object, bytes */
int64_t viso_seq_device_bytes(const viso_seq* seq);
/* Detector on the device: HarrisBinnedFeatureDetector::detectImpl, reference src/viso.cpp:911-979
 * (cv::cornerHarris(block 3, aperture 5, k, BORDER_DEFAULT); nbinx x nbiny bins of (width/nbinx) x (height/nbiny)
 * pixels; per bin the n_features/(nbinx*nbiny) largest |response| != 0; bins concatenated binx outer, biny inner).
 * cornerHarris is float32 with a rounding OpenCV does not define (FMA use differs between its vector body and tail
 * columns, its box filter keeps running sums that depend on the stripe split), so the response is evaluated in ONE
 * canonical order, float32, no fused multiply-adds, every border BORDER_REFLECT_101:
 *   s = 1/(16*3*255) in double; f0,f1,f2 = (float)(6s), (float)(4s), (float)(1s)
 *   r(x,y) = (p[x+2]-p[x-2]) + 2(p[x+1]-p[x-1]);   Dx = f0*r(y); Dx += f1*(r(y-1)+r(y+1)); Dx += f2*(r(y-2)+r(y+2))
 *   t(x,y) = f0*p[x]; t += f1*(p[x-1]+p[x+1]); t += f2*(p[x-2]+p[x+2]);   Dy = 2*(t(y+1)-t(y-1)); Dy += t(y+2)-t(y-2)
 *   a,b,c = 3x3 sums of Dx*Dx, Dx*Dy, Dy*Dy: rows (c[x-1]+c[x])+c[x+1], then (rs[y-1]+rs[y])+rs[y+1]
 *   response = (a*c - b*b) - (k*(a+c))*(a+c)
 * (within 2e-6 of the image maximum of cv2.cornerHarris; on the synthetic frames the kept keypoints are the same).
 * Inside a bin the kept keypoints are the largest by (|response|, x, y), emitted ascending -- std::nth_element leaves
 * an implementation-defined order there.  The reference never initialises its k (viso.cpp:915-919); 0.04f is its
 * constructor's default argument.
 *   viso_detect_harris: one image from host memory (rows `pitch` bytes apart) -> kp_xy (up to n_features rows of x, y),
 *   kp_response (nullable), *n_out.
 *   viso_seq_set_detector + viso_seq_upload_frame_raw / viso_seq_upload_chunk_raw: the sequence object detects and
 *   describes on the device, only the 8-bit images are uploaded (viso_seq_set_image_size first; n_features rounded
 *   down to a multiple of the bin count must not exceed the sequence's max_kp).
 *   viso_seq_get_keypoints: the keypoints of frame t, side 0 (left) / 1 (right), after a run. */
int viso_detect_harris(viso_ctx* ctx, const uint8_t* img, int width, int height, int pitch, int n_features, int nbinx,
                       int nbiny, float k, float* kp_xy, float* kp_response, int32_t* n_out);
/* Tuning / test knob of the RANSAC hypothesis stage (viso.cpp:1555-1562): a three-point Gauss-Newton solve that has not
 * converged after `cap` iterations moves from the four-lanes-per-hypothesis kernel to a warp-per-hypothesis kernel for
 * iterations cap .. 99.  Results do not depend on it (same operations, differently spread over lanes); 100 keeps
 * everything in the first kernel.  Default 8; VISO_HYP_IT_CAP in the environment at viso_create overrides it. */
int viso_set_hyp_iteration_cap(viso_ctx* ctx, int cap);

/* Test hook: sin(x[i]), cos(x[i]) as the estimation kernels evaluate them (glibc's algorithm, libviso_b200/csrc/
 * glibc_sincos.h): bit-identical to the host libm for |x| < 0x1.921fbp+26 = 105414336. */
int viso_debug_sincos(viso_ctx* ctx, const double* x, int n, double* s, double* c);

/* MyFeatureExtractor::computeImpl for one image (reference src/viso.cpp:1004-1024, descriptor radius 5): desc = n x 121
 * float, the reference's cv::Mat layout.  (Inside a sequence object the descriptors stay on the device in the packed
 * layout; this entry point exists for callers that use the extractor on its own.) */
int viso_extract_descriptors(viso_ctx* ctx, const uint8_t* img, int width, int height, int pitch, const float* kp_xy, int n,
                             float* desc);
int viso_seq_set_detector(viso_seq* seq, int n_features, int nbinx, int nbiny, float k);
int viso_seq_upload_frame_raw(viso_seq* seq, int t, const uint8_t* imgL, const uint8_t* imgR);
int viso_seq_upload_chunk_raw(viso_seq* seq, int t0, int count, const uint8_t* images);
int viso_seq_get_keypoints(viso_seq* seq, int t, int side, float* kp_xy, int32_t* n);
/* Uploads run on a dedicated copy stream.  viso_seq_run_range() orders itself after every upload enqueued so far and
 * enqueues the pipeline for frames [t0, t1) (frame pairs (t-1, t) for t in [max(t0,1), t1); frame t0-1 must have been
 * run before), so uploading chunk k+1 from pinned memory overlaps the kernels of chunk k. */
/* Frames are processed in order: t0 may not exceed the number of frames already run since they were last uploaded
 * (frame pair t0 reads frame t0 - 1's matches and 3-D points from the previous submission); VISO_ERR_ARG otherwise. */
int viso_seq_run_range(viso_seq* seq, const viso_param* param, int t0, int t1);
/* enqueue the whole pipeline for frames [0, n_frames).  seeds: host [n_frames][ransac_iter][3] uint32 (copied). */
int viso_seq_run(viso_seq* seq, const viso_param* param, const uint32_t* seeds);
/* same but the seeds are already on the device from a previous viso_seq_run / viso_seq_set_seeds */
int viso_seq_set_seeds(viso_seq* seq, const uint32_t* seeds, int ransac_iter);
int viso_seq_run_resident(viso_seq* seq, const viso_param* param);
/* device -> host: records[n_frames] (record 0 zero: the first frame has no pose); synchronises */
int viso_seq_download(viso_seq* seq, viso_record* records);
/* rank-0 bookkeeping of viso.cpp:1189-1190,1313-1321: poses[0] = I, one 4x4 appended per ok record. returns count */
int viso_chain_poses(const viso_record* records, int n_frames, double* poses /* (n_frames) x 16 */);
/* algorithmic HBM bytes of the sad_match launch of one viso_seq_run (SURVEY 8d formula, int16 layout); the number of
 * candidate pairs that reach the SAD in the reference (the survey's P); and the number of exact SADs the kernel
 * actually evaluated after lower-bound pruning.  Any pointer may be NULL. */
int viso_seq_stats(viso_seq* seq, int64_t* match_bytes, int64_t* sad_pairs, int64_t* sad_evaluated);
/* elapsed ms of the sad_match kernel in the last run (CUDA events on the context stream) */
int viso_seq_match_ms(viso_seq* seq, float* ms);
/* Measurement hook: launch ONLY the stereo (which = 0) or ONLY the temporal (which = 1) match_desc jobs of an already run
 * sequence again and return the device time of that launch (CUDA events), the (query, candidate) pairs that reached the
 * SAD and the queries the tile kernel left to the generic kernel.  Results are rewritten with identical values. */
int viso_seq_time_match(viso_seq* seq, int which, float* ms, int64_t* sad_pairs, int32_t* n_pending);
/* queries of the last submission that the tile kernel left to the generic kernel (top-K cut needed, candidate list or
 * staging buffer too small): a performance diagnostic, results do not depend on it */
int viso_seq_last_pending(viso_seq* seq, int32_t* n_pending);
/* parity-test getters (host copies).  which: 0 = stereo (frame t), 1 = temporal left (t vs t-1), 2 = temporal right */
int viso_seq_get_dense(viso_seq* seq, int which, int t, int32_t* out4 /* n x 4: idx,d1,d2,valid */, int32_t* n);
/* packed descriptor rows (n x 128 u16, value + 1024, pad 0) of frame t; side 0 = left, 1 = right */
int viso_seq_get_packed(viso_seq* seq, int t, int side, uint16_t* rows, int32_t* n);
int viso_seq_get_lr_matches(viso_seq* seq, int t, int32_t* matches3, int32_t* n);
int viso_seq_get_circ(viso_seq* seq, int t, int32_t* circ4, int32_t* pcl2, int32_t* n);
int viso_seq_get_inliers(viso_seq* seq, int t, int32_t* inliers, int32_t* n);
int viso_seq_get_hyp(viso_seq* seq, int t, double* hyp_tr, int32_t* hyp_ok, int32_t* hyp_count);

#ifdef __cplusplus
}
#endif
#endif
