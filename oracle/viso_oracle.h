/*
 * viso_oracle.h -- CPU ORACLE for the libviso hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a dependency-free restatement (plain C++17, libstdc++ only) of the reference's
 * per-frame hot path, following /root/reference/src/viso.cpp line by line including every
 * quirk.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library
 * (libviso_b200/csrc) never includes, links or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself -- the reference ships no golden
 * vectors and cannot be compiled in this container (no OpenCV / Eigen / Boost headers).  The
 * third-party arithmetic on the path (OpenCV cvflann radiusSearch ordering, cv::mulTransposed,
 * cv::solve(DECOMP_LU), cv::invert, cv::determinant, cv::Sobel) IS pinned: the restatements
 * below are checked bit-for-bit against cv2 4.13 and the resulting vectors are committed under
 * tests/golden/ (generator: tools/make_golden.py).  The glue between them is checked against tests/golden/path_cv2.npz,
 * a literal Python transcription of viso.cpp's control flow that executes the real OpenCV routine at every third-party
 * call site (tools/make_golden_path.py).  No reference BINARY output exists to pin against.
 *
 * All matrices are row-major, like cv::Mat.
 */
#ifndef VISO_ORACLE_H_
#define VISO_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* MatchParams, viso.cpp:48-75 */
typedef struct {
    int32_t enforce_epipolar;
    int32_t enforce_2nd_best;
    int32_t max_neighbors;
    int32_t _pad;
    double radius;
    double sampson_thresh;
    double ratio_2nd_best;
    double F[9];
} vo_match_params;

/* struct param, viso.h:58-72 */
typedef struct {
    double base;
    double f, cu, cv;
    double inlier_threshold;
    double thresh;
    int32_t ransac_iter;
    int32_t _pad;
} vo_param;

/* per-frame-pair record produced by the pipeline (what sequence_odometry needs to chain poses) */
typedef struct {
    double tr[6];
    int32_t ok;        /* 1: ransac_minimize_reproj returned true */
    int32_t n_inliers;
    int32_t n_circ;    /* circular matches; <3 => no RANSAC (viso.cpp:1283-1288) */
    int32_t best_hyp;  /* winning hypothesis index or -1 */
} vo_record;

void vo_match_params_stereo(vo_match_params* p, const double F[9]); /* viso.cpp:62-71 */
void vo_match_params_temporal(vo_match_params* p);                  /* viso.cpp:72-74 */
void vo_param_default(vo_param* p);                                 /* viso.h:60 */

/* viso.cpp:652-666 (+390-407) */
double vo_sampson_distance(const double F[9], float p1x, float p1y, float p2x, float p2y);

/* viso.cpp:668-726 (+170-203 radiusSearch wrapper, cvflann linear index semantics).
 * kp: n x 2 float (x,y); d: n x dlen float row-major.
 * matches: out, capacity n1*3 ints (i1,i2,dist), ordered exactly as the reference's std::sort leaves them.
 * dense_*: optional (may be NULL) per-query results before the ratio test:
 *   dense_idx = best_idx (-1 none), dense_d1/d2 = (int)best_d1 / best_d2 (INT32_MAX for DBL_MAX),
 *   dense_valid = 1 iff the query produced a Match.
 * n_sad: optional, number of SAD evaluations (the survey's P). */
int vo_match_desc(const float* kp1, int n1, const float* kp2, int n2,
                  const float* d1, const float* d2, int dlen,
                  const vo_match_params* sp,
                  int32_t* matches, int32_t* n_matches,
                  int32_t* dense_idx, int32_t* dense_d1, int32_t* dense_d2, int32_t* dense_valid,
                  int64_t* n_sad);

/* one row of the radiusSearch wrapper (viso.cpp:170-203): nb[K] neighbour indices (-1 tail), dists[K]; returns total found */
int vo_radius_search(float qx, float qy, const float* kp2, int n2, float radius, int K, int32_t* nb, float* dists);

/* the reference's std::sort call on Matches, viso.cpp:724 (exposed to test the device introsort) */
void vo_sort_matches(int32_t* matches, int n);

/* viso.cpp:206-243.  matches are n x 3 ints.  circ4: cap nlr*4, pcl: cap nlr*3 (i,k,0). */
int vo_match_circle(const int32_t* mlr, int nlr, const int32_t* mlrp, int nlrp,
                    const int32_t* m11, int n11, const int32_t* m22, int n22,
                    int32_t* circ4, int32_t* pcl3);

/* viso.cpp:501-514: x is 4 x m row-major double */
void vo_collect_matches(const float* kp1, const float* kp2, const int32_t* matches, int m, double* x);

/* viso.cpp:1137-1162 (T=double): x 4 x m, X 3 x m */
void vo_triangulate_rectified_f64(const double* x, int m, double f, double base, double cu, double cv, double* X);
/* mvg.cpp:172-192: x1,x2 2 x m float, X 3 x m float */
void vo_triangulate_rectified_f32(const float* x1, const float* x2, int m, double f, double base,
                                  double c1u, double c1v, float* X);

/* viso.cpp:1401-1497.  X 3 x n, observe 4 x n, active[na].  J (4na x 6), predict (4 x na), residual (4na). */
void vo_compute_J(const double* X, const double* observe, int n, const double tr[6], const vo_param* p,
                  const int32_t* active, int na, double* J, double* predict, double* residual);

/* viso.cpp:1509-1537.  inliers cap n.  min_margin (optional): min_i |err2_i - thr^2| */
int vo_get_inliers(const double* X, const double* observe, int n, const double tr[6], const vo_param* p,
                   int32_t* inliers, double* rms, double* min_margin);

/* cv::mulTransposed(J,JtJ,true) and J.t()*residual as used at viso.cpp:1599-1602; J is rows x 6 */
void vo_mul_transposed(const double* J, int rows, double JtJ[36]);
void vo_Jt_times_r(const double* J, const double* r, int rows, double Jtr[6]);

/* viso.cpp:1583-1623.  returns 1/0; tr updated in place; iters (optional) = iterations executed */
int vo_minimize_reproj(const double* X, const double* observe, int n, double tr[6], const vo_param* p,
                       const int32_t* active, int na, int32_t* iters);

/* viso.cpp:1543-1580 with randomsample (viso.cpp:87-107) replaced by a host-supplied table
 * sample_table[H][3] (H = p->ransac_iter).  best_tr is in/out (keeps caller's value on failure).
 * Optional diagnostics: hyp_tr[H][6], hyp_ok[H], hyp_count[H], best_hyp. */
int vo_ransac_minimize_reproj(const double* X, const double* observe, int n, const vo_param* p,
                              const int32_t* sample_table,
                              double best_tr[6], int32_t* best_inliers, int32_t* n_best,
                              double* hyp_tr, int32_t* hyp_ok, int32_t* hyp_count, int32_t* best_hyp);

/* viso.cpp:87-107 Algorithm S, one std::mt19937(seed) stream for the whole table */
void vo_randomsample_table(uint32_t seed, int H, int N, int32_t* table);
/* pipeline mapping seeds[H][3] (uint32) -> ascending distinct triples in [0,N) (N>=3); integer-exact,
 * shared definition with the device (DESIGN.md "sample seeds") */
void vo_samples_from_seeds(const uint32_t* seeds, int H, int N, int32_t* table);

/* viso.cpp:109-133 */
void vo_tr2mat(const double tr[6], double T[16]);
/* cv::Mat::inv() (DECOMP_LU) on n x n double, used at viso.cpp:1319.  returns 0 if singular */
int vo_invert_lu(const double* A, int n, double* Ainv);
/* cv::solve(A,b,x,DECOMP_LU), viso.cpp:1602 */
int vo_solve_lu(const double* A, const double* b, int n, double* x);
/* cv::determinant (n>3: LU), used by F_from_P (mvg.h:41-66) */
double vo_determinant(const double* A, int n);
/* mvg.h:41-66 (T=double) + the normalisation at viso.cpp:1176-1180 when normalise!=0 */
void vo_F_from_P(const double P1[12], const double P2[12], int normalise, double F[9]);
/* pose = pose * inv(tr2mat(tr)), viso.cpp:1315-1321.  returns 0 if singular */
int vo_pose_update(const double pose[16], const double tr[6], double pose_out[16]);

/* mvg.cpp:124-169: triangulate_dlt, x1,x2 2 x m float; P1,P2 3x4 double; X 3 x m float.
 * SVD via one-sided Jacobi on A^T (the null vector is unique up to sign; the /vt(3,3) divide fixes sign). */
void vo_triangulate_dlt(const float* x1, const float* x2, int m, const double P1[12], const double P2[12], float* X);
/* estimation.cpp:29-51 (float): A,B 3 x n; T 4x4 row-major (maps B->A) */
void vo_solve_rigid_motion(const float* A, const float* B, int n, float T[16]);
/* viso.cpp:326-333: x = h2e(P * e2h(X)), X 3 x n double, P 3x4, x 2 x n; returns -1 on w~0 (overflow_error) */
int vo_project_points(const double* X, int n, const double P[12], double* x);

/* MyFeatureExtractor::computeImpl (viso.cpp:1004-1024) given the Sobel image (h x w float):
 * 11x11 patches (radius r) -> n x (2r+1)^2 float, border rule >0 && <size */
void vo_extract_descriptors(const float* sob, int h, int w, const float* kp, int n, int radius, float* d);

/*
 * Front end.
 *
 * vo_sobel_x: cv::Sobel(image, CV_32F, 1, 0, 3, 1, 0, BORDER_DEFAULT) (viso.cpp:1010) -- integer valued, bit-exact
 * against cv2 (tests/golden/sobel.npz).
 *
 * vo_harris_response: cv::cornerHarris(image, block 3, aperture 5, k, BORDER_DEFAULT) (viso.cpp:930) for 8-bit input.
 * cornerHarris is a float32 pipeline whose rounding is NOT defined by OpenCV (FMA use differs between the vector body
 * and the tail columns, the box filter keeps running sums whose order depends on the stripe split), so this is the
 * CANONICAL float32 evaluation both the oracle and the device follow, no fused multiply-adds, operations in this order:
 *   s = 1/(2^(aperture-1) * block * 255) in double; f0,f1,f2 = (float)(6s), (float)(4s), (float)(1s)
 *   r(x,y) = (p[x+2]-p[x-2]) + 2(p[x+1]-p[x-1])           (exact), pixels reflected (BORDER_REFLECT_101)
 *   Dx = f0*r(y); Dx += f1*(r(y-1)+r(y+1)); Dx += f2*(r(y-2)+r(y+2))
 *   t(x,y) = f0*p[x]; t += f1*(p[x-1]+p[x+1]); t += f2*(p[x-2]+p[x+2])
 *   Dy = 2*(t(y+1)-t(y-1)); Dy += (t(y+2)-t(y-2))
 *   xx = Dx*Dx, xy = Dx*Dy, yy = Dy*Dy; box: rs = (c[x-1]+c[x])+c[x+1]; S = (rs[y-1]+rs[y])+rs[y+1], the covariance
 *   images reflected (BORDER_REFLECT_101 on positions, as cv::boxFilter does)
 *   response = (a*c - b*b) - (k*(a+c))*(a+c)
 * Checked against cv2.cornerHarris to 1e-6 of the image maximum (tests/test_oracle_golden.py); parity with the
 * reference for this function is therefore "within float rounding", not bit-exact -- the device is bit-exact with THIS.
 *
 * vo_detect_harris_binned: HarrisBinnedFeatureDetector::detectImpl (viso.cpp:925-976): per bin (binx outer, biny inner)
 * the n/(nbinx*nbiny) largest |response| != 0.  order_rule 0: literally std::nth_element on the scan-ordered vector
 * (libstdc++ introselect; the order of the kept elements is whatever it leaves); order_rule 1: canonical -- kept set =
 * largest by (|response|, x, y), emitted ascending by that key (the rule the device implements; same SET as rule 0
 * whenever no two responses tie at the cut).  kp_xy: up to n rows of (x, y); returns the keypoint count.
 */
void vo_sobel_x(const uint8_t* img, int h, int w, float* sob);
void vo_harris_response(const uint8_t* img, int h, int w, float k, float* resp);
int vo_detect_harris_binned(const uint8_t* img, int h, int w, int n, int nbinx, int nbiny, float k, int order_rule,
                            float* kp_xy, float* kp_resp /* nullable */);

/*
 * Whole-sequence pipeline, viso.cpp:1205-1327 minus detection/description/debug output.
 * Frame t has nL[t]/nR[t] keypoints starting at row offL[t]/offR[t] of kpL/kpR (x,y float) and dL/dR (dlen floats/row).
 * seeds: [n_frames][H][3] uint32 sample seeds (frame 0's block unused).
 * records: n_frames entries; record[0] is zeroed (first frame yields no pose).
 * Optional dumps for parity tests (may be NULL):
 *   lr_matches: [sum nL][3] rows at offL[t]..; lr_count[t]
 *   m11_dense/m22_dense: per query (offL / offR rows) {idx,d1,d2,valid}
 *   circ: [sum nL][4] at offL[t]; inliers: [sum nL] at offL[t]
 */
int vo_sequence(int n_frames, const int32_t* nL, const int32_t* nR, const int64_t* offL, const int64_t* offR,
                const float* kpL, const float* kpR, const float* dL, const float* dR, int dlen,
                const double P1[12], const double P2[12], const vo_param* param_in /* ransac_iter etc; calib overwritten */,
                const uint32_t* seeds,
                vo_record* records,
                int32_t* lr_matches, int32_t* lr_count,
                int32_t* m11_dense, int32_t* m22_dense,
                int32_t* circ, int32_t* inliers,
                double* poses /* [n_frames+?][16], first = identity */, int32_t* n_poses);

#ifdef __cplusplus
}
#endif
#endif
