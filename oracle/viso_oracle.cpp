/*
 * viso_oracle.cpp -- CPU ORACLE (test infrastructure only; see viso_oracle.h).
 *
 * Restates /root/reference/src/{viso.cpp,mvg.cpp,mvg.h,estimation.cpp} for the hot path.
 * Each function cites the reference lines it follows.  Build: oracle/Makefile
 * (g++ -O2 -ffp-contract=off: x86-64 g++ without -march flags never contracts to FMA, which is
 * what a stock build of the reference does; the flag makes that explicit).
 *
 * Descriptor domain note: the reference evaluates cv::norm(d2.row(j)-d1.row(i), NORM_L1) in
 * float/double (viso.cpp:702).  Descriptors on this path are integer-valued Sobel responses in
 * [-1020,1020] (viso.cpp:1010), for which every order of accumulation is exact; this oracle
 * accumulates |float difference| in double in index order.
 */
#include "viso_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <random>
#include <utility>
#include <vector>

namespace {

struct Match3 { int32_t v[3]; };

/* ------------------------------------------------------------------------------------------
 * OpenCV hal LUImpl<double> (modules/core/src/matrix_decomp.cpp, OpenCV 4.x), eps = DBL_EPSILON*100.
 * Third-party code absent from /root/reference; restated from its published algorithm and
 * pinned bit-for-bit against cv2 4.13 (tests/golden/lu_*.npz).
 * A is m x m (row stride astep), b is m x n (row stride bstep) or NULL.  Returns 0 if singular,
 * else the permutation sign.
 * ------------------------------------------------------------------------------------------ */
int lu_impl(double* A, int astep, int m, double* b, int bstep, int n)
{
    const double eps = DBL_EPSILON * 100;
    int p = 1;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::fabs(A[j * astep + i]) > std::fabs(A[k * astep + i]))
                k = j;
        if (std::fabs(A[k * astep + i]) < eps)
            return 0;
        if (k != i) {
            for (int j = i; j < m; j++)
                std::swap(A[i * astep + j], A[k * astep + j]);
            if (b)
                for (int j = 0; j < n; j++)
                    std::swap(b[i * bstep + j], b[k * bstep + j]);
            p = -p;
        }
        double d = -1 / A[i * astep + i];
        for (int j = i + 1; j < m; j++) {
            double alpha = A[j * astep + i] * d;
            for (int kk = i + 1; kk < m; kk++)
                A[j * astep + kk] += alpha * A[i * astep + kk];
            if (b)
                for (int kk = 0; kk < n; kk++)
                    b[j * bstep + kk] += alpha * b[i * bstep + kk];
        }
    }
    if (b) {
        for (int i = m - 1; i >= 0; i--)
            for (int j = 0; j < n; j++) {
                double s = b[i * bstep + j];
                for (int k = i + 1; k < m; k++)
                    s -= A[i * astep + k] * b[k * bstep + j];
                b[i * bstep + j] = s / A[i * astep + i];
            }
    }
    return p;
}

/* algebricDistance, viso.cpp:390-407: float coordinates times double F, summed left to right */
double algebraic_distance(const double* F, float p1x, float p1y, float p2x, float p2y)
{
    float a0 = p1x, a1 = p1y, a2 = 1, b0 = p2x, b1 = p2y, b2 = 1;
    return b0 * F[0] * a0 + b0 * F[1] * a1 + b0 * F[2] * a2 +
           b1 * F[3] * a0 + b1 * F[4] * a1 + b1 * F[5] * a2 +
           b2 * F[6] * a0 + b2 * F[7] * a1 + b2 * F[8] * a2;
}


/* radiusSearch wrapper, viso.cpp:170-203, over cvflann::Index<L1<float>> + LinearIndexParams (viso.cpp:684).
 * cvflann (OpenCV's bundled FLANN 1.6.10; absent from /root/reference): LinearIndex::findNeighbors visits
 * every dataset row in index order; L1<float> accumulates |a-b| in float (x, then y);
 * RadiusUniqueResultSet::addPoint keeps dist <= radius in a std::set ordered by (dist, index);
 * sortAndCopy emits the first K; NNIndex::radiusSearch returns the TOTAL number found; the wrapper then
 * overwrites slots [found, K) with -1 (the Mat was pre-filled with -1, viso.cpp:680-681).
 * Pinned against cv2.flann_Index(algorithm=0, distType=L1).radiusSearch (tests/golden/flann_*.npz). */
int radius_search_row(float qx, float qy, const float* kp2, int n2, float radius, int K,
                      std::vector<std::pair<float, int>>& found, int32_t* nb)
{
    found.clear();
    for (int j = 0; j < n2; ++j) {
        float dist = 0;
        dist += std::fabs(kp2[2 * j] - qx);
        dist += std::fabs(kp2[2 * j + 1] - qy);
        if (dist <= radius) found.push_back(std::make_pair(dist, j));
    }
    std::sort(found.begin(), found.end()); /* == iteration order of std::set<DistIndex> */
    for (int j = 0; j < K; ++j) nb[j] = j < (int)found.size() ? found[j].second : -1;
    return (int)found.size();
}

} // namespace

extern "C" {

void vo_match_params_stereo(vo_match_params* p, const double F[9])
{
    /* MatchParams(Mat F), viso.cpp:62-71 */
    std::memset(p, 0, sizeof(*p));
    p->enforce_epipolar = 1;
    p->sampson_thresh = 1;
    p->enforce_2nd_best = 0;
    p->ratio_2nd_best = .8;
    p->max_neighbors = 200;
    p->radius = 80;
    for (int i = 0; i < 9; i++) p->F[i] = F[i];
}

void vo_match_params_temporal(vo_match_params* p)
{
    /* MatchParams(), viso.cpp:72-74 */
    std::memset(p, 0, sizeof(*p));
    p->enforce_epipolar = 0;
    p->enforce_2nd_best = 1;
    p->ratio_2nd_best = .9;
    p->max_neighbors = 250;
    p->radius = 80;
}

void vo_param_default(vo_param* p)
{
    /* param(), viso.h:60.  base / calib are uninitialised in the reference; zero here. */
    std::memset(p, 0, sizeof(*p));
    p->ransac_iter = 50;
    p->inlier_threshold = 2;
    p->thresh = 1e-4;
}

double vo_sampson_distance(const double F[9], float p1x, float p1y, float p2x, float p2y)
{
    /* viso.cpp:655-666 */
    double Fx0 = F[0] * p1x + F[1] * p1y + F[2],
           Fx1 = F[3] * p1x + F[4] * p1y + F[5],
           Ftx0 = F[0] * p2x + F[3] * p2y + F[6],
           Ftx1 = F[1] * p2x + F[4] * p2y + F[7];
    float ad = (float)algebraic_distance(F, p1x, p1y, p2x, p2y); /* float truncation, :664 */
    return ad * ad / (Fx0 * Fx0 + Fx1 * Fx1 + Ftx0 * Ftx0 + Ftx1 * Ftx1);
}

int vo_radius_search(float qx, float qy, const float* kp2, int n2, float radius, int K, int32_t* nb, float* dists)
{
    std::vector<std::pair<float, int>> found;
    int total = radius_search_row(qx, qy, kp2, n2, radius, K, found, nb);
    if (dists)
        for (int j = 0; j < K; ++j) dists[j] = j < total ? found[j].first : -1.f;
    return total;
}

void vo_sort_matches(int32_t* matches, int n)
{
    /* viso.cpp:724 -- libstdc++ std::sort (introsort), unstable; comparator on dist only */
    Match3* m = reinterpret_cast<Match3*>(matches);
    std::sort(m, m + n, [](const Match3& a, const Match3& b) { return a.v[2] < b.v[2]; });
}

int vo_match_desc(const float* kp1, int n1, const float* kp2, int n2,
                  const float* d1, const float* d2, int dlen,
                  const vo_match_params* sp,
                  int32_t* matches, int32_t* n_matches,
                  int32_t* dense_idx, int32_t* dense_d1, int32_t* dense_d2, int32_t* dense_valid,
                  int64_t* n_sad)
{
    const int K = sp->max_neighbors;
    if (K < 1) return -1;
    const float radius = (float)sp->radius; /* radiusSearch(..., float radius, ...), viso.cpp:171-172 */
    std::vector<std::pair<float, int>> found;
    std::vector<int> nb(K);
    std::vector<Match3> out;
    int64_t sad_count = 0;
    for (int i = 0; i < n1; ++i) {
        /* --- radiusSearch, viso.cpp:685 */
        const float qx = kp1[2 * i], qy = kp1[2 * i + 1];
        radius_search_row(qx, qy, kp2, n2, radius, K, found, nb.data());

        /* --- scan, viso.cpp:688-710 */
        double best_d1 = DBL_MAX, best_d2 = DBL_MAX;
        int best_idx = -1;
        for (int j = 0; j < K && nb[j] > 0; ++j) { /* nind>0: index 0 terminates the scan (:693) */
            const int nind = nb[j];
            if (sp->enforce_epipolar) {
                double sd = vo_sampson_distance(sp->F, qx, qy, kp2[2 * nind], kp2[2 * nind + 1]);
                if (!std::isfinite(sd) || sd > sp->sampson_thresh) continue;
            }
            const float* a = d2 + (size_t)nind * dlen;
            const float* b = d1 + (size_t)i * dlen;
            double d = 0;
            for (int k = 0; k < dlen; ++k) d += (double)std::fabs(a[k] - b[k]); /* cv::norm(a-b, NORM_L1), :702 */
            ++sad_count;
            if (d <= best_d1) {
                best_d2 = best_d1;
                best_d1 = d;
                best_idx = nind;
            } else if (d <= best_d2)
                best_d2 = d;
        }
        int valid = 0;
        if (best_idx >= 0) {
            if (sp->enforce_2nd_best) {
                if (best_d1 < best_d2 * sp->ratio_2nd_best) valid = 1;
            } else
                valid = 1;
        }
        if (valid) {
            Match3 m;
            m.v[0] = i; m.v[1] = best_idx; m.v[2] = (int)best_d1; /* Match(i,best_idx,best_d1): double->int */
            out.push_back(m);
        }
        if (dense_idx) dense_idx[i] = best_idx;
        if (dense_d1) dense_d1[i] = best_d1 == DBL_MAX ? INT32_MAX : (int)best_d1;
        if (dense_d2) dense_d2[i] = best_d2 == DBL_MAX ? INT32_MAX : (int)best_d2;
        if (dense_valid) dense_valid[i] = valid;
    }
    std::sort(out.begin(), out.end(), [](const Match3& a, const Match3& b) { return a.v[2] < b.v[2]; }); /* :724 */
    if (matches) std::memcpy(matches, out.data(), out.size() * sizeof(Match3));
    if (n_matches) *n_matches = (int)out.size();
    if (n_sad) *n_sad = sad_count;
    return 0;
}

int vo_match_circle(const int32_t* mlr, int nlr, const int32_t* mlrp, int nlrp,
                    const int32_t* m11, int n11, const int32_t* m22, int n22,
                    int32_t* circ4, int32_t* pcl3)
{
    /* viso.cpp:207-243, literal nested scans */
    int c = 0;
    for (int i = 0; i < nlr; ++i) {
        int ileft = mlr[3 * i], iright = mlr[3 * i + 1];
        for (int j = 0; j < n11; ++j) {
            if (m11[3 * j] == ileft) {
                int ileft_prev = m11[3 * j + 1];
                for (int k = 0; k < nlrp; ++k) {
                    if (mlrp[3 * k] == ileft_prev) {
                        int iright_prev = mlrp[3 * k + 1];
                        for (int l = 0; l < n22; ++l) {
                            if (m22[3 * l + 1] == iright_prev) {
                                if (m22[3 * l] == iright) {
                                    circ4[4 * c] = ileft; circ4[4 * c + 1] = iright;
                                    circ4[4 * c + 2] = ileft_prev; circ4[4 * c + 3] = iright_prev;
                                    pcl3[3 * c] = i; pcl3[3 * c + 1] = k; pcl3[3 * c + 2] = 0; /* Match(i,k) */
                                    ++c;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    return c;
}

void vo_collect_matches(const float* kp1, const float* kp2, const int32_t* matches, int m, double* x)
{
    /* viso.cpp:501-514 */
    for (int i = 0; i < m; ++i) {
        int i1 = matches[3 * i], i2 = matches[3 * i + 1];
        x[0 * m + i] = kp1[2 * i1];
        x[1 * m + i] = kp1[2 * i1 + 1];
        x[2 * m + i] = kp2[2 * i2];
        x[3 * m + i] = kp2[2 * i2 + 1];
    }
}

void vo_triangulate_rectified_f64(const double* x, int m, double f, double base, double cu, double cv, double* X)
{
    /* viso.cpp:1137-1154, T=double; no disparity clamp */
    for (int i = 0; i < m; ++i) {
        double d = x[0 * m + i] - x[2 * m + i];
        X[0 * m + i] = base * (x[0 * m + i] - cu) / d;
        X[1 * m + i] = base * (x[1 * m + i] - cv) / d;
        X[2 * m + i] = f * base / d;
    }
}

void vo_triangulate_rectified_f32(const float* x1, const float* x2, int m, double f, double base,
                                  double c1u, double c1v, float* X)
{
    /* mvg.cpp:172-192 */
    for (int i = 0; i < m; ++i) {
        double d = std::max(x1[0 * m + i] - x2[0 * m + i], 0.0001f);
        X[0 * m + i] = (float)((x1[0 * m + i] - c1u) * base / d);
        X[1 * m + i] = (float)((x1[1 * m + i] - c1v) * base / d);
        X[2 * m + i] = (float)(f * base / d);
    }
}

void vo_compute_J(const double* X, const double* observe, int n, const double tr[6], const vo_param* param,
                  const int32_t* active, int na, double* J, double* predict, double* residual)
{
    /* viso.cpp:1401-1497 */
    double rx = tr[0], ry = tr[1], rz = tr[2];
    double tx = tr[3], ty = tr[4], tz = tr[5];
    double sx = sin(rx), cx = cos(rx), sy = sin(ry);
    double cy = cos(ry), sz = sin(rz), cz = cos(rz);

    double r00 = +cy * cz;                double r01 = -cy * sz;                double r02 = +sy;
    double r10 = +sx * sy * cz + cx * sz; double r11 = -sx * sy * sz + cx * cz; double r12 = -sx * cy;
    double r20 = -cx * sy * cz + sx * sz; double r21 = +cx * sy * sz + sx * cz; double r22 = +cx * cy;
    double rdrx10 = +cx * sy * cz - sx * sz; double rdrx11 = -cx * sy * sz - sx * cz; double rdrx12 = -cx * cy;
    double rdrx20 = +sx * sy * cz + cx * sz; double rdrx21 = -sx * sy * sz + cx * cz; double rdrx22 = -sx * cy;
    double rdry00 = -sy * cz;      double rdry01 = +sy * sz;      double rdry02 = +cy;
    double rdry10 = +sx * cy * cz; double rdry11 = -sx * cy * sz; double rdry12 = +sx * sy;
    double rdry20 = -cx * cy * cz; double rdry21 = +cx * cy * sz; double rdry22 = -cx * sy;
    double rdrz00 = -cy * sz;                double rdrz01 = -cy * cz;
    double rdrz10 = -sx * sy * sz + cx * cz; double rdrz11 = -sx * sy * cz - cx * sz;
    double rdrz20 = +cx * sy * sz + sx * cz; double rdrz21 = +cx * sy * cz - sx * sz;

    const double f = param->f, cu = param->cu, cv = param->cv;
    double X1p, Y1p, Z1p, X1c, Y1c, Z1c, X2c, X1cd = 0, Y1cd = 0, Z1cd = 0;
    for (int i = 0; i < na; i++) {
        X1p = X[0 * n + active[i]];
        Y1p = X[1 * n + active[i]];
        Z1p = X[2 * n + active[i]];

        X1c = r00 * X1p + r01 * Y1p + r02 * Z1p + tx;
        Y1c = r10 * X1p + r11 * Y1p + r12 * Z1p + ty;
        Z1c = r20 * X1p + r21 * Y1p + r22 * Z1p + tz;

        /* weight read from column i, NOT active[i] (viso.cpp:1449) */
        double weight = 1.0 / (fabs(observe[0 * n + i] - cu) / fabs(cu) + 0.05);

        X2c = X1c - param->base;
        for (int j = 0; j < 6; j++) {
            switch (j) {
            case 0: X1cd = 0;
                Y1cd = rdrx10 * X1p + rdrx11 * Y1p + rdrx12 * Z1p;
                Z1cd = rdrx20 * X1p + rdrx21 * Y1p + rdrx22 * Z1p;
                break;
            case 1: X1cd = rdry00 * X1p + rdry01 * Y1p + rdry02 * Z1p;
                Y1cd = rdry10 * X1p + rdry11 * Y1p + rdry12 * Z1p;
                Z1cd = rdry20 * X1p + rdry21 * Y1p + rdry22 * Z1p;
                break;
            case 2: X1cd = rdrz00 * X1p + rdrz01 * Y1p;
                Y1cd = rdrz10 * X1p + rdrz11 * Y1p;
                Z1cd = rdrz20 * X1p + rdrz21 * Y1p;
                break;
            case 3: X1cd = 1; Y1cd = 0; Z1cd = 0; break;
            case 4: X1cd = 0; Y1cd = 1; Z1cd = 0; break;
            case 5: X1cd = 0; Y1cd = 0; Z1cd = 1; break;
            }
            J[(4 * i + 0) * 6 + j] = weight * f * (X1cd * Z1c - X1c * Z1cd) / (Z1c * Z1c);
            J[(4 * i + 1) * 6 + j] = weight * f * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c);
            J[(4 * i + 2) * 6 + j] = weight * f * (X1cd * Z1c - X2c * Z1cd) / (Z1c * Z1c);
            J[(4 * i + 3) * 6 + j] = weight * f * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c);
        }
        predict[0 * na + i] = f * X1c / Z1c + cu;
        predict[1 * na + i] = f * Y1c / Z1c + cv;
        predict[2 * na + i] = f * X2c / Z1c + cu;
        predict[3 * na + i] = f * Y1c / Z1c + cv;

        residual[4 * i + 0] = weight * (observe[0 * n + active[i]] - predict[0 * na + i]);
        residual[4 * i + 1] = weight * (observe[1 * n + active[i]] - predict[1 * na + i]);
        residual[4 * i + 2] = weight * (observe[2 * n + active[i]] - predict[2 * na + i]);
        residual[4 * i + 3] = weight * (observe[3 * n + active[i]] - predict[3 * na + i]);
    }
}

int vo_get_inliers(const double* X, const double* observe, int n, const double tr[6], const vo_param* param,
                   int32_t* inliers, double* rms, double* min_margin)
{
    /* viso.cpp:1509-1537 */
    std::vector<int32_t> active(n);
    for (int i = 0; i < n; ++i) active[i] = i;
    std::vector<double> J((size_t)4 * n * 6), residual((size_t)4 * n), predict((size_t)4 * n);
    vo_compute_J(X, observe, n, tr, param, active.data(), n, J.data(), predict.data(), residual.data());
    int cnt = 0;
    double err2 = 0, margin = DBL_MAX;
    const double thr2 = param->inlier_threshold * param->inlier_threshold;
    for (int i = 0; i < n; ++i) {
        err2 = pow(observe[0 * n + i] - predict[0 * n + i], 2) +
               pow(observe[1 * n + i] - predict[1 * n + i], 2) +
               pow(observe[2 * n + i] - predict[2 * n + i], 2) +
               pow(observe[3 * n + i] - predict[3 * n + i], 2);
        if (err2 < thr2) inliers[cnt++] = i;
        double mg = fabs(err2 - thr2);
        if (mg < margin) margin = mg;
    }
    if (rms) *rms = n > 0 ? sqrt(err2 / n) : 0; /* last point only, as in the reference (:1535) */
    if (min_margin) *min_margin = margin;
    return cnt;
}

int vo_solve_lu(const double* A, const double* b, int n, double* x)
{
    /* cv::solve(A,b,x,DECOMP_LU) for n>3: copy A, copy b into x, hal::LU64f */
    std::vector<double> a(A, A + (size_t)n * n);
    for (int i = 0; i < n; i++) x[i] = b[i];
    return lu_impl(a.data(), n, n, x, 1, 1) != 0;
}

int vo_invert_lu(const double* A, int n, double* Ainv)
{
    /* cv::invert(DECOMP_LU) general path (n>3): dst = I; LU64f(src copy, dst) ; zero on failure */
    std::vector<double> a(A, A + (size_t)n * n);
    for (int i = 0; i < n * n; i++) Ainv[i] = 0;
    for (int i = 0; i < n; i++) Ainv[i * n + i] = 1;
    int ok = lu_impl(a.data(), n, n, Ainv, n, n) != 0;
    if (!ok)
        for (int i = 0; i < n * n; i++) Ainv[i] = 0;
    return ok;
}

double vo_determinant(const double* A, int n)
{
    /* cv::determinant, rows>3 branch: LU then product of the diagonal times the permutation sign */
    std::vector<double> a(A, A + (size_t)n * n);
    double result = lu_impl(a.data(), n, n, nullptr, 0, 0);
    if (result != 0)
        for (int i = 0; i < n; i++) result *= a[i * n + i];
    return result;
}

void vo_mul_transposed(const double* J, int rows, double JtJ[36])
{
    /* cv::mulTransposed(J,JtJ,true), viso.cpp:1599 -- MulTransposedR<double,double>: upper triangle, each
     * entry a sequential sum over rows starting from 0.0, then cv::completeSymm mirrors it
     * (pinned bit-exact vs cv2: tests/golden/linalg.npz mt_*) */
    for (int i = 0; i < 6; i++)
        for (int j = i; j < 6; j++) {
            double s = 0;
            for (int k = 0; k < rows; k++) s += J[k * 6 + i] * J[k * 6 + j];
            JtJ[i * 6 + j] = s;
            JtJ[j * 6 + i] = s;
        }
}

void vo_Jt_times_r(const double* J, const double* r, int rows, double Jtr[6])
{
    /* J.t()*residual, viso.cpp:1602 -- cv::gemm's accumulation order is build dependent; the oracle
     * fixes row order (cv2 4.13 agrees to ~1e-15 relative) */
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int k = 0; k < rows; k++) s += J[k * 6 + i] * r[k];
        Jtr[i] = s;
    }
}

int vo_minimize_reproj(const double* X, const double* observe, int n, double tr[6], const vo_param* param,
                       const int32_t* active, int na, int32_t* iters)
{
    /* viso.cpp:1583-1623 */
    std::vector<double> J((size_t)4 * na * 6), residual((size_t)4 * na), predict((size_t)4 * na);
    const double step_size = 1.0f;
    if (iters) *iters = 0;
    for (int it = 0; it < 100; ++it) {
        if (iters) *iters = it + 1;
        vo_compute_J(X, observe, n, tr, param, active, na, J.data(), predict.data(), residual.data());
        double JtJ[36], Jtr[6], p_gn[6];
        vo_mul_transposed(J.data(), 4 * na, JtJ);          /* mulTransposed(J,JtJ,true), :1599 */
        vo_Jt_times_r(J.data(), residual.data(), 4 * na, Jtr); /* J.t()*residual, :1602 */
        if (!vo_solve_lu(JtJ, Jtr, 6, p_gn)) return 0;
        bool converged = true;
        for (int j = 0; j < 6; ++j) {
            if (fabs((double)(p_gn[j] > param->thresh))) { /* sic: fabs(p > thresh), :1610 */
                converged = false;
                break;
            }
        }
        if (converged) return 1; /* without applying p_gn, :1616-1617 */
        for (int j = 0; j < 6; ++j) tr[j] = tr[j] + step_size * p_gn[j];
    }
    return 0;
}

int vo_ransac_minimize_reproj(const double* X, const double* observe, int n, const vo_param* param,
                              const int32_t* sample_table,
                              double best_tr[6], int32_t* best_inliers, int32_t* n_best,
                              double* hyp_tr, int32_t* hyp_ok, int32_t* hyp_count, int32_t* best_hyp)
{
    /* viso.cpp:1543-1580 */
    std::vector<int32_t> cur(n > 0 ? n : 1), best;
    double tr[6];
    int bh = -1;
    for (int i = 0; i < param->ransac_iter; ++i) {
        for (int j = 0; j < 6; j++) tr[j] = 0;
        const int32_t* sample = sample_table + 3 * i; /* randomsample(3,X.cols,sample), :1558 */
        int ok = vo_minimize_reproj(X, observe, n, tr, param, sample, 3, nullptr);
        int cnt = -1;
        if (ok) {
            cnt = vo_get_inliers(X, observe, n, tr, param, cur.data(), nullptr, nullptr);
            if ((size_t)cnt > best.size()) { /* strict >, first best wins, :1564 */
                best.assign(cur.begin(), cur.begin() + cnt);
                for (int j = 0; j < 6; j++) best_tr[j] = tr[j];
                bh = i;
            }
        }
        if (hyp_tr) for (int j = 0; j < 6; j++) hyp_tr[6 * i + j] = tr[j];
        if (hyp_ok) hyp_ok[i] = ok;
        if (hyp_count) hyp_count[i] = cnt;
    }
    if (best_hyp) *best_hyp = bh;
    if (best.size() < 6 ||
        !vo_minimize_reproj(X, observe, n, best_tr, param, best.data(), (int)best.size(), nullptr)) {
        /* the reference leaves best_inliers = RANSAC support set on failure */
        if (best_inliers) std::copy(best.begin(), best.end(), best_inliers);
        if (n_best) *n_best = (int)best.size();
        return 0;
    }
    int cnt = vo_get_inliers(X, observe, n, best_tr, param, cur.data(), nullptr, nullptr);
    if (best_inliers) std::copy(cur.begin(), cur.begin() + cnt, best_inliers);
    if (n_best) *n_best = cnt;
    return 1;
}

void vo_randomsample_table(uint32_t seed, int H, int N, int32_t* table)
{
    /* viso.cpp:87-107, Knuth Algorithm S (n=3), one generator stream instead of a fresh
     * random_device-seeded one per call */
    std::mt19937 gen(seed);
    std::uniform_real_distribution<> dis(0, 1);
    for (int h = 0; h < H; ++h) {
        int t = 0, m = 0, n = 3;
        while (m < n) {
            double u = dis(gen);
            if ((N - t) * u >= n - m) {
                t++;
            } else {
                table[3 * h + m] = t;
                t++; m++;
            }
        }
    }
}

void vo_samples_from_seeds(const uint32_t* seeds, int H, int N, int32_t* table)
{
    for (int h = 0; h < H; ++h) {
        uint32_t r0 = seeds[3 * h], r1 = seeds[3 * h + 1], r2 = seeds[3 * h + 2];
        int a = (int)(((uint64_t)r0 * (uint64_t)N) >> 32);
        int b = (int)(((uint64_t)r1 * (uint64_t)(N - 1)) >> 32);
        int c = (int)(((uint64_t)r2 * (uint64_t)(N - 2)) >> 32);
        if (b >= a) b++;
        int lo = a < b ? a : b, hi = a < b ? b : a;
        if (c >= lo) c++;
        if (c >= hi) c++;
        int s0 = lo, s1 = hi, s2 = c;
        if (s2 < s0) { int t = s2; s2 = s1; s1 = s0; s0 = t; }
        else if (s2 < s1) { int t = s2; s2 = s1; s1 = t; }
        table[3 * h] = s0; table[3 * h + 1] = s1; table[3 * h + 2] = s2;
    }
}

void vo_tr2mat(const double tr[6], double T[16])
{
    /* viso.cpp:109-133 */
    double rx = tr[0], ry = tr[1], rz = tr[2], tx = tr[3], ty = tr[4], tz = tr[5];
    double sx = sin(rx), cx = cos(rx), sy = sin(ry), cy = cos(ry), sz = sin(rz), cz = cos(rz);
    T[0] = +cy * cz;                T[1] = -cy * sz;                T[2] = +sy;       T[3] = tx;
    T[4] = +sx * sy * cz + cx * sz; T[5] = -sx * sy * sz + cx * cz; T[6] = -sx * cy;  T[7] = ty;
    T[8] = -cx * sy * cz + sx * sz; T[9] = +cx * sy * sz + sx * cz; T[10] = +cx * cy; T[11] = tz;
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

int vo_pose_update(const double pose[16], const double tr[6], double pose_out[16])
{
    /* viso.cpp:1315-1321: pose = pose * tr_mat.inv() */
    double T[16], Ti[16];
    vo_tr2mat(tr, T);
    if (!vo_invert_lu(T, 4, Ti)) return 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += pose[i * 4 + k] * Ti[k * 4 + j];
            pose_out[i * 4 + j] = s;
        }
    return 1;
}

void vo_F_from_P(const double P1[12], const double P2[12], int normalise, double F[9])
{
    /* mvg.h:41-66: Xj / Yj are P1 / P2 with row j omitted (cyclic order); F(r,c) = det([X_c; Y_r]) */
    static const int rows[3][2] = { {1, 2}, {2, 0}, {0, 1} };
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double M[16];
            for (int k = 0; k < 4; k++) {
                M[0 * 4 + k] = P1[rows[c][0] * 4 + k];
                M[1 * 4 + k] = P1[rows[c][1] * 4 + k];
                M[2 * 4 + k] = P2[rows[r][0] * 4 + k];
                M[3 * 4 + k] = P2[rows[r][1] * 4 + k];
            }
            F[r * 3 + c] = vo_determinant(M, 4);
        }
    if (normalise && F[8] > DBL_MIN) {
        /* viso.cpp:1177-1180, `F /= F.at<double>(2,2)`: OpenCV's operator/=(Mat&, double) is a.convertTo(a, -1, 1./s)
         * (opencv2/core/mat.inl.hpp) -- every element is MULTIPLIED by the reciprocal */
        const double inv = 1. / F[8];
        for (int i = 0; i < 9; i++) F[i] = F[i] * inv;
    }
}

int vo_project_points(const double* X, int n, const double P[12], double* x)
{
    /* viso.cpp:326-333 with misc.h:90-124: Xh = [X;1]; xh = P*Xh; x = xh[0:2]/xh[2];
     * h2e throws overflow_error when isEqual(|w|,0) (misc.cpp:3-8: |w-0| <= 1e-6*|w| <=> w==0) */
    for (int i = 0; i < n; i++) {
        double Xh[4] = { X[0 * n + i], X[1 * n + i], X[2 * n + i], 1.0 };
        double xh[3];
        for (int r = 0; r < 3; r++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += P[r * 4 + k] * Xh[k];
            xh[r] = s;
        }
        if (std::fabs(xh[2]) == 0) return -1;
        x[0 * n + i] = xh[0] / xh[2];
        x[1 * n + i] = xh[1] / xh[2];
    }
    return 0;
}

void vo_extract_descriptors(const float* sob, int h, int w, const float* kp, int n, int radius, float* d)
{
    /* viso.cpp:1011-1023 */
    const int side = 2 * radius + 1;
    for (int k = 0; k < n; ++k) {
        int px = (int)lrintf(kp[2 * k]), py = (int)lrintf(kp[2 * k + 1]); /* Point2i p = kp.pt: saturate_cast<int> = cvRound (half to even) */
        int col = 0;
        for (int i = -radius; i <= radius; i += 1)
            for (int j = -radius; j <= radius; j += 1, ++col) {
                float val = (py + i > 0 && py + i < h && px + j > 0 && px + j < w) ? sob[(size_t)(py + i) * w + (px + j)] : 0;
                d[(size_t)k * side * side + col] = val;
            }
    }
}

/* ---- front end: Sobel-x, Harris response, binned detector ---- */

static inline int reflect101(int i, int n)
{
    /* cv::borderInterpolate(BORDER_REFLECT_101) for |overshoot| < n */
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

void vo_sobel_x(const uint8_t* img, int h, int w, float* sob)
{
    /* viso.cpp:1010: 3x3 Sobel, kx = [-1 0 1], ky = [1 2 1]; integers, exact in float */
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = img + (size_t)reflect101(y - 1, h) * w;
        const uint8_t* r1 = img + (size_t)y * w;
        const uint8_t* r2 = img + (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; ++x) {
            const int xl = reflect101(x - 1, w), xr = reflect101(x + 1, w);
            const int v = (r0[xr] - r0[xl]) + 2 * (r1[xr] - r1[xl]) + (r2[xr] - r2[xl]);
            sob[(size_t)y * w + x] = (float)v;
        }
    }
}

void vo_harris_response(const uint8_t* img, int h, int w, float k, float* resp)
{
    /* see viso_oracle.h for the canonical operation order (cv::cornerHarris, viso.cpp:930) */
    const double s = 1.0 / ((double)(1 << 4) * 3 * 255.0);
    const float f0 = (float)(6.0 * s), f1 = (float)(4.0 * s), f2 = (float)(1.0 * s);
    const size_t N = (size_t)h * w;
    std::vector<float> r(N), t(N), xx(N), xy(N), yy(N), rsa(N), rsb(N), rsc(N);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const uint8_t* row = img + (size_t)y * w;
            const float pl2 = row[reflect101(x - 2, w)], pl1 = row[reflect101(x - 1, w)], p0 = row[x],
                        pr1 = row[reflect101(x + 1, w)], pr2 = row[reflect101(x + 2, w)];
            r[(size_t)y * w + x] = (pr2 - pl2) + 2.0f * (pr1 - pl1);
            float tt = f0 * p0;
            tt = tt + f1 * (pl1 + pr1);
            tt = tt + f2 * (pl2 + pr2);
            t[(size_t)y * w + x] = tt;
        }
    for (int y = 0; y < h; ++y) {
        const size_t ym2 = (size_t)reflect101(y - 2, h) * w, ym1 = (size_t)reflect101(y - 1, h) * w,
                     yp1 = (size_t)reflect101(y + 1, h) * w, yp2 = (size_t)reflect101(y + 2, h) * w, y0 = (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float dx = f0 * r[y0 + x];
            dx = dx + f1 * (r[ym1 + x] + r[yp1 + x]);
            dx = dx + f2 * (r[ym2 + x] + r[yp2 + x]);
            float dy = 2.0f * (t[yp1 + x] - t[ym1 + x]);
            dy = dy + (t[yp2 + x] - t[ym2 + x]);
            xx[y0 + x] = dx * dx; xy[y0 + x] = dx * dy; yy[y0 + x] = dy * dy;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const size_t o = (size_t)y * w, l = o + reflect101(x - 1, w), c = o + x, rr = o + reflect101(x + 1, w);
            rsa[c] = (xx[l] + xx[c]) + xx[rr];
            rsb[c] = (xy[l] + xy[c]) + xy[rr];
            rsc[c] = (yy[l] + yy[c]) + yy[rr];
        }
    for (int y = 0; y < h; ++y) {
        const size_t u = (size_t)reflect101(y - 1, h) * w, o = (size_t)y * w, d = (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; ++x) {
            const float a = (rsa[u + x] + rsa[o + x]) + rsa[d + x];
            const float b = (rsb[u + x] + rsb[o + x]) + rsb[d + x];
            const float c = (rsc[u + x] + rsc[o + x]) + rsc[d + x];
            const float tr = a + c;
            resp[o + x] = (a * c - b * b) - (k * tr) * tr;
        }
    }
}

int vo_detect_harris_binned(const uint8_t* img, int h, int w, int n, int nbinx, int nbiny, float k, int order_rule,
                            float* kp_xy, float* kp_resp)
{
    /* viso.cpp:925-976 */
    std::vector<float> resp((size_t)h * w);
    vo_harris_response(img, h, w, k, resp.data());
    const int stridex = w / nbinx, stridey = h / nbiny;               /* :932-933 */
    if (stridex <= 0 || stridey <= 0) return -1;                       /* :934 assert */
    struct elem {
        int x, y; float val;
        bool operator<(const elem& o) const { return val < o.val; }    /* :941 */
    };
    const int per = n / (nbinx * nbiny);                               /* :943 */
    std::vector<elem> v;
    v.reserve((size_t)stridex * stridey);
    int out = 0;
    for (int binx = 0; binx < nbinx; ++binx)
        for (int biny = 0; biny < nbiny; ++biny) {
            for (int x = binx * stridex; x < (binx + 1) * stridex && x < w; ++x)
                for (int y = biny * stridey; y < (biny + 1) * stridey && y < h; ++y) {
                    const float response = fabsf(resp[(size_t)y * w + x]);
                    /* isEqual(response, 0.f), misc.cpp:10-15: |r - 0| <= 1e-6 |r|  <=>  r == 0 (NaN is kept) */
                    if (fabsf(response - 0.0f) <= 1e-6f * fabsf(response)) continue;
                    v.push_back(elem{x, y, response});
                }
            const int m = ((int)v.size() > per) ? (int)v.size() - per : 0; /* :961 */
            if (order_rule == 0) {
                if (m > 0) std::nth_element(v.begin(), v.begin() + m, v.end()); /* :963 */
            } else {
                std::sort(v.begin(), v.end(), [](const elem& a, const elem& b) {
                    if (a.val != b.val) return a.val < b.val;
                    if (a.x != b.x) return a.x < b.x;
                    return a.y < b.y;
                });
            }
            for (size_t i = m; i < v.size(); ++i, ++out) {
                kp_xy[2 * out] = (float)v[i].x; kp_xy[2 * out + 1] = (float)v[i].y; /* :967 */
                if (kp_resp) kp_resp[out] = v[i].val;
            }
            v.clear();
        }
    return out;
}

/* ---- lower-priority geometry; not bit-pinned: SVD vectors are unique only up to rounding ---- */

static void jacobi_eig_sym(double* A, int n, double* V)
{
    /* cyclic Jacobi on symmetric A (n<=4), V accumulates eigenvectors (columns) */
    for (int i = 0; i < n * n; i++) V[i] = 0;
    for (int i = 0; i < n; i++) V[i * n + i] = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A[p * n + q];
                if (apq == 0) continue;
                double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
}

void vo_triangulate_dlt(const float* x1, const float* x2, int m, const double P1[12], const double P2[12], float* X)
{
    /* mvg.cpp:124-169.  cv::SVD's last right singular vector == eigenvector of A^T A with the smallest
     * eigenvalue (sign fixed by the division by vt(3,3)) */
    for (int i = 0; i < m; ++i) {
        double A[16];
        for (int k = 0; k < 4; k++) {
            A[0 * 4 + k] = x1[0 * m + i] * P1[2 * 4 + k] - P1[0 * 4 + k];
            A[1 * 4 + k] = x1[1 * m + i] * P1[2 * 4 + k] - P1[1 * 4 + k];
            A[2 * 4 + k] = x2[0 * m + i] * P2[2 * 4 + k] - P2[0 * 4 + k];
            A[3 * 4 + k] = x2[1 * m + i] * P2[2 * 4 + k] - P2[1 * 4 + k];
        }
        double AtA[16], V[16];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += A[k * 4 + r] * A[k * 4 + c];
                AtA[r * 4 + c] = s;
            }
        jacobi_eig_sym(AtA, 4, V);
        int best = 0;
        for (int c = 1; c < 4; c++) if (AtA[c * 4 + c] < AtA[best * 4 + best]) best = c;
        double v[4] = { V[0 * 4 + best], V[1 * 4 + best], V[2 * 4 + best], V[3 * 4 + best] };
        double d = (fabs(v[3]) < DBL_MIN) ? 1.0 : v[3];
        X[0 * m + i] = (float)((float)v[0] / d);
        X[1 * m + i] = (float)((float)v[1] / d);
        X[2 * m + i] = (float)((float)v[2] / d);
    }
}

void vo_solve_rigid_motion(const float* A, const float* B, int n, float T[16])
{
    /* estimation.cpp:29-51: C = A_zm * B_zm^T, R = U diag(1,1,det(UV^T)) V^T, t = mean1 - R*mean2 */
    double m1[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
    for (int r = 0; r < 3; r++) {
        for (int i = 0; i < n; i++) { m1[r] += A[r * n + i]; m2[r] += B[r * n + i]; }
        m1[r] /= n; m2[r] /= n;
    }
    double C[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int i = 0; i < n; i++) s += (A[r * n + i] - m1[r]) * (B[c * n + i] - m2[c]);
            C[r * 3 + c] = s;
        }
    /* SVD of C via eigen-decomposition of C^T C (V) and U = C V / sigma with Gram-Schmidt completion */
    double CtC[9], V[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += C[k * 3 + r] * C[k * 3 + c];
            CtC[r * 3 + c] = s;
        }
    jacobi_eig_sym(CtC, 3, V);
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int a, int b) { return CtC[a * 3 + a] > CtC[b * 3 + b]; });
    double Vs[9], U[9];
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) Vs[r * 3 + c] = V[r * 3 + ord[c]];
    for (int c = 0; c < 3; c++) {
        double u[3];
        for (int r = 0; r < 3; r++) { double s = 0; for (int k = 0; k < 3; k++) s += C[r * 3 + k] * Vs[k * 3 + c]; u[r] = s; }
        for (int pc = 0; pc < c; pc++) {
            double dot = 0; for (int r = 0; r < 3; r++) dot += u[r] * U[r * 3 + pc];
            for (int r = 0; r < 3; r++) u[r] -= dot * U[r * 3 + pc];
        }
        double nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (nrm < 1e-12) {
            /* rank-deficient: complete with a cross product / any orthogonal vector */
            if (c == 2) {
                u[0] = U[1 * 3 + 0] * U[2 * 3 + 1] - U[2 * 3 + 0] * U[1 * 3 + 1];
                u[1] = U[2 * 3 + 0] * U[0 * 3 + 1] - U[0 * 3 + 0] * U[2 * 3 + 1];
                u[2] = U[0 * 3 + 0] * U[1 * 3 + 1] - U[1 * 3 + 0] * U[0 * 3 + 1];
            } else {
                double e[3] = {0, 0, 0}; e[c == 0 ? 0 : (fabs(U[0]) < 0.9 ? 0 : 1)] = 1;
                for (int pc = 0; pc < c; pc++) {
                    double dot = 0; for (int r = 0; r < 3; r++) dot += e[r] * U[r * 3 + pc];
                    for (int r = 0; r < 3; r++) e[r] -= dot * U[r * 3 + pc];
                }
                for (int r = 0; r < 3; r++) u[r] = e[r];
            }
            nrm = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        }
        for (int r = 0; r < 3; r++) U[r * 3 + c] = u[r] / nrm;
    }
    double UVt[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double s = 0; for (int k = 0; k < 3; k++) s += U[r * 3 + k] * Vs[c * 3 + k]; UVt[r * 3 + c] = s;
    }
    double det = UVt[0] * (UVt[4] * UVt[8] - UVt[5] * UVt[7]) - UVt[1] * (UVt[3] * UVt[8] - UVt[5] * UVt[6]) +
                 UVt[2] * (UVt[3] * UVt[7] - UVt[4] * UVt[6]);
    double R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += U[r * 3 + k] * (k == 2 ? det : 1.0) * Vs[c * 3 + k];
        R[r * 3 + c] = s;
    }
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) T[r * 4 + c] = (float)R[r * 3 + c];
        T[r * 4 + 3] = (float)(m1[r] - (R[r * 3 + 0] * m2[0] + R[r * 3 + 1] * m2[1] + R[r * 3 + 2] * m2[2]));
    }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

int vo_sequence(int n_frames, const int32_t* nL, const int32_t* nR, const int64_t* offL, const int64_t* offR,
                const float* kpL, const float* kpR, const float* dL, const float* dR, int dlen,
                const double P1[12], const double P2[12], const vo_param* param_in,
                const uint32_t* seeds,
                vo_record* records,
                int32_t* lr_matches, int32_t* lr_count,
                int32_t* m11_dense, int32_t* m22_dense,
                int32_t* circ, int32_t* inliers_out,
                double* poses, int32_t* n_poses)
{
    /* viso.cpp:1167-1330 without detection / description / debug images */
    double F[9];
    vo_F_from_P(P1, P2, 1, F);                       /* :1176-1180 */
    vo_param param = *param_in;
    param.base = fabs(P2[3] / P2[0]);                 /* :1184 */
    param.f = P1[0]; param.cu = P1[2]; param.cv = P1[6]; /* :1185-1187 */
    vo_match_params mp_lr, mp_t;
    vo_match_params_stereo(&mp_lr, F);
    vo_match_params_temporal(&mp_t);
    const int H = param.ransac_iter;

    std::vector<int32_t> match_lr, match_lr_prev, m11, m22;
    std::vector<double> x, X, X_prev;
    int n_lr = 0, n_lr_prev = 0;
    double pose[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int np = 0;
    if (poses) { std::memcpy(poses, pose, sizeof(pose)); }
    np = 1;                                           /* :1189-1190 */
    bool first = true;
    for (int t = 0; t < n_frames; ++t) {
        std::memset(&records[t], 0, sizeof(vo_record));
        records[t].best_hyp = -1;
        if (!first) {                                  /* :1208-1222 */
            match_lr_prev.swap(match_lr); n_lr_prev = n_lr;
            X_prev.swap(X);
        }
        const float* kp1 = kpL + 2 * offL[t]; const float* kp2 = kpR + 2 * offR[t];
        const float* d1 = dL + (size_t)offL[t] * dlen; const float* d2 = dR + (size_t)offR[t] * dlen;
        match_lr.assign((size_t)3 * std::max(nL[t], 1), 0);
        vo_match_desc(kp1, nL[t], kp2, nR[t], d1, d2, dlen, &mp_lr, match_lr.data(), &n_lr,
                      nullptr, nullptr, nullptr, nullptr, nullptr);   /* :1240 */
        if (lr_matches) std::memcpy(lr_matches + 3 * offL[t], match_lr.data(), (size_t)n_lr * 12);
        if (lr_count) lr_count[t] = n_lr;
        x.assign((size_t)4 * std::max(n_lr, 1), 0); X.assign((size_t)3 * std::max(n_lr, 1), 0);
        vo_collect_matches(kp1, kp2, match_lr.data(), n_lr, x.data());              /* :1246 */
        vo_triangulate_rectified_f64(x.data(), n_lr, param.f, param.base, param.cu, param.cv, X.data()); /* :1247 */
        if (first) { first = false; continue; }        /* :1256-1260 */

        const float* kp1p = kpL + 2 * offL[t - 1]; const float* kp2p = kpR + 2 * offR[t - 1];
        const float* d1p = dL + (size_t)offL[t - 1] * dlen; const float* d2p = dR + (size_t)offR[t - 1] * dlen;
        int n11 = 0, n22 = 0;
        m11.assign((size_t)3 * std::max(nL[t], 1), 0); m22.assign((size_t)3 * std::max(nR[t], 1), 0);
        {
            std::vector<int32_t> di(nL[t]), dd1(nL[t]), dd2(nL[t]), dv(nL[t]);
            vo_match_desc(kp1, nL[t], kp1p, nL[t - 1], d1, d1p, dlen, &mp_t, m11.data(), &n11,
                          di.data(), dd1.data(), dd2.data(), dv.data(), nullptr);   /* :1264 */
            if (m11_dense)
                for (int i = 0; i < nL[t]; i++) {
                    int32_t* o = m11_dense + 4 * (offL[t] + i);
                    o[0] = di[i]; o[1] = dd1[i]; o[2] = dd2[i]; o[3] = dv[i];
                }
        }
        {
            std::vector<int32_t> di(nR[t]), dd1(nR[t]), dd2(nR[t]), dv(nR[t]);
            vo_match_desc(kp2, nR[t], kp2p, nR[t - 1], d2, d2p, dlen, &mp_t, m22.data(), &n22,
                          di.data(), dd1.data(), dd2.data(), dv.data(), nullptr);   /* :1275 */
            if (m22_dense)
                for (int i = 0; i < nR[t]; i++) {
                    int32_t* o = m22_dense + 4 * (offR[t] + i);
                    o[0] = di[i]; o[1] = dd1[i]; o[2] = dd2[i]; o[3] = dv[i];
                }
        }
        std::vector<int32_t> circ4((size_t)4 * std::max(n_lr, 1)), pcl((size_t)3 * std::max(n_lr, 1));
        int C = vo_match_circle(match_lr.data(), n_lr, match_lr_prev.data(), n_lr_prev,
                                m11.data(), n11, m22.data(), n22, circ4.data(), pcl.data()); /* :1282 */
        records[t].n_circ = C;
        if (circ) std::memcpy(circ + 4 * offL[t], circ4.data(), (size_t)C * 16);
        if (C < 3) continue;                           /* :1283-1288 */
        std::vector<double> Xp_c((size_t)3 * C), x_c((size_t)4 * C);
        for (int i = 0; i < C; ++i) {                  /* :1294-1305 */
            for (int r = 0; r < 4; r++) x_c[(size_t)r * C + i] = x[(size_t)r * n_lr + pcl[3 * i]];
            for (int r = 0; r < 3; r++) Xp_c[(size_t)r * C + i] = X_prev[(size_t)r * n_lr_prev + pcl[3 * i + 1]];
        }
        std::vector<int32_t> table((size_t)3 * H), inl(C);
        vo_samples_from_seeds(seeds + (size_t)t * H * 3, H, C, table.data());
        double tr[6] = {0, 0, 0, 0, 0, 0};             /* :1312 */
        int n_inl = 0, bh = -1;
        int ok = vo_ransac_minimize_reproj(Xp_c.data(), x_c.data(), C, &param, table.data(), tr, inl.data(), &n_inl,
                                           nullptr, nullptr, nullptr, &bh);  /* :1313 */
        records[t].ok = ok; records[t].n_inliers = n_inl; records[t].best_hyp = bh;
        for (int j = 0; j < 6; j++) records[t].tr[j] = tr[j];
        if (inliers_out) std::memcpy(inliers_out + offL[t], inl.data(), (size_t)n_inl * 4);
        if (ok) {                                      /* :1315-1321 */
            double np_[16];
            vo_pose_update(pose, tr, np_);
            std::memcpy(pose, np_, sizeof(pose));
            if (poses) std::memcpy(poses + 16 * np, pose, sizeof(pose));
            np++;
        }
    }
    if (n_poses) *n_poses = np;
    return 0;
}

} /* extern "C" */
