/*
 * oracle/shim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE (oracle/_ref build only; never on the product path).
 *
 * Overlay on compat/opencv2/core/core.hpp that adds the OpenCV routines the reference's hot path calls, so that the
 * reference's own functions can be compiled UNCHANGED from /root/reference/src into oracle/_ref/libviso_ref.so:
 *   cv::mulTransposed (viso.cpp:1599), cv::solve(DECOMP_LU) (viso.cpp:1602), cv::SVD (mvg.cpp:162).
 * OpenCV itself is third-party code that is absent from /root/reference and from this image; these are
 * restatements of its published algorithms, each pinned bit for bit (SVD: to tolerance) against OpenCV 4.13 through
 * the golden vectors under tests/golden/ (tests/test_ref_pin.py).
 */
#ifndef VISO_ORACLE_SHIM_OPENCV2_CORE_CORE_HPP_
#define VISO_ORACLE_SHIM_OPENCV2_CORE_CORE_HPP_

#include_next <opencv2/core/core.hpp>

namespace cv {

/* cv::mulTransposed(src, dst, aTa = true), CV_64F: MulTransposedR -- the upper triangle, each entry a sum over the
 * rows in order starting from 0, mirrored by completeSymm (modules/core/src/matmul.cpp) */
inline void mulTransposed(const Mat& src, Mat& dst, bool aTa)
{
    if (src.type() != CV_64FC1 || !aTa) throw std::invalid_argument("cv::mulTransposed (shim): CV_64F, aTa = true only");
    const int n = src.cols;
    Mat out(n, n, CV_64FC1);
    for (int i = 0; i < n; ++i)
        for (int j = i; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < src.rows; ++k) s += src.at<double>(k, i) * src.at<double>(k, j);
            out.at<double>(i, j) = s;
            out.at<double>(j, i) = s;
        }
    dst = out;
}

/* cv::solve(A, b, x, DECOMP_LU), square CV_64F with n > 3: A is copied, b is copied into x, hal::LU64f solves in place
 * (modules/core/src/lapack.cpp); false when a pivot is below 100 * DBL_EPSILON */
inline bool solve(const Mat& A, const Mat& b, Mat& x, int flags = DECOMP_LU)
{
    if (flags != DECOMP_LU || A.type() != CV_64FC1 || b.type() != CV_64FC1 || A.rows != A.cols || A.rows != b.rows || A.rows <= 3)
        throw std::invalid_argument("cv::solve (shim): square CV_64F systems with n > 3, DECOMP_LU only");
    Mat a = A.clone();
    Mat out = b.clone();
    const int ok = compat_detail::lu<double>(a.ptr<double>(), a.step / sizeof(double), a.rows, out.ptr<double>(),
                                             out.step / sizeof(double), out.cols, DBL_EPSILON * 100);
    x = out;
    return ok != 0;
}

/* cv::SVD of a small CV_64F matrix: w (decreasing), u, vt.  One-sided Jacobi like OpenCV's JacobiSVDImpl_; only the
 * null vector vt.row(n-1) is consumed (mvg.cpp:163-166), where the sign cancels in the division by vt(3,3). */
class SVD {
public:
    enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
    Mat u, w, vt;
    SVD() {}
    SVD(const Mat& src, int = 0)
    {
        if (src.type() != CV_64FC1 || src.rows < src.cols) throw std::invalid_argument("cv::SVD (shim): CV_64F with rows >= cols only");
        const int m = src.rows, n = src.cols;
        Mat W = src.clone(), V = Mat::eye(n, n, CV_64FC1);
        for (int sweep = 0; sweep < 60; ++sweep) {
            double off = 0;
            for (int p = 0; p < n - 1; ++p)
                for (int q = p + 1; q < n; ++q) {
                    double al = 0, be = 0, ga = 0;
                    for (int i = 0; i < m; ++i) { const double a = W.at<double>(i, p), b = W.at<double>(i, q); al += a * a; be += b * b; ga += a * b; }
                    if (std::abs(ga) <= 1e-300) continue;
                    off = std::max(off, std::abs(ga) / std::sqrt(std::max(al * be, 1e-300)));
                    const double zeta = (be - al) / (2 * ga);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::abs(zeta) + std::sqrt(1 + zeta * zeta));
                    const double c = 1 / std::sqrt(1 + t * t), s = c * t;
                    for (int i = 0; i < m; ++i) { const double a = W.at<double>(i, p), b = W.at<double>(i, q); W.at<double>(i, p) = c * a - s * b; W.at<double>(i, q) = s * a + c * b; }
                    for (int i = 0; i < n; ++i) { const double a = V.at<double>(i, p), b = V.at<double>(i, q); V.at<double>(i, p) = c * a - s * b; V.at<double>(i, q) = s * a + c * b; }
                }
            if (off < 1e-16) break;
        }
        std::vector<double> sv(n);
        std::vector<int> order(n);
        for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < m; ++i) s += W.at<double>(i, j) * W.at<double>(i, j); sv[j] = std::sqrt(s); order[j] = j; }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sv[a] > sv[b]; });
        w.create(n, 1, CV_64FC1); u.create(m, n, CV_64FC1); vt.create(n, n, CV_64FC1);
        for (int k = 0; k < n; ++k) {
            const int j = order[k];
            w.at<double>(k, 0) = sv[j];
            for (int i = 0; i < m; ++i) u.at<double>(i, k) = sv[j] > 0 ? W.at<double>(i, j) / sv[j] : 0.0;
            for (int i = 0; i < n; ++i) vt.at<double>(k, i) = V.at<double>(i, j);
        }
    }
};

} // namespace cv
#endif
