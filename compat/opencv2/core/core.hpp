/*
 * compat/opencv2/core/core.hpp -- stand-in for the OpenCV core types the libviso API is written in, for builds where
 * OpenCV itself is not installed (this image has no OpenCV C++ headers: SURVEY.md 8c).  It lets the reference's own
 * headers (src/viso.h, src/mvg.h, src/misc.h, src/estimation.h) and callers (src/kitti.cpp, test/test.cpp) compile
 * unchanged against libviso_b200/host.  With a real OpenCV on the include path this directory is simply not used.
 *
 * Scope: the container (2-D, ref-counted, row-major cv::Mat with at / ptr / row / clone / copyTo / create), the small
 * value types (Point_, Vec, Scalar, Size, KeyPoint, DataType), Mat_ with the comma initialiser, eager element-wise
 * and matrix operators, t(), inv() and determinant() for the once-per-pose host bookkeeping (viso.cpp:1176-1180,
 * 1315-1321), norm, cv::format and stream output.  Arithmetic follows OpenCV's published algorithms where the result
 * is observable: LU with partial pivoting (hal::LUImpl, eps = 100 * DBL_EPSILON) for inv / determinant of n > 3 and
 * the closed forms for n <= 3, `Mat /= s` as a multiplication by 1 / s (mat.inl.hpp: a.convertTo(a, -1, 1./s)).
 * Nothing of the per-frame hot path is implemented here -- that is libviso_b200.so.
 */
#ifndef VISO_COMPAT_OPENCV2_CORE_CORE_HPP_
#define VISO_COMPAT_OPENCV2_CORE_CORE_HPP_

#define VISO_B200_COMPAT_OPENCV 1

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <limits>
#include <map>
#include <set>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

using std::string;
typedef std::string String;

inline int cvRound(double v) { return (int)std::nearbyint(v); } /* round half to even, like the SSE2 cvRound */

template <class T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(double v) { int i = cvRound(v); return (uchar)(i < 0 ? 0 : i > 255 ? 255 : i); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }
template <> inline short saturate_cast<short>(double v) { int i = cvRound(v); return (short)(i < SHRT_MIN ? SHRT_MIN : i > SHRT_MAX ? SHRT_MAX : i); }
template <> inline ushort saturate_cast<ushort>(double v) { int i = cvRound(v); return (ushort)(i < 0 ? 0 : i > 65535 ? 65535 : i); }

template <class T> struct DataType { enum { depth = -1, channels = 1, type = -1 }; };
template <> struct DataType<uchar> { enum { depth = CV_8U, channels = 1, type = CV_8UC1 }; };
template <> struct DataType<signed char> { enum { depth = CV_8S, channels = 1, type = CV_8S }; };
template <> struct DataType<ushort> { enum { depth = CV_16U, channels = 1, type = CV_16U }; };
template <> struct DataType<short> { enum { depth = CV_16S, channels = 1, type = CV_16S }; };
template <> struct DataType<int> { enum { depth = CV_32S, channels = 1, type = CV_32SC1 }; };
template <> struct DataType<float> { enum { depth = CV_32F, channels = 1, type = CV_32FC1 }; };
template <> struct DataType<double> { enum { depth = CV_64F, channels = 1, type = CV_64FC1 }; };

/* ---- small value types ---- */

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    /* OpenCV converts between point types with saturate_cast (Point2i p = kp.pt rounds, viso.cpp:1013) */
    template <class U> Point_(const Point_<U>& o) : x(saturate_cast<T>(o.x)), y(saturate_cast<T>(o.y)) {}
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
template <> template <> inline Point_<float>::Point_(const Point_<int>& o) : x((float)o.x), y((float)o.y) {}
template <> template <> inline Point_<float>::Point_(const Point_<double>& o) : x((float)o.x), y((float)o.y) {}
template <> template <> inline Point_<double>::Point_(const Point_<int>& o) : x(o.x), y(o.y) {}
template <> template <> inline Point_<double>::Point_(const Point_<float>& o) : x(o.x), y(o.y) {}
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;
template <class T> inline std::ostream& operator<<(std::ostream& os, const Point_<T>& p) { return os << "[" << p.x << ", " << p.y << "]"; }

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
};
typedef Size_<int> Size;

template <class T, int N> struct Vec {
    enum { channels = N };
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a) { for (int i = 0; i < N; ++i) val[i] = T(); val[0] = a; }
    Vec(T a, T b) { for (int i = 0; i < N; ++i) val[i] = T(); val[0] = a; if (N > 1) val[1 % N] = b; }
    Vec(T a, T b, T c) { static_assert(N >= 3, "Vec"); for (int i = 0; i < N; ++i) val[i] = T(); val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { static_assert(N >= 4, "Vec"); for (int i = 0; i < N; ++i) val[i] = T(); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T& operator()(int i) { return val[i]; }
    const T& operator()(int i) const { return val[i]; }
    bool operator==(const Vec& o) const { for (int i = 0; i < N; ++i) if (!(val[i] == o.val[i])) return false; return true; }
    bool operator!=(const Vec& o) const { return !(*this == o); }
};
typedef Vec<int, 2> Vec2i;
typedef Vec<int, 3> Vec3i;
typedef Vec<int, 4> Vec4i;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<float, 6> Vec6f;
typedef Vec<double, 2> Vec2d;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 4> Vec4d;
typedef Vec<uchar, 3> Vec3b;
template <class T, int N> inline std::ostream& operator<<(std::ostream& os, const Vec<T, N>& v)
{
    os << "[";
    for (int i = 0; i < N; ++i) os << (i ? ", " : "") << v.val[i];
    return os << "]";
}
template <class T, int N> struct DataType<Vec<T, N> > {
    enum { depth = DataType<T>::depth, channels = N, type = CV_MAKETYPE(DataType<T>::depth, N) };
};

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(p), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };
enum { DECOMP_LU = 0, DECOMP_SVD = 1, DECOMP_EIG = 2, DECOMP_CHOLESKY = 3, DECOMP_QR = 4, DECOMP_NORMAL = 16 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4 };

class Mat;
template <class T> class Mat_;
template <class T> class MatCommaInitializer_;

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;

/* ---- Mat ---- */

class Mat {
public:
    enum { AUTO_STEP = 0 };
    int flags;          /* the type (depth + channels) */
    int dims;
    int rows, cols;
    uchar* data;
    size_t step;        /* bytes between rows */

    Mat() : flags(0), dims(0), rows(0), cols(0), data(0), step(0) {}
    Mat(int r, int c, int type) : flags(0), dims(0), rows(0), cols(0), data(0), step(0) { create(r, c, type); }
    Mat(Size s, int type) : flags(0), dims(0), rows(0), cols(0), data(0), step(0) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : flags(0), dims(0), rows(0), cols(0), data(0), step(0) { create(r, c, type); setTo(s); }
    Mat(Size sz, int type, const Scalar& s) : flags(0), dims(0), rows(0), cols(0), data(0), step(0) { create(sz.height, sz.width, type); setTo(s); }
    /* a header over memory the caller owns (no copy, never freed) */
    Mat(int r, int c, int type, void* ext, size_t step_ = AUTO_STEP)
        : flags(type), dims(2), rows(r), cols(c), data((uchar*)ext), step(step_ ? step_ : (size_t)c * esz(type)) {}

    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && flags == type && step == (size_t)c * esz(type)) return;
        flags = type; dims = 2; rows = r; cols = c; step = (size_t)c * esz(type);
        const size_t bytes = step * (size_t)r;
        buf_ = std::shared_ptr<uchar>(new uchar[bytes ? bytes : 1], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf_.reset(); data = 0; rows = cols = 0; dims = 0; step = 0; }

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize() const { return esz(flags); }
    size_t elemSize1() const { return esz(CV_MAT_DEPTH(flags)); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }

    template <class T> T& at(int r, int c) { return ((T*)(data + step * (size_t)r))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + step * (size_t)r))[c]; }
    /* element i of a row or column vector (or of a continuous matrix, row-major) */
    template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    template <class T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    template <class T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <class T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    template <class T> T* ptr(int r = 0) { return (T*)(data + step * (size_t)r); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + step * (size_t)r); }
    uchar* ptr(int r = 0) { return data + step * (size_t)r; }
    const uchar* ptr(int r = 0) const { return data + step * (size_t)r; }

    /* views share the buffer */
    Mat row(int r) const { Mat m(*this); m.rows = 1; m.data = data + step * (size_t)r; return m; }
    Mat col(int c) const { Mat m(*this); m.cols = 1; m.data = data + (size_t)c * elemSize(); return m; }
    Mat rowRange(int r0, int r1) const { Mat m(*this); m.rows = r1 - r0; m.data = data + step * (size_t)r0; return m; }
    Mat colRange(int c0, int c1) const { Mat m(*this); m.cols = c1 - c0; m.data = data + (size_t)c0 * elemSize(); return m; }

    Mat clone() const
    {
        Mat m;
        if (data) {
            m.create(rows, cols, flags);
            const size_t rb = (size_t)cols * elemSize();
            for (int r = 0; r < rows; ++r) std::memcpy(m.data + m.step * (size_t)r, data + step * (size_t)r, rb);
        }
        return m;
    }
    void copyTo(Mat& dst) const
    {
        if (!data) { dst.release(); return; }
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        dst.create(rows, cols, flags);
        const size_t rb = (size_t)cols * elemSize();
        for (int r = 0; r < rows; ++r) std::memcpy(dst.data + dst.step * (size_t)r, data + step * (size_t)r, rb);
    }
    inline void copyTo(const _OutputArray& dst) const;

    double getd(int r, int c) const
    {
        const uchar* p = data + step * (size_t)r;
        switch (depth()) {
        case CV_8U: return p[c];
        case CV_8S: return ((const signed char*)p)[c];
        case CV_16U: return ((const ushort*)p)[c];
        case CV_16S: return ((const short*)p)[c];
        case CV_32S: return ((const int*)p)[c];
        case CV_32F: return ((const float*)p)[c];
        default: return ((const double*)p)[c];
        }
    }
    void setd(int r, int c, double v)
    {
        uchar* p = data + step * (size_t)r;
        switch (depth()) {
        case CV_8U: p[c] = saturate_cast<uchar>(v); break;
        case CV_8S: ((signed char*)p)[c] = (signed char)cvRound(v); break;
        case CV_16U: ((ushort*)p)[c] = saturate_cast<ushort>(v); break;
        case CV_16S: ((short*)p)[c] = saturate_cast<short>(v); break;
        case CV_32S: ((int*)p)[c] = saturate_cast<int>(v); break;
        case CV_32F: ((float*)p)[c] = (float)v; break;
        default: ((double*)p)[c] = v; break;
        }
    }
    Mat& setTo(const Scalar& s)
    {
        const int cn = channels();
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * cn; ++c) setd(r, c, s.val[c % cn]);
        return *this;
    }
    Mat& operator=(const Scalar& s) { return setTo(s); }

    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const
    {
        const int dt = rtype < 0 ? flags : CV_MAKETYPE(rtype, channels());
        Mat out(rows, cols, dt);
        const int cn = channels();
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols * cn; ++c) out.setd(r, c, getd(r, c) * alpha + beta);
        dst = out;
    }

    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0)); }
    static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type, Scalar(0)); }
    static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
    static Mat eye(int r, int c, int type)
    {
        Mat m(r, c, type, Scalar(0));
        for (int i = 0; i < std::min(r, c); ++i) m.setd(i, i, 1.0);
        return m;
    }

    inline Mat t() const;
    inline Mat inv(int method = DECOMP_LU) const;
    inline Mat mul(const Mat& m, double scale = 1) const;
    inline double dot(const Mat& m) const;

private:
    static size_t esz(int type)
    {
        static const size_t d[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return d[CV_MAT_DEPTH(type)] * CV_MAT_CN(type);
    }
    std::shared_ptr<uchar> buf_;
};

/* ---- proxies for function arguments (only plain Mat arguments are supported) ---- */

class _InputArray {
public:
    _InputArray() : m_(0) {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    int rows() const { return m_ ? m_->rows : 0; }
    int cols() const { return m_ ? m_->cols : 0; }
    Size size() const { return Size(cols(), rows()); }
    bool empty() const { return !m_ || m_->empty(); }
    int type() const { return m_ ? m_->type() : 0; }

protected:
    const Mat* m_;
};

class _OutputArray : public _InputArray {
public:
    _OutputArray() : w_(0) {}
    _OutputArray(Mat& m) : _InputArray(m), w_(&m) {}
    Mat& getMatRef() const { return *w_; }
    void create(int r, int c, int type) const { w_->create(r, c, type); }
    void create(Size s, int type) const { w_->create(s, type); }
    void release() const { w_->release(); }
    bool needed() const { return w_ != 0; }

private:
    Mat* w_;
};

inline const _InputArray& noArray()
{
    static _OutputArray none;
    return none;
}

inline void Mat::copyTo(const _OutputArray& dst) const { copyTo(dst.getMatRef()); }

/* ---- Mat_<T> and the comma initialiser: (Mat_<double>(3,4) << a, b, ...) ---- */

template <class T> class Mat_ : public Mat {
public:
    Mat_() : Mat() { flags = DataType<T>::type; }
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(int r, int c, const T& v) : Mat(r, c, DataType<T>::type, Scalar(v)) {}
    Mat_(const Mat& m) : Mat()
    {
        if (m.type() == DataType<T>::type) Mat::operator=(m);
        else m.convertTo(*this, DataType<T>::depth);
    }
    inline Mat_(const MatCommaInitializer_<T>& ci);
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
    T& operator()(int i) { return this->template at<T>(i); }
    const T& operator()(int i) const { return this->template at<T>(i); }
};
typedef Mat_<uchar> Mat1b;
typedef Mat_<int> Mat1i;
typedef Mat_<float> Mat1f;
typedef Mat_<double> Mat1d;

template <class T> class MatCommaInitializer_ {
public:
    explicit MatCommaInitializer_(Mat_<T>* m) : m_(m), i_(0) {}
    template <class U> MatCommaInitializer_& operator,(U v)
    {
        assert(i_ < m_->rows * m_->cols);
        m_->template at<T>(i_ / m_->cols, i_ % m_->cols) = T(v);
        ++i_;
        return *this;
    }
    operator Mat_<T>() const { return *m_; }
    operator Mat() const { return *m_; }
    const Mat_<T>& mat() const { return *m_; }

private:
    Mat_<T>* m_;
    int i_;
};

template <class T> inline Mat_<T>::Mat_(const MatCommaInitializer_<T>& ci) : Mat(ci.mat()) {}

template <class T, class U> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v)
{
    MatCommaInitializer_<T> ci(const_cast<Mat_<T>*>(&m));
    return (ci, v);
}

/* ---- eager operators ---- */

namespace compat_detail {

inline void same_shape(const Mat& a, const Mat& b)
{
    if (a.rows != b.rows || a.cols != b.cols || a.type() != b.type()) throw std::invalid_argument("cv::Mat (compat): operands differ in size or type");
}

template <class F> inline Mat zip(const Mat& a, const Mat& b, F f)
{
    same_shape(a, b);
    Mat r(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < n; ++j) r.setd(i, j, f(a.getd(i, j), b.getd(i, j)));
    return r;
}

template <class F> inline Mat map(const Mat& a, F f)
{
    Mat r(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < n; ++j) r.setd(i, j, f(a.getd(i, j)));
    return r;
}

/* dst = a * b: each element a sequential sum over k in the operands' precision (float for CV_32F, else double) */
inline Mat gemm(const Mat& a, const Mat& b)
{
    if (a.cols != b.rows || a.type() != b.type() || (a.type() != CV_32FC1 && a.type() != CV_64FC1))
        throw std::invalid_argument("cv::Mat (compat): operator* needs CV_32F or CV_64F operands with matching inner size");
    Mat r(a.rows, b.cols, a.type());
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            if (a.type() == CV_32FC1) {
                float s = 0;
                for (int k = 0; k < a.cols; ++k) s += a.at<float>(i, k) * b.at<float>(k, j);
                r.at<float>(i, j) = s;
            } else {
                double s = 0;
                for (int k = 0; k < a.cols; ++k) s += a.at<double>(i, k) * b.at<double>(k, j);
                r.at<double>(i, j) = s;
            }
        }
    return r;
}

/* OpenCV hal::LUImpl (modules/core/src/matrix_decomp.cpp): in-place LU with partial pivoting, eps = 100 * epsilon;
 * returns 0 when singular, else the sign of the permutation; b (m x n) is overwritten with the solution */
template <class T> inline int lu(T* A, size_t astep, int m, T* b, size_t bstep, int n, T eps)
{
    int p = 1;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (std::abs(A[j * astep + i]) > std::abs(A[k * astep + i])) k = j;
        if (std::abs(A[k * astep + i]) < eps) return 0;
        if (k != i) {
            for (int j = i; j < m; j++) std::swap(A[i * astep + j], A[k * astep + j]);
            if (b)
                for (int j = 0; j < n; j++) std::swap(b[i * bstep + j], b[k * bstep + j]);
            p = -p;
        }
        const T d = -1 / A[i * astep + i];
        for (int j = i + 1; j < m; j++) {
            const T alpha = A[j * astep + i] * d;
            for (int k2 = i + 1; k2 < m; k2++) A[j * astep + k2] += alpha * A[i * astep + k2];
            if (b)
                for (int k2 = 0; k2 < n; k2++) b[j * bstep + k2] += alpha * b[i * bstep + k2];
        }
    }
    if (b) {
        for (int i = m - 1; i >= 0; i--)
            for (int j = 0; j < n; j++) {
                T s = b[i * bstep + j];
                for (int k = i + 1; k < m; k++) s -= A[i * astep + k] * b[k * bstep + j];
                b[i * bstep + j] = s / A[i * astep + i];
            }
    }
    return p;
}

} // namespace compat_detail

inline Mat operator+(const Mat& a, const Mat& b) { return compat_detail::zip(a, b, [](double x, double y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return compat_detail::zip(a, b, [](double x, double y) { return x - y; }); }
inline Mat operator-(const Mat& a) { return compat_detail::map(a, [](double x) { return -x; }); }
inline Mat operator*(const Mat& a, const Mat& b) { return compat_detail::gemm(a, b); }
inline Mat operator*(const Mat& a, double s) { return compat_detail::map(a, [s](double x) { return x * s; }); }
inline Mat operator*(double s, const Mat& a) { return a * s; }
inline Mat operator+(const Mat& a, const Scalar& s) { const double v = s.val[0]; return compat_detail::map(a, [v](double x) { return x + v; }); }
inline Mat operator-(const Mat& a, const Scalar& s) { const double v = s.val[0]; return compat_detail::map(a, [v](double x) { return x - v; }); }
/* mat.inl.hpp: operator/(Mat, double) and /= scale by the reciprocal (a.convertTo(a, -1, 1./s)) */
inline Mat operator/(const Mat& a, double s) { const double inv = 1. / s; return compat_detail::map(a, [inv](double x) { return x * inv; }); }
inline Mat& operator+=(Mat& a, const Mat& b) { a = a + b; return a; }
inline Mat& operator-=(Mat& a, const Mat& b) { a = a - b; return a; }
inline Mat& operator*=(Mat& a, double s) { a = a * s; return a; }
inline Mat& operator/=(Mat& a, double s) { a = a / s; return a; }

inline Mat Mat::t() const
{
    Mat r(cols, rows, flags);
    const size_t e = elemSize();
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) std::memcpy(r.data + r.step * (size_t)j + (size_t)i * e, data + step * (size_t)i + (size_t)j * e, e);
    return r;
}

inline Mat Mat::mul(const Mat& m, double scale) const
{
    return compat_detail::zip(*this, m, [scale](double x, double y) { return x * y * scale; });
}

inline double Mat::dot(const Mat& m) const
{
    compat_detail::same_shape(*this, m);
    double s = 0;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols * channels(); ++j) s += getd(i, j) * m.getd(i, j);
    return s;
}

/* cv::determinant: closed forms for n <= 3, LU otherwise (matrix_decomp / lapack.cpp) */
inline double determinant(const Mat& m)
{
    if (m.rows != m.cols || (m.type() != CV_32FC1 && m.type() != CV_64FC1)) throw std::invalid_argument("cv::determinant (compat): square CV_32F / CV_64F only");
    const int n = m.rows;
    if (m.type() == CV_64FC1) {
#define VISO_M(y, x) m.at<double>(y, x)
        if (n == 1) return VISO_M(0, 0);
        if (n == 2) return VISO_M(0, 0) * VISO_M(1, 1) - VISO_M(0, 1) * VISO_M(1, 0);
        if (n == 3)
            return VISO_M(0, 0) * (VISO_M(1, 1) * VISO_M(2, 2) - VISO_M(1, 2) * VISO_M(2, 1)) -
                   VISO_M(0, 1) * (VISO_M(1, 0) * VISO_M(2, 2) - VISO_M(1, 2) * VISO_M(2, 0)) +
                   VISO_M(0, 2) * (VISO_M(1, 0) * VISO_M(2, 1) - VISO_M(1, 1) * VISO_M(2, 0));
#undef VISO_M
        Mat a = m.clone();
        double result = compat_detail::lu<double>(a.ptr<double>(), a.step / sizeof(double), n, 0, 0, 0, DBL_EPSILON * 100);
        if (result != 0)
            for (int i = 0; i < n; ++i) result *= a.at<double>(i, i);
        return result;
    }
#define VISO_M(y, x) ((double)m.at<float>(y, x))
    if (n == 1) return VISO_M(0, 0);
    if (n == 2) return VISO_M(0, 0) * VISO_M(1, 1) - VISO_M(0, 1) * VISO_M(1, 0);
    if (n == 3)
        return VISO_M(0, 0) * (VISO_M(1, 1) * VISO_M(2, 2) - VISO_M(1, 2) * VISO_M(2, 1)) -
               VISO_M(0, 1) * (VISO_M(1, 0) * VISO_M(2, 2) - VISO_M(1, 2) * VISO_M(2, 0)) +
               VISO_M(0, 2) * (VISO_M(1, 0) * VISO_M(2, 1) - VISO_M(1, 1) * VISO_M(2, 0));
#undef VISO_M
    Mat a = m.clone();
    double result = compat_detail::lu<float>(a.ptr<float>(), a.step / sizeof(float), n, 0, 0, 0, FLT_EPSILON * 10);
    if (result != 0)
        for (int i = 0; i < n; ++i) result *= a.at<float>(i, i);
    return result;
}

/* cv::invert(DECOMP_LU): dst = I, LU of a copy of src with dst as the right-hand side; zero matrix when singular.
 * (OpenCV uses closed forms for n <= 3; the pose bookkeeping only inverts 4 x 4, viso.cpp:1319.) */
inline double invert(const Mat& src, Mat& dst, int method = DECOMP_LU)
{
    if (method != DECOMP_LU) throw std::invalid_argument("cv::invert (compat): DECOMP_LU only");
    if (src.rows != src.cols || (src.type() != CV_32FC1 && src.type() != CV_64FC1)) throw std::invalid_argument("cv::invert (compat): square CV_32F / CV_64F only");
    const int n = src.rows;
    Mat a = src.clone();
    Mat out = Mat::eye(n, n, src.type());
    int ok;
    if (src.type() == CV_64FC1) ok = compat_detail::lu<double>(a.ptr<double>(), a.step / sizeof(double), n, out.ptr<double>(), out.step / sizeof(double), n, DBL_EPSILON * 100);
    else ok = compat_detail::lu<float>(a.ptr<float>(), a.step / sizeof(float), n, out.ptr<float>(), out.step / sizeof(float), n, FLT_EPSILON * 10);
    if (!ok) out.setTo(Scalar(0));
    dst = out;
    return ok != 0;
}

inline Mat Mat::inv(int method) const
{
    Mat r;
    invert(*this, r, method);
    return r;
}

inline double norm(const Mat& m, int type = NORM_L2)
{
    double s = 0;
    const int n = m.cols * m.channels();
    for (int i = 0; i < m.rows; ++i)
        for (int j = 0; j < n; ++j) {
            const double v = m.getd(i, j);
            if (type == NORM_L1) s += std::abs(v);
            else if (type == NORM_INF) s = std::max(s, std::abs(v));
            else s += v * v;
        }
    return type == NORM_L2 ? std::sqrt(s) : s;
}
inline double norm(const Mat& a, const Mat& b, int type = NORM_L2) { return norm(a - b, type); }

inline void transpose(const Mat& src, Mat& dst) { dst = src.t(); }
inline void hconcat(const Mat& a, const Mat& b, Mat& dst)
{
    Mat r(a.rows, a.cols + b.cols, a.type());
    for (int i = 0; i < a.rows; ++i) {
        for (int j = 0; j < a.cols * a.channels(); ++j) r.setd(i, j, a.getd(i, j));
        for (int j = 0; j < b.cols * b.channels(); ++j) r.setd(i, a.cols * a.channels() + j, b.getd(i, j));
    }
    dst = r;
}
inline void vconcat(const Mat& a, const Mat& b, Mat& dst)
{
    Mat r(a.rows + b.rows, a.cols, a.type());
    for (int j = 0; j < a.cols * a.channels(); ++j) {
        for (int i = 0; i < a.rows; ++i) r.setd(i, j, a.getd(i, j));
        for (int i = 0; i < b.rows; ++i) r.setd(a.rows + i, j, b.getd(i, j));
    }
    dst = r;
}

inline std::ostream& operator<<(std::ostream& os, const Mat& m)
{
    os << "[";
    const int n = m.cols * m.channels();
    for (int i = 0; i < m.rows; ++i) {
        for (int j = 0; j < n; ++j) os << (j ? ", " : "") << m.getd(i, j);
        os << (i + 1 < m.rows ? ";\n " : "");
    }
    return os << "]";
}

/* cv::format: printf into a string */
inline String format(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (n < (int)sizeof(buf)) return String(buf, n > 0 ? n : 0);
    std::vector<char> big(n + 1);
    va_start(ap, fmt);
    vsnprintf(big.data(), big.size(), fmt, ap);
    va_end(ap);
    return String(big.data(), n);
}

/* declared so `using cv::FileStorage` (src/viso.h:39) resolves; persistence is not part of the path */
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const String&, int) {}
    bool isOpened() const { return false; }
    void release() {}
};

/* cv::RNG (Multiply-With-Carry, modules/core/include/opencv2/core/operations.hpp) */
class RNG {
public:
    uint64_t state;
    RNG() : state(0xffffffff) {}
    explicit RNG(uint64_t s) : state(s ? s : 0xffffffff) {}
    unsigned next() { state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32); return (unsigned)state; }
    operator unsigned() { return next(); }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
    float uniform(float a, float b) { return ((float)next() * 2.3283064365386962890625e-10f) * (b - a) + a; }
    double uniform(double a, double b)
    {
        unsigned t = next();
        const double d = ((uint64_t)t << 32 | next()) * 5.4210108624275221700372640043497e-20;
        return d * (b - a) + a;
    }
};

} // namespace cv

#endif
