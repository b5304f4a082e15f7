/*
 * cvcompat.h -- the few OpenCV core types the libviso hot-path API is written in (cv::Mat of CV_64F / CV_32F /
 * CV_32S / CV_8U, KeyPoint, Point2f, Vec3i / Vec4i), for builds where OpenCV itself is not installed (this image has no
 * OpenCV C++ headers: SURVEY.md 8c).  When <opencv2/core/core.hpp> exists it is used instead and this file adds
 * nothing.  Only what libviso_b200/host/ and its tests touch is provided; layouts follow OpenCV (row-major,
 * ref-counted buffer, at<T>(r,c), ptr<T>(r)).
 */
#ifndef VISO_B200_CVCOMPAT_H_
#define VISO_B200_CVCOMPAT_H_

#if defined(__has_include)
#if __has_include(<opencv2/core/core.hpp>) && !defined(VISO_B200_FORCE_COMPAT)
#define VISO_B200_HAVE_OPENCV 1
#endif
#endif

#ifdef VISO_B200_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else

#include <cassert>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6

namespace cv {

template <class T> struct DataType;
template <> struct DataType<unsigned char> { enum { type = CV_8U }; };
template <> struct DataType<int> { enum { type = CV_32S }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;

struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(float x, float y, float size_ = 1.f) : pt(x, y), size(size_) {}
};

template <class T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    Vec(T a, T b, T c) { static_assert(N == 3, "Vec3"); val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b, T c, T d) { static_assert(N == 4, "Vec4"); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    bool operator==(const Vec& o) const { for (int i = 0; i < N; ++i) if (!(val[i] == o.val[i])) return false; return true; }
};
typedef Vec<int, 3> Vec3i;
typedef Vec<int, 4> Vec4i;

class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, double fill) { create(r, c, type); setTo(fill); }

    void create(int r, int c, int type)
    {
        assert(type == CV_8U || type == CV_32S || type == CV_32F || type == CV_64F);
        rows = r; cols = c; type_ = type;
        const size_t bytes = (size_t)r * c * elemSize();
        buf_ = std::shared_ptr<unsigned char>(new unsigned char[bytes ? bytes : 1], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, 0.0); }
    static Mat eye(int r, int c, int type)
    {
        Mat m(r, c, type, 0.0);
        for (int i = 0; i < (r < c ? r : c); ++i) m.set(i, i, 1.0);
        return m;
    }
    int type() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    size_t elemSize() const { return type_ == CV_64F ? 8 : type_ == CV_8U ? 1 : 4; }
    size_t total() const { return (size_t)rows * cols; }
    template <class T> T& at(int r, int c) { assert(DataType<T>::type == type_); return reinterpret_cast<T*>(data)[(size_t)r * cols + c]; }
    template <class T> const T& at(int r, int c) const { assert(DataType<T>::type == type_); return reinterpret_cast<const T*>(data)[(size_t)r * cols + c]; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data) + (size_t)r * cols; }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data) + (size_t)r * cols; }
    Mat clone() const
    {
        Mat m;
        if (type_ >= 0) { m.create(rows, cols, type_); std::memcpy(m.data, data, total() * elemSize()); }
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    void setTo(double v)
    {
        for (size_t i = 0; i < total(); ++i) {
            if (type_ == CV_64F) reinterpret_cast<double*>(data)[i] = v;
            else if (type_ == CV_32F) reinterpret_cast<float*>(data)[i] = (float)v;
            else if (type_ == CV_8U) data[i] = (unsigned char)v;
            else reinterpret_cast<int*>(data)[i] = (int)v;
        }
    }

private:
    void set(int r, int c, double v)
    {
        if (type_ == CV_64F) at<double>(r, c) = v;
        else if (type_ == CV_32F) at<float>(r, c) = (float)v;
        else if (type_ == CV_8U) at<unsigned char>(r, c) = (unsigned char)v;
        else at<int>(r, c) = (int)v;
    }
    int type_ = -1;
    std::shared_ptr<unsigned char> buf_;
};

} // namespace cv
#endif /* !VISO_B200_HAVE_OPENCV */
#endif
