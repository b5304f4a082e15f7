/*
 * misc.h (B200) -- the reference's src/misc.h restated: the small cv::Mat helpers its headers and callers use (isEqual,
 * _str, vcat / hcat, e2h / h2e, measure).  Host-side utilities with the reference's names, signatures and error
 * behaviour (h2e throws std::overflow_error on a zero last coordinate, misc.h:118-119); nothing here is on the per-frame path.
 */
#ifndef VISO_B200_HOST_MISC_H_
#define VISO_B200_HOST_MISC_H_

#include <chrono>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <opencv2/core/core.hpp>

using std::string;
using cv::Mat;
using std::stringstream;

bool isEqual(double x, double y); /* |x - y| <= 1e-6 |x|, misc.cpp:3-8 */
bool isEqual(float x, float y);

/* "(rows x cols) [a,b;c,d]", cut after `truncate` elements (misc.h:17-47) */
template <typename T> string _str(const Mat& m, bool include_dims = true, int truncate = 16, int precision = 2)
{
    (void)precision;
    stringstream ss;
    if (include_dims) ss << "(" << m.rows << "x" << m.cols << ") ";
    ss << "[";
    int k = 0;
    for (int i = 0; i < m.rows; ++i) {
        if (i > 0) ss << " ";
        for (int j = 0; j < m.cols; ++j, ++k) {
            ss << m.at<T>(i, j);
            if (j < m.cols - 1) ss << ",";
            if (k == truncate) { ss << "...]"; return ss.str(); }
        }
        if (i < m.rows - 1) ss << ";";
    }
    ss << "]";
    return ss.str();
}

/* m1 on top of m2 (misc.h:49-71) */
template <class T> Mat vcat(const Mat& m1, const Mat& m2)
{
    Mat res(m1.rows + m2.rows, m1.cols, m1.type());
    for (int i = 0; i < m1.rows; ++i)
        for (int j = 0; j < m1.cols; ++j) res.at<T>(i, j) = m1.at<T>(i, j);
    for (int i = 0; i < m2.rows; ++i)
        for (int j = 0; j < m2.cols; ++j) res.at<T>(m1.rows + i, j) = m2.at<T>(i, j);
    return res;
}

/* m1 left of m2 (misc.h:73-88) */
template <class T> Mat hcat(const Mat& m1, const Mat& m2)
{
    Mat m(m1.rows, m1.cols + m2.cols, m1.type());
    for (int i = 0; i < m1.rows; ++i)
        for (int j = 0; j < m1.cols + m2.cols; ++j) m.at<T>(i, j) = (j < m1.cols) ? m1.at<T>(i, j) : m2.at<T>(i, j - m1.cols);
    return m;
}

/* euclidean -> homogeneous: a row of ones is appended (misc.h:90-107) */
template <class T> Mat e2h(const Mat& X)
{
    Mat Xh(X.rows + 1, X.cols, cv::DataType<T>::type);
    for (int i = 0; i < X.rows; ++i)
        for (int j = 0; j < X.cols; ++j) Xh.at<T>(i, j) = X.at<T>(i, j);
    for (int j = 0; j < X.cols; ++j) Xh.at<T>(Xh.rows - 1, j) = 1.0;
    return Xh;
}

/* homogeneous -> euclidean: division by the last row (misc.h:109-124) */
template <class T> Mat h2e(const Mat& X)
{
    Mat Xe(X.rows - 1, X.cols, X.type());
    for (int i = 0; i < Xe.rows; ++i)
        for (int j = 0; j < Xe.cols; ++j) {
            if (isEqual(std::abs(X.at<T>(X.rows - 1, j)), .0f)) throw std::overflow_error("divide by zero in h2e");
            Xe.at<T>(i, j) = X.at<T>(i, j) / X.at<T>(X.rows - 1, j);
        }
    return Xe;
}

/* measure<>::execution(f): wall time of f() in TimeT units (misc.h:126-139) */
template <typename TimeT = std::chrono::milliseconds> struct measure {
    template <typename F> static typename TimeT::rep execution(F const& func)
    {
        const auto start = std::chrono::system_clock::now();
        func();
        return std::chrono::duration_cast<TimeT>(std::chrono::system_clock::now() - start).count();
    }
};

#endif /* VISO_B200_HOST_MISC_H_ */
