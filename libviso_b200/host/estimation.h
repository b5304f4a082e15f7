/*
 * estimation.h (B200) -- solveRigidMotion, reference src/estimation.h:7-9 / src/estimation.cpp:29-51 (Kabsch), over
 * the C-ABI.  The reference's signature uses Eigen (MatrixXf, Affine3f); with Eigen available the overload below is
 * compiled with those types, otherwise (this image has no Eigen) the cv::Mat form is the only one:
 * A, B: 3 x n CV_32F, T: 4 x 4 CV_32F with T * B ~ A (the direction the reference's code implements).
 */
#ifndef VISO_B200_HOST_ESTIMATION_H_
#define VISO_B200_HOST_ESTIMATION_H_

#include "cvcompat.h"

void solveRigidMotion(const cv::Mat& A, const cv::Mat& B, cv::Mat& T);

#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
inline void solveRigidMotion(const Eigen::MatrixXf& A, const Eigen::MatrixXf& B, Eigen::Affine3f& T)
{
    cv::Mat a(3, (int)A.cols(), CV_32F), b(3, (int)B.cols(), CV_32F), t;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < (int)A.cols(); ++c) { a.at<float>(r, c) = A(r, c); b.at<float>(r, c) = B(r, c); }
    solveRigidMotion(a, b, t);
    Eigen::Matrix4f m;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m(r, c) = t.at<float>(r, c);
    T = Eigen::Affine3f(m);
}
#endif
#endif

#endif
