/*
 * estimation.h (B200) -- solveRigidMotion with the reference's signature (reference src/estimation.h:7-9,
 * src/estimation.cpp:29-51: Kabsch / orthogonal Procrustes), computed on the device (viso_solve_rigid_motion):
 * A, B are 3 x n, T maps B onto A (T * B ~ A: the direction the reference's code implements).
 * The cv::Mat overload (3 x n CV_32F in, 4 x 4 CV_32F out) is an addition for callers without Eigen types.
 */
#ifndef VISO_B200_HOST_ESTIMATION_H_
#define VISO_B200_HOST_ESTIMATION_H_

#include <Eigen/Dense>
#include "viso.h"

void solveRigidMotion(const Eigen::MatrixXf& A, const Eigen::MatrixXf& B, Eigen::Affine3f& T);
void solveRigidMotion(const cv::Mat& A, const cv::Mat& B, cv::Mat& T);

#endif /* VISO_B200_HOST_ESTIMATION_H_ */
