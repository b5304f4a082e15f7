/*
 * mvg.h (B200) -- the hot-path functions of the reference's src/mvg.h / src/mvg.cpp with the same signatures, bodies
 * over the C-ABI (include/viso_b200.h):
 *   triangulate_dlt        mvg.cpp:124-169 (decl mvg.h:77-78)
 *   triangulate_rectified  mvg.cpp:172-192 (decl mvg.h:80-86; the float version -- the pipeline's double version is
 *                          the template in viso.h)
 *   F_from_P               mvg.h:41-66 (T = double; declared in viso.h)
 */
#ifndef VISO_B200_HOST_MVG_H_
#define VISO_B200_HOST_MVG_H_

#include "cvcompat.h"

cv::Mat triangulate_dlt(const cv::Mat& x1, const cv::Mat& x2, const cv::Mat& P1, const cv::Mat& P2);
cv::Mat triangulate_rectified(const cv::Mat& x1, const cv::Mat& x2, double f, double base, double c1u, double c1v);

#endif
