/* compat/opencv2/calib3d/calib3d.hpp -- see core/core.hpp.  Only included by the reference's headers; the mono
 * pipeline that calls findEssentialMat (viso.cpp:1332-1398) is out of scope. */
#ifndef VISO_COMPAT_OPENCV2_CALIB3D_HPP_
#define VISO_COMPAT_OPENCV2_CALIB3D_HPP_
#include "../core/core.hpp"
namespace cv {
enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8 };
}
#endif
