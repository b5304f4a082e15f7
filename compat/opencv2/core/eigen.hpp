/* compat/opencv2/core/eigen.hpp -- included by src/viso.cpp:8 (cv2eigen is only called from dead helpers) */
#ifndef VISO_COMPAT_OPENCV2_CORE_EIGEN_HPP_
#define VISO_COMPAT_OPENCV2_CORE_EIGEN_HPP_
#include "core.hpp"
#endif
