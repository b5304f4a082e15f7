/* compat/opencv2/imgproc/imgproc.hpp -- see core/core.hpp.  The reference's headers only include this file; none of
 * the image-processing routines is part of the C++ API surface (the device front end replaces cv::cornerHarris and
 * cv::Sobel: libviso_b200/csrc/detect.cu, match.cu). */
#ifndef VISO_COMPAT_OPENCV2_IMGPROC_HPP_
#define VISO_COMPAT_OPENCV2_IMGPROC_HPP_
#include "../core/core.hpp"
#include "types_c.h"
#endif
