/* compat/opencv2/video/tracking.hpp -- included by src/viso.cpp:7; nothing of it is used */
#ifndef VISO_COMPAT_OPENCV2_VIDEO_TRACKING_HPP_
#define VISO_COMPAT_OPENCV2_VIDEO_TRACKING_HPP_
#include "../core/core.hpp"
#endif
