/* compat/opencv2/nonfree/nonfree.hpp -- included by src/viso.h:11; nothing of it is used */
#ifndef VISO_COMPAT_OPENCV2_NONFREE_HPP_
#define VISO_COMPAT_OPENCV2_NONFREE_HPP_
#include "../core/core.hpp"
#endif
