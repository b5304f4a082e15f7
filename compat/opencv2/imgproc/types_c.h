/* compat/opencv2/imgproc/types_c.h -- colour-conversion codes named by the reference's debug drawing code */
#ifndef VISO_COMPAT_OPENCV2_IMGPROC_TYPES_C_H_
#define VISO_COMPAT_OPENCV2_IMGPROC_TYPES_C_H_
enum { CV_BGR2GRAY = 6, CV_RGB2GRAY = 7, CV_GRAY2BGR = 8, CV_GRAY2RGB = 8 };
#endif
