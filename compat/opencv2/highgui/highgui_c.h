/* compat/opencv2/highgui/highgui_c.h -- the C-API constants src/viso.h uses (viso.h:92, :111) */
#ifndef VISO_COMPAT_OPENCV2_HIGHGUI_C_H_
#define VISO_COMPAT_OPENCV2_HIGHGUI_C_H_
enum { CV_LOAD_IMAGE_UNCHANGED = -1, CV_LOAD_IMAGE_GRAYSCALE = 0, CV_LOAD_IMAGE_COLOR = 1 };
enum { CV_WINDOW_AUTOSIZE = 1 };
#endif
