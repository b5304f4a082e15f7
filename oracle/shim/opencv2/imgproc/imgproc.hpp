/*
 * oracle/shim/opencv2/imgproc/imgproc.hpp -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * cv::Sobel and cv::cornerHarris as the reference's front-end classes call them (viso.cpp:930, :1010), so that
 * HarrisBinnedFeatureDetector and MyFeatureExtractor compile unchanged into oracle/_ref/libviso_ref.so.
 *   Sobel(8-bit -> CV_32F, dx = 1, dy = 0, ksize 3, scale 1, delta 0, BORDER_REFLECT_101): integer arithmetic, so every
 *     correct implementation agrees; pinned against OpenCV 4.13 (tests/golden/sobel.npz).
 *   cornerHarris(block 3, aperture 5, k): OpenCV's float32 rounding is not reproducible across builds (DESIGN.md
 *     section 2), so this forwards to the oracle's canonical float32 evaluation (vo_harris_response): the parity that
 *     _ref adds for the detector is the reference's own binning / selection code (viso.cpp:933-976) running unchanged.
 *     The `k` ARGUMENT IS IGNORED: the reference passes m_k, which its constructor never initialises (viso.cpp:915-919
 *     vs :978) -- whatever the stack held.  cv::shim::harris_k() (default 0.04, the constructor's default argument, the
 *     value every other implementation here uses) is taken instead.
 */
#ifndef VISO_ORACLE_SHIM_OPENCV2_IMGPROC_IMGPROC_HPP_
#define VISO_ORACLE_SHIM_OPENCV2_IMGPROC_IMGPROC_HPP_

#include_next <opencv2/imgproc/imgproc.hpp>

extern "C" void vo_harris_response(const unsigned char* img, int h, int w, float k, float* resp); /* oracle/viso_oracle.h */

namespace cv {

namespace shim {
inline float& harris_k() { static float k = 0.04f; return k; }
}

inline void Sobel(InputArray src_, OutputArray dst_, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0,
                  int borderType = BORDER_DEFAULT)
{
    const Mat src = src_.getMat();
    if (src.type() != CV_8UC1 || CV_MAT_DEPTH(ddepth) != CV_32F || dx != 1 || dy != 0 || ksize != 3 || scale != 1 || delta != 0 ||
        borderType != BORDER_REFLECT_101)
        throw std::invalid_argument("cv::Sobel (shim): only the call of viso.cpp:1010 is supported");
    const int h = src.rows, w = src.cols;
    Mat out(h, w, CV_32FC1);
    auto refl = [](int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * n - 2 - i; return std::min(std::max(i, 0), n - 1); };
    for (int y = 0; y < h; ++y) {
        const uchar *r0 = src.ptr<uchar>(refl(y - 1, h)), *r1 = src.ptr<uchar>(y), *r2 = src.ptr<uchar>(refl(y + 1, h));
        float* o = out.ptr<float>(y);
        for (int x = 0; x < w; ++x) {
            const int xl = refl(x - 1, w), xr = refl(x + 1, w);
            o[x] = (float)((r0[xr] + 2 * r1[xr] + r2[xr]) - (r0[xl] + 2 * r1[xl] + r2[xl]));
        }
    }
    dst_.getMatRef() = out;
}

inline void cornerHarris(InputArray src_, OutputArray dst_, int blockSize, int ksize, double k, int borderType = BORDER_DEFAULT)
{
    const Mat src = src_.getMat();
    if (src.type() != CV_8UC1 || blockSize != 3 || ksize != 5 || borderType != BORDER_DEFAULT || !src.isContinuous())
        throw std::invalid_argument("cv::cornerHarris (shim): only the call of viso.cpp:930 is supported");
    Mat out(src.rows, src.cols, CV_32FC1);
    (void)k; /* uninitialised in the reference: see the header comment */
    vo_harris_response(src.data, src.rows, src.cols, shim::harris_k(), out.ptr<float>());
    dst_.getMatRef() = out;
}

} // namespace cv
#endif
