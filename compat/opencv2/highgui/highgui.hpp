/*
 * compat/opencv2/highgui/highgui.hpp -- cv::imread for 8-bit grayscale images, which is all StereoImageGenerator needs
 * (src/viso.h:92: cv::imread(name, CV_LOAD_IMAGE_GRAYSCALE) of KITTI's image_0 / image_1 PNGs).  Decodes binary PGM
 * (P5) natively and non-interlaced 8-bit gray PNG through zlib (link with -lz; libpng is not in this image).  Anything
 * else yields an empty Mat, which ends the sequence exactly like an unreadable file does in the reference
 * (viso.h:95).  imwrite / imshow / waitKey exist for the debug code paths and do nothing.
 */
#ifndef VISO_COMPAT_OPENCV2_HIGHGUI_HPP_
#define VISO_COMPAT_OPENCV2_HIGHGUI_HPP_

#include "../core/core.hpp"
#include "highgui_c.h"

#include <cstdint>
#include <cstdlib>
#include <zlib.h>

namespace cv {

enum { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum { WINDOW_AUTOSIZE = 1 };

namespace compat_detail {

inline bool read_file(const String& name, std::vector<uchar>& out)
{
    FILE* fp = fopen(name.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    const long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (n <= 0) { fclose(fp); return false; }
    out.resize((size_t)n);
    const size_t got = fread(out.data(), 1, (size_t)n, fp);
    fclose(fp);
    return got == (size_t)n;
}

inline Mat decode_pgm(const std::vector<uchar>& f)
{
    size_t p = 2;
    int vals[3], nv = 0;
    while (nv < 3 && p < f.size()) {
        while (p < f.size() && (f[p] == ' ' || f[p] == '\n' || f[p] == '\r' || f[p] == '\t')) ++p;
        if (p < f.size() && f[p] == '#') { while (p < f.size() && f[p] != '\n') ++p; continue; }
        int v = 0, digits = 0;
        while (p < f.size() && f[p] >= '0' && f[p] <= '9') { v = v * 10 + (f[p] - '0'); ++p; ++digits; }
        if (!digits) return Mat();
        vals[nv++] = v;
    }
    ++p; /* the single whitespace after maxval */
    if (nv < 3 || vals[2] != 255 || vals[0] <= 0 || vals[1] <= 0 || p + (size_t)vals[0] * vals[1] > f.size()) return Mat();
    Mat m(vals[1], vals[0], CV_8UC1);
    std::memcpy(m.data, f.data() + p, (size_t)vals[0] * vals[1]);
    return m;
}

inline uint32_t be32(const uchar* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline Mat decode_png_gray8(const std::vector<uchar>& f)
{
    static const uchar sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (f.size() < 33 || std::memcmp(f.data(), sig, 8) != 0) return Mat();
    size_t p = 8;
    int w = 0, h = 0, ctype = -1, bpp = 0;
    std::vector<uchar> idat;
    while (p + 12 <= f.size()) {
        const uint32_t len = be32(&f[p]);
        const uchar* tag = &f[p + 4];
        if (p + 12 + len > f.size()) return Mat();
        const uchar* body = &f[p + 8];
        if (!std::memcmp(tag, "IHDR", 4)) {
            if (len < 13) return Mat();
            w = (int)be32(body); h = (int)be32(body + 4);
            const int depth = body[8];
            ctype = body[9];
            if (depth != 8 || body[12] != 0 /* interlaced */) return Mat();
            bpp = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
            if (!bpp || w <= 0 || h <= 0) return Mat();
        } else if (!std::memcmp(tag, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(tag, "IEND", 4)) {
            break;
        }
        p += 12 + (size_t)len;
    }
    if (ctype < 0 || idat.empty()) return Mat();
    const size_t stride = (size_t)w * bpp;
    std::vector<uchar> raw((stride + 1) * (size_t)h);
    uLongf rawlen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return Mat();
    std::vector<uchar> prev(stride, 0), cur(stride);
    Mat m(h, w, CV_8UC1);
    for (int y = 0; y < h; ++y) {
        const uchar* src = &raw[(stride + 1) * (size_t)y];
        const int ft = src[0];
        ++src;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)bpp ? cur[x - bpp] : 0, b = prev[x], c = x >= (size_t)bpp ? prev[x - bpp] : 0;
            int pred = 0;
            if (ft == 1) pred = a;
            else if (ft == 2) pred = b;
            else if (ft == 3) pred = (a + b) >> 1;
            else if (ft == 4) {
                const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            } else if (ft != 0) return Mat();
            cur[x] = (uchar)(src[x] + pred);
        }
        uchar* dst = m.ptr<uchar>(y);
        if (bpp <= 2) {
            for (int x = 0; x < w; ++x) dst[x] = cur[(size_t)x * bpp];
        } else { /* RGB(A) -> gray, OpenCV's fixed-point weights (R 4899, G 9617, B 1868) / 16384 */
            for (int x = 0; x < w; ++x) {
                const uchar* px = &cur[(size_t)x * bpp];
                dst[x] = (uchar)((px[0] * 4899 + px[1] * 9617 + px[2] * 1868 + 8192) >> 14);
            }
        }
        prev.swap(cur);
    }
    return m;
}

} // namespace compat_detail

inline Mat imread(const String& name, int flags = IMREAD_COLOR)
{
    (void)flags; /* every supported file decodes to one 8-bit channel */
    std::vector<uchar> f;
    if (!compat_detail::read_file(name, f) || f.size() < 8) return Mat();
    if (f[0] == 'P' && f[1] == '5') return compat_detail::decode_pgm(f);
    return compat_detail::decode_png_gray8(f);
}

/* binary PGM writer (8-bit, one channel); other inputs are ignored */
inline bool imwrite(const String& name, const Mat& m, const std::vector<int>& = std::vector<int>())
{
    if (m.empty() || m.type() != CV_8UC1) return false;
    const size_t n = name.size();
    if (n < 4 || name.compare(n - 4, 4, ".pgm") != 0) return false;
    FILE* fp = fopen(name.c_str(), "wb");
    if (!fp) return false;
    fprintf(fp, "P5\n%d %d\n255\n", m.cols, m.rows);
    for (int r = 0; r < m.rows; ++r) fwrite(m.ptr<uchar>(r), 1, (size_t)m.cols, fp);
    fclose(fp);
    return true;
}

inline void namedWindow(const String&, int = WINDOW_AUTOSIZE) {}
inline void imshow(const String&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }

} // namespace cv
#endif
