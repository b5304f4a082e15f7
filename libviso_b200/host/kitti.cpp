/*
 * kitti.cpp (B200) -- the reference's KITTI driver (reference src/kitti.cpp:80-111) on top of the device pipeline:
 *
 *     KITTI_HOME=/data/kitti  kitti <result_sha> <seq_name> [begin [end]]
 *
 * reads $KITTI_HOME/sequences/<seq>/calib.txt (loadCalib, kitti.cpp:23-46), the stereo pairs image_0 / image_1
 * "%06d" from index `begin` until the first missing file or `end` (StereoImageGenerator, viso.h:81-101), runs
 * sequence_odometry (detector, extractor, matching, circle closure, triangulation, RANSAC / Gauss-Newton all on the
 * GPU) and writes $KITTI_HOME/results/<seq>/<result_sha>/data/<seq>.txt (savePoses, kitti.cpp:49-64).
 *
 * Image decoding is the one thing that stays outside the library: with OpenCV the frames are read with cv::imread
 * exactly as the reference does (.png); without it (this image has no OpenCV headers) binary PGM (P5, maxval 255)
 * files with the same names and the extension .pgm are read.
 *
 *     g++ -std=c++17 -O2 libviso_b200/host/kitti.cpp libviso_b200/host/viso.cpp -Llibviso_b200 -lviso_b200 -o kitti
 */
#include "viso.h"
#include "kitti_io.h"

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <sys/stat.h>

#ifdef VISO_B200_HAVE_OPENCV
#include <opencv2/highgui/highgui.hpp>
#endif

namespace {

std::string index_name(const std::string& mask, int index)
{
    char buf[4096];
    std::snprintf(buf, sizeof(buf), mask.c_str(), index);
    return buf;
}

/* binary PGM (P5), 8 bit; comments after the magic number are skipped */
bool read_pgm(const std::string& name, Mat& out)
{
    FILE* fp = std::fopen(name.c_str(), "rb");
    if (!fp) return false;
    auto token = [&](int& v) {
        int c = std::fgetc(fp);
        while (c == '#' || c == ' ' || c == '\n' || c == '\r' || c == '\t') {
            if (c == '#') while (c != '\n' && c != EOF) c = std::fgetc(fp);
            c = std::fgetc(fp);
        }
        v = 0;
        bool any = false;
        while (c >= '0' && c <= '9') { v = v * 10 + (c - '0'); any = true; c = std::fgetc(fp); }
        return any; /* the single whitespace after the token has been consumed */
    };
    int w = 0, h = 0, maxv = 0;
    bool ok = std::fgetc(fp) == 'P' && std::fgetc(fp) == '5' && token(w) && token(h) && token(maxv) && w > 0 && h > 0 && maxv == 255;
    if (ok) {
        out.create(h, w, CV_8U);
        ok = std::fread(out.ptr<unsigned char>(0), 1, (size_t)w * h, fp) == (size_t)w * h;
    }
    std::fclose(fp);
    return ok;
}

/* StereoImageGenerator (viso.h:81-101): the sequence ends at m_end or at the first pair that cannot be read */
class FileStereoSource : public StereoImageSource {
public:
    FileStereoSource(const std::string& mask0, const std::string& mask1, int begin, int end)
        : m_mask0(mask0), m_mask1(mask1), m_index(begin), m_end(end) {}
    bool next(image_pair& out) override
    {
        if (m_index > m_end) return false;
        const std::string n0 = index_name(m_mask0, m_index), n1 = index_name(m_mask1, m_index);
        ++m_index;
#ifdef VISO_B200_HAVE_OPENCV
        out = image_pair(cv::imread(n0, 0), cv::imread(n1, 0));
        return out.first.data && out.second.data;
#else
        return read_pgm(n0, out.first) && read_pgm(n1, out.second);
#endif
    }

private:
    std::string m_mask0, m_mask1;
    int m_index, m_end;
};

void make_dirs(const std::string& path)
{
    for (size_t i = 1; i <= path.size(); ++i)
        if (i == path.size() || path[i] == '/') mkdir(path.substr(0, i).c_str(), 0777);
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 3) {
        std::printf("usage: %s result_sha seq_name begin end\n", argv[0]); /* kitti.cpp:84 */
        return 1;
    }
    int begin = 0, end = INT_MAX;
    if (argc > 3) begin = std::atoi(argv[3]);
    if (argc > 4) end = std::atoi(argv[4]);
    const char* home = std::getenv("KITTI_HOME");
    if (!home) { std::fprintf(stderr, "KITTI_HOME is not set\n"); return 1; } /* assert(KITTI_HOME), kitti.cpp:98 */
    if (const char* nf = std::getenv("VISO_MAX_FEATURES")) viso_b200::set_max_features(std::atoi(nf));
    const std::string result_sha = argv[1], seq_name = argv[2];
    const std::string seq_dir = std::string(home) + "/sequences/" + seq_name;
    const std::string result_dir = std::string(home) + "/results/" + seq_name + "/" + result_sha;
    Mat P1(3, 4, CV_64F), P2(3, 4, CV_64F);
    if (!loadCalib(seq_dir + "/calib.txt", P1, P2)) { std::fprintf(stderr, "cannot read %s/calib.txt\n", seq_dir.c_str()); return 1; }
#ifdef VISO_B200_HAVE_OPENCV
    const char* ext = "png";
#else
    const char* ext = "pgm";
#endif
    FileStereoSource images(seq_dir + "/image_0/%06d." + ext, seq_dir + "/image_1/%06d." + ext, begin, end);
    vector<Mat> poses;
    try {
        poses = sequence_odometry(P1, P2, images);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "sequence_odometry failed: %s\n", e.what());
        return 2;
    }
    make_dirs(result_dir + "/data");
    const std::string poses_file = result_dir + "/data/" + seq_name + ".txt";
    if (!savePoses(poses_file, poses)) { std::fprintf(stderr, "cannot write %s\n", poses_file.c_str()); return 1; }
    std::printf("%zu poses -> %s (%lld kernel launches)\n", poses.size(), poses_file.c_str(), viso_b200::kernel_launches());
    return 0;
}
