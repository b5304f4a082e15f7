/*
 * viso.h (B200) -- the reference's hot-path API (reference src/viso.h, plus the header-less hot-path functions of
 * src/viso.cpp that have external linkage) with the same names, argument meaning and error behaviour; the bodies in
 * viso.cpp call the C-ABI of include/viso_b200.h and nothing else.  A caller written against the reference
 * (test/test.cpp:152-168, the per-frame loop of sequence_odometry) compiles against this header unchanged for the
 * functions listed here.  There is no CPU fallback: without a CUDA device every call throws viso_b200_error.
 *
 *   reference                                              here
 *   struct param                       viso.h:58-72        identical
 *   struct MatchParams                 viso.cpp:48-75      identical
 *   match_desc                         viso.cpp:668-726    identical signature
 *   match_circle                       viso.cpp:206-243    identical signature
 *   collect_matches (Mat x)            viso.cpp:501-514    identical signature
 *   triangulate_rectified<double>      viso.cpp:1137-1162  identical signatures (T = double)
 *   get_inliers                        viso.cpp:1509-1537  identical signature
 *   minimize_reproj                    viso.h:77-79        identical signature
 *   ransac_minimize_reproj             viso.h:74-75        identical signature
 *   tr2mat                             viso.h:162          identical signature
 *   F_from_P<double>                   mvg.h:41-66         F_from_P(P1, P2)
 *   per-frame loop of sequence_odometry viso.cpp:1205-1327 sequence_odometry(P1, P2, FeatureSequence&)
 *   HarrisBinnedFeatureDetector        viso.cpp:911-979    same constructor, detect(image, kp)
 *   MyFeatureExtractor                 viso.cpp:981-1025   same constructor, compute(image, kp, d)
 *   sequence_odometry                  viso.h:138-139      sequence_odometry(P1, P2, StereoImageSource&): images in,
 *                                                          poses out; StereoImageGenerator's cv::imread (viso.h:81-101)
 *                                                          is replaced by any source of 8-bit image pairs
 *
 * RANSAC sampling: the reference seeds a fresh std::mt19937 from std::random_device per hypothesis
 * (viso.cpp:93-95), which is not reproducible.  Here the triples come from one std::mt19937 stream through the
 * reference's own Algorithm S (viso.cpp:87-107); viso_b200::set_ransac_seed() chooses the stream
 * (default 424242), viso_b200::set_sample_table() installs an explicit table for the next call.
 */
#ifndef VISO_B200_HOST_VISO_H_
#define VISO_B200_HOST_VISO_H_

#include "cvcompat.h"

#include <climits>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

using cv::KeyPoint;
using cv::Mat;
using cv::Point2f;
using cv::Vec3i;
using cv::Vec4i;
using std::pair;
using std::vector;

typedef vector<KeyPoint> KeyPoints;
typedef Vec3i Match; // i1, i2, dist
typedef vector<Match> Matches;
typedef vector<Point2f> Points2f;
typedef Mat Descriptors;

/* reference src/viso.h:58-72 */
struct param
{
    param() : ransac_iter(50), inlier_threshold(2), thresh(1e-4), save_debug(true) {}
    double base;
    int ransac_iter;
    double inlier_threshold;
    double thresh; /* gradient norm threshold */
    bool save_debug;
    struct
    {
        double f;
        double cu;
        double cv;
    } calib;
};

/* reference src/viso.cpp:48-75 */
struct MatchParams
{
    bool enforce_epipolar;
    Mat F;
    double alg_thresh;
    double sampson_thresh;
    bool enforce_2nd_best;
    double ratio_2nd_best;
    bool allow_ann;
    int max_neighbors;
    double radius;

    MatchParams(Mat F_) : enforce_epipolar(true), alg_thresh(0), sampson_thresh(1), enforce_2nd_best(false),
                          ratio_2nd_best(.8), allow_ann(true), max_neighbors(200), radius(80)
    {
        F_.copyTo(this->F);
    }
    MatchParams() : enforce_epipolar(false), alg_thresh(0), sampson_thresh(0), enforce_2nd_best(true),
                    ratio_2nd_best(.9), allow_ann(true), max_neighbors(250), radius(80) {}
};

struct viso_b200_error : std::runtime_error {
    int status;
    viso_b200_error(int s, const std::string& what) : std::runtime_error(what), status(s) {}
};

namespace viso_b200 {
void set_device(int device);            /* before the first call; default 0 */
void set_ransac_seed(uint32_t seed);    /* restarts the sample stream */
void set_sample_table(const vector<int>& table /* ransac_iter x 3, consumed by the next ransac call */);
long long kernel_launches();
}

void match_desc(const KeyPoints& kp1, const KeyPoints& kp2, const Descriptors& d1, const Descriptors& d2,
                Matches& match, const MatchParams& sp = MatchParams());

void match_circle(const Matches& match_lr, const Matches& match_lr_prev, const Matches& match11,
                  const Matches& match22, vector<Vec4i>& circ_match, Matches& match_pcl);

void collect_matches(const KeyPoints& kp1, const KeyPoints& kp2, const Matches& match, Mat& x);

/* viso.cpp:1137-1154 (only T = double is used by the pipeline, viso.cpp:1247) and the param overload :1155-1162 */
template <typename T> Mat triangulate_rectified(const Mat& x, double f, double base, double c1u, double c1v);
template <> Mat triangulate_rectified<double>(const Mat& x, double f, double base, double c1u, double c1v);
template <typename T> Mat triangulate_rectified(const Mat& x, const struct param& param)
{
    return triangulate_rectified<T>(x, param.calib.f, param.base, param.calib.cu, param.calib.cv);
}

pair<vector<int>, double> get_inliers(const Mat& X, const Mat& observe, vector<double>& tr,
                                      const struct param& param);

bool minimize_reproj(const Mat& X, const Mat& observe, vector<double>& tr, const struct param& param,
                     const vector<int>& active);

bool ransac_minimize_reproj(const Mat& X, const Mat& observe, vector<double>& best_tr, vector<int>& best_inliers,
                            const struct param& param);

void tr2mat(vector<double> tr, Mat& Tr);

Mat F_from_P(const Mat& P1, const Mat& P2);

/* reference src/viso.cpp:911-979.  detect() clears nothing (cv::FeatureDetector::detect does: kp is replaced). The
 * reference leaves m_k uninitialised (:915-919, :978); here the constructor argument is stored.  Keypoint order
 * inside a bin: ascending (|response|, x, y) (include/viso_b200.h, viso_detect_harris). */
class HarrisBinnedFeatureDetector {
public:
    HarrisBinnedFeatureDetector(int radius, int n, int nbinx = 24, int nbiny = 5, float k = .04f, int block_size = 3,
                                int aperture_size = 5);
    void detect(const Mat& image /* CV_8U */, KeyPoints& kp) const;

private:
    int m_radius, m_nbinx, m_nbiny, m_n;
    float m_k;
};

/* reference src/viso.cpp:981-1025: d = kp.size() x (2r+1)^2 CV_32F; only r = 5 (the pipeline's, viso.cpp:1174) */
class MyFeatureExtractor {
public:
    explicit MyFeatureExtractor(int descriptor_radius);
    int descriptorSize() const { return (2 * m_descriptor_radius + 1) * (2 * m_descriptor_radius + 1); }
    void compute(const Mat& image /* CV_8U */, KeyPoints& kp, Mat& d) const;

private:
    int m_descriptor_radius;
};

typedef pair<Mat, Mat> image_pair;   /* reference src/viso.h: left, right (CV_8U, same size) */

/* Source of stereo pairs: the analogue of StereoImageGenerator::operator() (viso.h:86-96) without cv::imread;
 * next() returns false when the sequence ends (the reference stops at the first unreadable pair). */
class StereoImageSource {
public:
    virtual ~StereoImageSource() {}
    virtual bool next(image_pair& out) = 0;
};

namespace viso_b200 {
void set_max_features(int n);   /* MAX_FEATURE_NUM of sequence_odometry; the reference hard-codes 1200 (viso.cpp:1172) */
}

/* sequence_odometry (viso.h:138-139, viso.cpp:1167-1330) from images: Harris detection, descriptors, matching, circle
 * closure, triangulation and RANSAC/Gauss-Newton all on the device; only the images go up and 64 bytes per frame
 * pair come back.  Debug image dumps (dbg_dir) are not produced. */
vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, StereoImageSource& images);

/* One frame's front-end output (HarrisBinnedFeatureDetector + MyFeatureExtractor, viso.cpp:1226-1231). */
struct FrameFeatures {
    KeyPoints kp1, kp2;   /* left, right */
    Descriptors d1, d2;   /* n x 121 CV_32F */
};

/* Source of per-frame features, the analogue of StereoImageGenerator (viso.h:81-101) after the front-end. */
class FeatureSequence {
public:
    virtual ~FeatureSequence() {}
    virtual size_t size() const = 0;
    virtual const FrameFeatures& frame(size_t t) = 0;
};

/* the per-frame loop of sequence_odometry (viso.cpp:1205-1327): returns the chained 4x4 poses, identity first; frames
 * with < 3 circular matches or a failed RANSAC append nothing (viso.cpp:1283-1288, 1322-1324).  All frames are
 * processed in ONE batched device submission (frame pairs are independent given the features). */
vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, FeatureSequence& frames);

#endif
