/*
 * viso.h (B200) -- the reference's public header (reference src/viso.h) restated declaration for declaration, so that
 * code written against the reference -- src/kitti.cpp, test/test.cpp -- builds against this library unchanged: same
 * includes, same using-declarations and typedefs (callers rely on them), `struct param`, the two image generators,
 * and every function the reference defines, with the reference's signatures.  The bodies (viso.cpp in this
 * directory) are marshalling layers over the C-ABI of include/viso_b200.h; all arithmetic of the per-frame path runs
 * in libviso_b200.so on the GPU.  There is no CPU fallback: without a CUDA device every call throws viso_b200_error.
 *
 *   reference src/viso.h                         here
 *   :29-56   using / typedefs                    identical
 *   :58-72   struct param                        identical
 *   :74-79   ransac_minimize_reproj, minimize_reproj            identical signatures
 *   :81-119  StereoImageGenerator, MonoImageGenerator           same members, same behaviour (cv::imread per frame)
 *   :138-139 sequence_odometry(P1, P2, StereoImageGenerator&, dbg_dir)   identical signature; dbg_dir is accepted
 *            and unused (the debug JPEG dumps of viso.cpp:1232-1310 are not produced)
 *   :145-147 collect_matches(..., Points2f&, Points2f&, lim)    identical signature
 *   :162     tr2mat                                             identical signature
 *   :124-136, :142-144, :149-160  readCameraParams, getFundamentalMat, findConstrainedCorrespondences,
 *            match_l2_2nd_best, match_epip_constraint (declared, never defined by the reference), save2 (debug
 *            drawing), calibratedSFM (mono pipeline): declared for source compatibility, not defined
 *
 * Below the reference's surface: the hot-path functions that have external linkage in the reference's viso.cpp but no
 * declaration in its header (match_desc, match_circle, collect_matches(Mat), triangulate_rectified<T>, get_inliers,
 * MatchParams), the two front-end classes, and the knobs of this implementation (namespace viso_b200).
 *
 * RANSAC sampling: the reference seeds a fresh std::mt19937 from std::random_device per hypothesis
 * (viso.cpp:93-95), which is not reproducible.  Here the triples come from one std::mt19937 stream through the
 * reference's own Algorithm S (viso.cpp:87-107); viso_b200::set_ransac_seed() chooses the stream
 * (default 424242), viso_b200::set_sample_table() installs an explicit table for the next call.
 */
#ifndef VISO_B200_HOST_VISO_H_
#define VISO_B200_HOST_VISO_H_

/* without an installed OpenCV / Boost / Eigen, compile with -I<repo>/compat (header stand-ins) */
#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <opencv2/calib3d/calib3d.hpp>
#include <opencv2/features2d/features2d.hpp>
#include <opencv2/highgui/highgui.hpp>
#include <opencv2/highgui/highgui_c.h>
#include <opencv2/imgproc/types_c.h>

#include <climits>
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include <boost/format.hpp>
#include <boost/filesystem.hpp>

#include <Eigen/Dense>

#include "mvg.h"
#include "misc.h"

using namespace std;
using namespace boost;

using cv::Mat;
using cv::KeyPoint;
using cv::Vec2i;
using cv::FeatureDetector;
using cv::DescriptorExtractor;
using cv::Scalar;
using cv::FileStorage;
using cv::Vec6f;
using cv::Point2i;
using cv::waitKey;
using cv::Size;
using cv::Point;
using cv::Point2f;
using cv::Vec3i;
using cv::Vec4i;
using Eigen::MatrixXf;
using Eigen::Affine3f;

typedef vector<KeyPoint> KeyPoints;
typedef Vec3i Match; // i1, i2, dist
typedef vector<Match> Matches;
typedef vector<Point2f> Points2f;
typedef Mat Descriptors;
typedef pair<Mat, Mat> image_pair;

/* reference src/viso.h:58-72 (member order is part of the contract: callers and the library share the layout) */
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

/* reference src/viso.h:74-79 */
bool ransac_minimize_reproj(const Mat& X, const Mat& observe, vector<double>& best_tr, vector<int>& best_inliers,
                            const struct param& param);
bool minimize_reproj(const Mat& X, const Mat& observe, vector<double>& tr, const struct param& param,
                     const vector<int>& active);

/* reference src/viso.h:81-101: yields the pairs `mask % index` for index = begin .. end, read with cv::imread as
 * 8-bit gray; the first unreadable pair ends the sequence */
class StereoImageGenerator
{
public:
    typedef boost::optional<image_pair> result_type;
    typedef pair<string, string> string_pair;
    StereoImageGenerator(const string_pair& mask, int begin = 0, int end = INT_MAX) : m_index(begin), m_end(end), m_mask(mask) {}
    result_type operator()()
    {
        if (m_index > m_end) return result_type();
        const string left = str(boost::format(m_mask.first) % m_index), right = str(boost::format(m_mask.second) % m_index);
        image_pair p(cv::imread(left, CV_LOAD_IMAGE_GRAYSCALE), cv::imread(right, CV_LOAD_IMAGE_GRAYSCALE));
        ++m_index;
        return (p.first.data && p.second.data) ? result_type(p) : result_type();
    }

private:
    int m_index, m_end;
    string_pair m_mask;
};

/* reference src/viso.h:103-119 */
class MonoImageGenerator
{
public:
    typedef boost::optional<Mat> result_type;
    MonoImageGenerator(const string& mask, int begin = 0, int end = INT_MAX) : m_index(begin), m_end(end), m_mask(mask) {}
    result_type operator()()
    {
        if (m_index > m_end) return result_type();
        Mat image = cv::imread(str(boost::format(m_mask) % m_index), CV_LOAD_IMAGE_GRAYSCALE);
        ++m_index;
        return image.data ? result_type(image) : result_type();
    }

private:
    int m_index, m_end;
    string m_mask;
};

/* reference src/viso.h:124-136: declared there, defined nowhere in the reference */
void readCameraParams(const string& intrinsics_name, const string& extrinsics_name, StereoCam& p);
Mat getFundamentalMat(const Mat& R1, const Mat& t1, const Mat& R2, const Mat& t2, const Mat& cameraMatrix);
void findConstrainedCorrespondences(const Mat& F, const KeyPoints& kp1, const KeyPoints& kp2, const Mat& d1, const Mat& d2,
                                    Matches& matches, double eps, double ratio);

/* reference src/viso.h:138-139, src/viso.cpp:1167-1330: detection, description, stereo + temporal matching, circle
 * closure, triangulation and RANSAC / Gauss-Newton for every stereo pair the generator yields; returns the chained
 * 4 x 4 CV_64F poses, identity first; a frame with fewer than 3 circular matches or a failed RANSAC appends nothing */
vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, StereoImageGenerator& images, const boost::filesystem::path& dbg_dir);

/* reference src/viso.h:142-160 */
void match_l2_2nd_best(const Descriptors& d1, const Descriptors& d2, Matches& match, float ratio = 0.7);     /* never defined */
void collect_matches(const KeyPoints& kp1, const KeyPoints& kp2, const Matches& match, Points2f& p1, Points2f& p2,
                     int lim = INT_MAX);                                                                      /* viso.cpp:469-483 */
void match_epip_constraint(const cv::Mat& F, const KeyPoints& kp1, const KeyPoints& kp2, const Descriptors& d1,
                           const Descriptors& d2, Matches& match, double ratio, double samp_thresh, double alg_thresh); /* never defined */
void save2(const cv::Mat& m1, const cv::Mat& m2, const KeyPoints& kp1, const KeyPoints& kp2, const Matches& match,
           const string& file_name, int lim = 50);                                                            /* debug drawing: not built */
void calibratedSFM(const Mat& K, MonoImageGenerator& images);                                                 /* mono pipeline: not built */

void tr2mat(vector<double> tr, Mat& Tr); /* reference src/viso.h:162, src/viso.cpp:109-133 */

/* ------------------------------------------------------------------------------------------------------------------
 * Hot-path functions of src/viso.cpp that the reference's header does not declare (same names and signatures)
 * ------------------------------------------------------------------------------------------------------------------ */

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
void set_max_features(int n);           /* MAX_FEATURE_NUM of sequence_odometry; the reference hard-codes 1200 (viso.cpp:1172);
                                           the environment variable VISO_MAX_FEATURES sets the initial value */
long long kernel_launches();
}

void match_desc(const KeyPoints& kp1, const KeyPoints& kp2, const Descriptors& d1, const Descriptors& d2,
                Matches& match, const MatchParams& sp = MatchParams());                       /* viso.cpp:668-726 */
void match_circle(const Matches& match_lr, const Matches& match_lr_prev, const Matches& match11,
                  const Matches& match22, vector<Vec4i>& circ_match, Matches& match_pcl);      /* viso.cpp:206-243 */
void collect_matches(const KeyPoints& kp1, const KeyPoints& kp2, const Matches& match, Mat& x); /* viso.cpp:501-514 */

/* viso.cpp:1137-1154 (only T = double is used by the pipeline, viso.cpp:1247) and the param overload :1155-1162 */
template <typename T> Mat triangulate_rectified(const Mat& x, double f, double base, double c1u, double c1v);
template <> Mat triangulate_rectified<double>(const Mat& x, double f, double base, double c1u, double c1v);
template <typename T> Mat triangulate_rectified(const Mat& x, const struct param& param)
{
    return triangulate_rectified<T>(x, param.calib.f, param.base, param.calib.cu, param.calib.cv);
}

pair<vector<int>, double> get_inliers(const Mat& X, const Mat& observe, vector<double>& tr,
                                      const struct param& param);                              /* viso.cpp:1509-1537 */

/* reference src/viso.cpp:911-979, on the device.  detect() replaces kp (cv::FeatureDetector::detect clears it).  The
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

/* reference src/viso.cpp:981-1025, on the device: d = kp.size() x (2r+1)^2 CV_32F; only r = 5 (viso.cpp:1174) */
class MyFeatureExtractor {
public:
    explicit MyFeatureExtractor(int descriptor_radius);
    int descriptorSize() const { return (2 * m_descriptor_radius + 1) * (2 * m_descriptor_radius + 1); }
    void compute(const Mat& image /* CV_8U */, KeyPoints& kp, Mat& d) const;

private:
    int m_descriptor_radius;
};

/* Any source of 8-bit stereo pairs (frames already in memory, a camera): next() returns false when the sequence ends */
class StereoImageSource {
public:
    virtual ~StereoImageSource() {}
    virtual bool next(image_pair& out) = 0;
};
vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, StereoImageSource& images);

/* One frame's front-end output (HarrisBinnedFeatureDetector + MyFeatureExtractor, viso.cpp:1226-1231). */
struct FrameFeatures {
    KeyPoints kp1, kp2;   /* left, right */
    Descriptors d1, d2;   /* n x 121 CV_32F */
};

class FeatureSequence {
public:
    virtual ~FeatureSequence() {}
    virtual size_t size() const = 0;
    virtual const FrameFeatures& frame(size_t t) = 0;
};

/* the per-frame loop of sequence_odometry (viso.cpp:1205-1327) from precomputed features, ONE batched device submission */
vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, FeatureSequence& frames);

#endif
