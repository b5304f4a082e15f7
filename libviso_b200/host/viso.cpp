/*
 * viso.cpp (B200) -- bodies of the reference's hot-path functions (see viso.h for the file:line map), each a thin
 * marshalling layer over the C-ABI of include/viso_b200.h.  No arithmetic of the path happens here: the only host
 * work is the std::vector / cv::Mat <-> plain-array conversion and the reference's own sampler (Algorithm S,
 * viso.cpp:87-107) that fills the RANSAC sample table.
 */
#include "viso.h"
#include "mvg.h"
#include "estimation.h"

#include "../../include/viso_b200.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>

namespace {

struct Global {
    std::mutex mu;
    viso_ctx* ctx = nullptr;
    int device = 0;
    uint32_t seed = 424242;
    uint64_t stream_pos = 0;          /* hypotheses drawn so far from the seed's stream */
    vector<int> table_override;
    ~Global() { if (ctx) viso_destroy(ctx); }
};

Global& G()
{
    static Global g;
    return g;
}

viso_ctx* ctx()
{
    Global& g = G();
    if (!g.ctx) {
        const int rc = viso_create(&g.ctx, g.device);
        if (rc != VISO_OK) throw viso_b200_error(rc, "viso_create failed: no usable CUDA device (libviso_b200 has no CPU fallback)");
    }
    return g.ctx;
}

void ck(int rc)
{
    if (rc != VISO_OK) throw viso_b200_error(rc, viso_last_error(G().ctx));
}

vector<float> kp_array(const KeyPoints& kp)
{
    vector<float> a(kp.size() * 2);
    for (size_t i = 0; i < kp.size(); ++i) { a[2 * i] = kp[i].pt.x; a[2 * i + 1] = kp[i].pt.y; }
    return a;
}

vector<int32_t> match_array(const Matches& m)
{
    vector<int32_t> a(m.size() * 3);
    for (size_t i = 0; i < m.size(); ++i) { a[3 * i] = m[i][0]; a[3 * i + 1] = m[i][1]; a[3 * i + 2] = m[i][2]; }
    return a;
}

viso_param to_c(const struct param& p)
{
    viso_param c;
    std::memset(&c, 0, sizeof(c));
    c.base = p.base; c.f = p.calib.f; c.cu = p.calib.cu; c.cv = p.calib.cv;
    c.inlier_threshold = p.inlier_threshold; c.thresh = p.thresh; c.ransac_iter = p.ransac_iter;
    return c;
}

/* 3 x n / 4 x n CV_64F, continuous */
const double* mat64(const Mat& m, int rows)
{
    assert(m.type() == cv::DataType<double>::type && m.rows == rows && m.isContinuous());
    (void)rows;
    return m.ptr<double>(0);
}

} // namespace

namespace viso_b200 {

void set_device(int device)
{
    std::lock_guard<std::mutex> l(G().mu);
    G().device = device;
}

void set_ransac_seed(uint32_t seed)
{
    std::lock_guard<std::mutex> l(G().mu);
    G().seed = seed;
    G().stream_pos = 0;
}

void set_sample_table(const vector<int>& table)
{
    std::lock_guard<std::mutex> l(G().mu);
    G().table_override = table;
}

long long kernel_launches()
{
    std::lock_guard<std::mutex> l(G().mu);
    return G().ctx ? (long long)viso_launch_count(G().ctx) : 0;
}

} // namespace viso_b200

/* reference src/viso.cpp:668-726 */
void match_desc(const KeyPoints& kp1, const KeyPoints& kp2, const Descriptors& d1, const Descriptors& d2,
                Matches& match, const MatchParams& sp)
{
    std::lock_guard<std::mutex> l(G().mu);
    match.clear(); /* viso.cpp:675 */
    assert(d1.cols == d2.cols || d1.rows == 0 || d2.rows == 0); /* BOOST_ASSERT_MSG, viso.cpp:676 */
    assert((int)kp1.size() == d1.rows && (int)kp2.size() == d2.rows);
    if (kp1.empty()) return;
    viso_match_params mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.enforce_epipolar = sp.enforce_epipolar;
    mp.enforce_2nd_best = sp.enforce_2nd_best;
    mp.max_neighbors = sp.max_neighbors;
    mp.radius = sp.radius;
    mp.sampson_thresh = sp.sampson_thresh;
    mp.ratio_2nd_best = sp.ratio_2nd_best;
    if (sp.enforce_epipolar) {
        assert(sp.F.rows == 3 && sp.F.cols == 3 && sp.F.type() == cv::DataType<double>::type); /* viso.cpp:393,658 */
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) mp.F[3 * r + c] = sp.F.at<double>(r, c);
    }
    const vector<float> a1 = kp_array(kp1), a2 = kp_array(kp2);
    const int dlen = d1.cols;
    vector<int32_t> out(kp1.size() * 3);
    int32_t n = 0;
    ck(viso_match_desc_sorted(ctx(), a1.data(), (int)kp1.size(), a2.data(), (int)kp2.size(),
                              d1.ptr<float>(0), kp2.empty() ? nullptr : d2.ptr<float>(0), dlen, &mp, out.data(), &n));
    match.reserve(n);
    for (int i = 0; i < n; ++i) match.push_back(Match(out[3 * i], out[3 * i + 1], out[3 * i + 2]));
}

/* reference src/viso.cpp:206-243 */
void match_circle(const Matches& match_lr, const Matches& match_lr_prev, const Matches& match11,
                  const Matches& match22, vector<Vec4i>& circ_match, Matches& match_pcl)
{
    std::lock_guard<std::mutex> l(G().mu);
    const vector<int32_t> a = match_array(match_lr), b = match_array(match_lr_prev), c = match_array(match11),
                          d = match_array(match22);
    vector<int32_t> circ(match_lr.size() * 4 + 4), pcl(match_lr.size() * 3 + 3);
    int32_t n = 0;
    ck(viso_match_circle(ctx(), a.data(), (int)match_lr.size(), b.data(), (int)match_lr_prev.size(), c.data(),
                         (int)match11.size(), d.data(), (int)match22.size(), circ.data(), pcl.data(), &n));
    for (int i = 0; i < n; ++i) {
        circ_match.push_back(Vec4i(circ[4 * i], circ[4 * i + 1], circ[4 * i + 2], circ[4 * i + 3])); /* :233 */
        match_pcl.push_back(Match(pcl[3 * i], pcl[3 * i + 1], 0));                                     /* :234 */
    }
}

/* reference src/viso.cpp:501-514 */
void collect_matches(const KeyPoints& kp1, const KeyPoints& kp2, const Matches& match, Mat& x)
{
    std::lock_guard<std::mutex> l(G().mu);
    x.create(4, (int)match.size(), cv::DataType<double>::type);
    if (match.empty()) return;
    const vector<float> a1 = kp_array(kp1), a2 = kp_array(kp2);
    const vector<int32_t> m = match_array(match);
    ck(viso_collect_triangulate(ctx(), a1.data(), (int)kp1.size(), a2.data(), (int)kp2.size(), m.data(),
                                (int)match.size(), 0, 0, 0, 0, x.ptr<double>(0), nullptr));
}

/* reference src/viso.cpp:1137-1154 */
template <> Mat triangulate_rectified<double>(const Mat& x, double f, double base, double c1u, double c1v)
{
    std::lock_guard<std::mutex> l(G().mu);
    assert(x.type() == cv::DataType<double>::type); /* viso.cpp:1145 */
    Mat X(3, x.cols, cv::DataType<double>::type);
    if (x.cols > 0) ck(viso_triangulate_rectified_f64(ctx(), mat64(x, 4), x.cols, f, base, c1u, c1v, X.ptr<double>(0)));
    return X;
}

/* reference src/viso.cpp:1509-1537.  The returned double is the reference's log-only "rms" (sqrt of the LAST
 * point's squared error over N, :1535); it is not computed on the device and is returned as NaN. */
pair<vector<int>, double> get_inliers(const Mat& X, const Mat& observe, vector<double>& tr, const struct param& param)
{
    std::lock_guard<std::mutex> l(G().mu);
    assert(tr.size() == 6);
    vector<int> inl(X.cols > 0 ? X.cols : 1);
    int32_t n = 0;
    const viso_param p = to_c(param);
    ck(viso_get_inliers(ctx(), X.cols ? mat64(X, 3) : nullptr, X.cols ? mat64(observe, 4) : nullptr, X.cols, tr.data(),
                        &p, inl.data(), &n));
    inl.resize(n);
    return std::make_pair(inl, std::nan(""));
}

/* reference src/viso.cpp:1583-1623 */
bool minimize_reproj(const Mat& X, const Mat& observe, vector<double>& tr, const struct param& param,
                     const vector<int>& active)
{
    std::lock_guard<std::mutex> l(G().mu);
    assert(tr.size() == 6);
    const viso_param p = to_c(param);
    int32_t ok = 0;
    ck(viso_minimize_reproj(ctx(), mat64(X, 3), mat64(observe, 4), X.cols, tr.data(), &p, active.data(),
                            (int)active.size(), &ok));
    return ok != 0;
}

/* reference src/viso.cpp:1543-1580 */
bool ransac_minimize_reproj(const Mat& X, const Mat& observe, vector<double>& best_tr, vector<int>& best_inliers,
                            const struct param& param)
{
    std::lock_guard<std::mutex> l(G().mu);
    Global& g = G();
    best_inliers.clear(); /* viso.cpp:1554 */
    assert(best_tr.size() == 6);
    const int N = X.cols, H = param.ransac_iter;
    vector<int> table;
    if (!g.table_override.empty()) {
        table.swap(g.table_override);
        if ((int)table.size() != 3 * H) throw viso_b200_error(VISO_ERR_ARG, "sample table must hold ransac_iter x 3 indices");
    } else if (N >= 3 && H > 0) {
        /* randomsample(3, N, sample) per hypothesis (viso.cpp:1558), drawn from ONE mt19937(seed) stream: the table
         * for this call starts where the previous call stopped, so successive calls see fresh samples like the
         * reference's, but reproducibly */
        vector<int> all((size_t)3 * (g.stream_pos + H));
        viso_randomsample_table(g.seed, (int)(g.stream_pos + H), N, all.data());
        table.assign(all.begin() + 3 * g.stream_pos, all.end());
        g.stream_pos += H;
    }
    const viso_param p = to_c(param);
    vector<int> inl(N > 0 ? N : 1);
    int32_t n_inl = 0, ok = 0;
    ck(viso_ransac_minimize_reproj(ctx(), N ? mat64(X, 3) : nullptr, N ? mat64(observe, 4) : nullptr, N, &p,
                                   table.empty() ? nullptr : table.data(), best_tr.data(), inl.data(), &n_inl, &ok,
                                   nullptr, nullptr, nullptr, nullptr));
    best_inliers.assign(inl.begin(), inl.begin() + n_inl);
    return ok != 0;
}

/* reference src/viso.cpp:109-133 */
void tr2mat(vector<double> tr, Mat& Tr)
{
    assert(tr.size() == 6);
    Tr.create(4, 4, cv::DataType<double>::type);
    viso_tr2mat(tr.data(), Tr.ptr<double>(0));
}

/* reference src/mvg.cpp:124-169 */
Mat triangulate_dlt(const Mat& x1, const Mat& x2, const Mat& P1, const Mat& P2)
{
    std::lock_guard<std::mutex> l(G().mu);
    assert(x1.cols == x2.cols && x1.rows == 2 && x2.rows == 2);                                   /* mvg.cpp:127-129 */
    assert(x1.type() == cv::DataType<float>::type && x2.type() == cv::DataType<float>::type);     /* :130-131 */
    assert(P1.type() == cv::DataType<double>::type && P2.type() == cv::DataType<double>::type);   /* :132-133 */
    Mat X(3, x1.cols, cv::DataType<float>::type);
    if (x1.cols > 0)
        ck(viso_triangulate_dlt(ctx(), x1.ptr<float>(0), x2.ptr<float>(0), x1.cols, P1.ptr<double>(0), P2.ptr<double>(0),
                                X.ptr<float>(0)));
    return X;
}

/* reference src/mvg.cpp:172-192 */
Mat triangulate_rectified(const Mat& x1, const Mat& x2, double f, double base, double c1u, double c1v)
{
    std::lock_guard<std::mutex> l(G().mu);
    assert(x1.cols == x2.cols);                                                                   /* mvg.cpp:180 */
    assert(x1.type() == cv::DataType<float>::type && x2.type() == cv::DataType<float>::type);     /* :181-182 */
    Mat X(3, x1.cols, cv::DataType<float>::type);
    if (x1.cols > 0)
        ck(viso_triangulate_rectified_f32(ctx(), x1.ptr<float>(0), x2.ptr<float>(0), x1.cols, f, base, c1u, c1v, X.ptr<float>(0)));
    return X;
}

/* reference src/estimation.cpp:29-51 */
void solveRigidMotion(const Mat& A, const Mat& B, Mat& T)
{
    std::lock_guard<std::mutex> l(G().mu);
    assert(A.cols > 1 && A.cols == B.cols && A.rows == B.rows && A.rows == 3); /* BOOST_ASSERT_MSG, estimation.cpp:32-39 */
    T.create(4, 4, cv::DataType<float>::type);
    ck(viso_solve_rigid_motion(ctx(), A.ptr<float>(0), B.ptr<float>(0), A.cols, T.ptr<float>(0)));
}

/* the reference's signature (src/estimation.h:7-9) */
void solveRigidMotion(const Eigen::MatrixXf& A, const Eigen::MatrixXf& B, Eigen::Affine3f& T)
{
    assert(A.cols() > 1 && A.cols() == B.cols() && A.rows() == B.rows() && A.rows() == 3); /* estimation.cpp:32-39 */
    const int n = (int)A.cols();
    Mat a(3, n, CV_32FC1), b(3, n, CV_32FC1), t;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < n; ++c) { a.at<float>(r, c) = A(r, c); b.at<float>(r, c) = B(r, c); }
    solveRigidMotion(a, b, t);
    Eigen::Matrix3f R;
    Eigen::Vector3f tv;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R(r, c) = t.at<float>(r, c); tv(r) = t.at<float>(r, 3); }
    T = R;                  /* estimation.cpp:49-50 */
    T.translation() = tv;
}

/* reference src/viso.cpp:469-483: host bookkeeping (point copies) */
void collect_matches(const KeyPoints& kp1, const KeyPoints& kp2, const Matches& match, Points2f& p1, Points2f& p2, int lim)
{
    p1.clear(); p2.clear();
    int i = 0;
    for (const Match& m : match) {
        if (i >= lim) break;
        p1.push_back(kp1.at(m[0]).pt);
        p2.push_back(kp2.at(m[1]).pt);
        ++i;
    }
}

/* reference src/mvg.cpp:33-42 */
vector<int> arange(int range)
{
    assert(range >= 0);
    vector<int> v(range);
    for (int i = 0; i < range; ++i) v[i] = i;
    return v;
}

/* reference src/mvg.cpp:92-107: P = K [R t], a once-per-rig 3 x 4 product (float storage, t read as T) */
template <class T> Mat P_from_KRt(const Mat& K, const Mat& R, const Mat& t)
{
    Mat P(3, 4, CV_32FC1);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) P.at<float>(i, j) = R.at<float>(i, j);
        P.at<float>(i, 3) = (float)t.at<T>(i);
    }
    return K * P;
}
template Mat P_from_KRt<float>(const Mat&, const Mat&, const Mat&);
template Mat P_from_KRt<double>(const Mat&, const Mat&, const Mat&);

/* reference src/misc.cpp:3-15 */
bool isEqual(double x, double y) { return std::abs(x - y) <= 1e-6 * std::abs(x); }
bool isEqual(float x, float y) { return std::abs(x - y) <= 1e-6f * std::abs(x); }

/* the per-frame loop of reference src/viso.cpp:1167-1330, batched */
vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, FeatureSequence& frames)
{
    std::lock_guard<std::mutex> l(G().mu);
    Global& g = G();
    vector<Mat> poses;
    poses.push_back(Mat::eye(4, 4, cv::DataType<double>::type)); /* viso.cpp:1189-1190 */
    const int F = (int)frames.size();
    if (F == 0) return poses;
    int cap = 1, dlen = 0;
    for (int t = 0; t < F; ++t) {
        const FrameFeatures& f = frames.frame(t);
        cap = std::max(cap, (int)std::max(f.kp1.size(), f.kp2.size()));
        if (f.d1.rows > 0) dlen = f.d1.cols;
    }
    if (dlen == 0) return poses;
    struct param prm; /* viso.cpp:1182: defaults; base / calib are derived from P1, P2 by viso_seq_set_calib */
    prm.base = 0; prm.calib.f = 0; prm.calib.cu = 0; prm.calib.cv = 0;
    viso_seq* seq = nullptr;
    ck(viso_seq_create(ctx(), F, cap, dlen, prm.ransac_iter, &seq));
    try {
        ck(viso_seq_set_calib(seq, p1.ptr<double>(0), p2.ptr<double>(0)));
        for (int t = 0; t < F; ++t) {
            const FrameFeatures& f = frames.frame(t);
            const vector<float> a1 = kp_array(f.kp1), a2 = kp_array(f.kp2);
            ck(viso_seq_upload_frame(seq, t, a1.data(), (int)f.kp1.size(), a2.data(), (int)f.kp2.size(),
                                     f.kp1.empty() ? nullptr : f.d1.ptr<float>(0),
                                     f.kp2.empty() ? nullptr : f.d2.ptr<float>(0)));
            ck(viso_sync(ctx())); /* a1 / a2 are pageable temporaries */
        }
        /* sample seeds for every (frame pair, hypothesis): one mt19937(seed) stream */
        vector<uint32_t> seeds((size_t)F * prm.ransac_iter * 3);
        {
            std::mt19937 gen(g.seed);
            for (auto& s : seeds) s = (uint32_t)gen();
        }
        viso_param p = to_c(prm);
        ck(viso_seq_run(seq, &p, seeds.data()));
        vector<viso_record> rec(F);
        ck(viso_seq_download(seq, rec.data()));
        vector<double> chained((size_t)F * 16);
        const int np = viso_chain_poses(rec.data(), F, chained.data()); /* viso.cpp:1313-1321 */
        for (int i = 1; i < np; ++i) {
            Mat pose(4, 4, cv::DataType<double>::type);
            std::memcpy(pose.ptr<double>(0), &chained[(size_t)16 * i], 16 * sizeof(double));
            poses.push_back(pose);
        }
    } catch (...) {
        viso_seq_destroy(seq);
        throw;
    }
    viso_seq_destroy(seq);
    return poses;
}

/* ---- front end: detector, extractor, images -> poses ---- */

namespace {
int initial_max_features()
{
    const char* e = std::getenv("VISO_MAX_FEATURES");
    const int v = e ? std::atoi(e) : 0;
    return v > 0 ? v : 1200;   /* MAX_FEATURE_NUM, viso.cpp:1172 */
}
int g_max_features = initial_max_features();

const unsigned char* image8(const Mat& image)
{
    if (image.empty() || image.type() != cv::DataType<unsigned char>::type || !image.isContinuous())
        throw viso_b200_error(VISO_ERR_ARG, "expected a continuous 8-bit single-channel image");
    return image.ptr<unsigned char>(0);
}
} // namespace

namespace viso_b200 {
void set_max_features(int n)
{
    std::lock_guard<std::mutex> l(G().mu);
    g_max_features = n;
}
} // namespace viso_b200

HarrisBinnedFeatureDetector::HarrisBinnedFeatureDetector(int radius, int n, int nbinx, int nbiny, float k, int block_size,
                                                         int aperture_size)
    : m_radius(radius), m_nbinx(nbinx), m_nbiny(nbiny), m_n(n), m_k(k)
{
    assert(nbinx > 0 && nbiny > 0);   /* viso.cpp:919 */
    if (block_size != 3 || aperture_size != 5)
        throw viso_b200_error(VISO_ERR_DOMAIN, "HarrisBinnedFeatureDetector: only block_size 3, aperture_size 5 (the reference's defaults)");
}

void HarrisBinnedFeatureDetector::detect(const Mat& image, KeyPoints& kp) const
{
    std::lock_guard<std::mutex> l(G().mu);
    const unsigned char* img = image8(image);
    vector<float> xy((size_t)std::max(m_n, 1) * 2), resp(std::max(m_n, 1));
    int32_t n = 0;
    ck(viso_detect_harris(ctx(), img, image.cols, image.rows, image.cols, m_n, m_nbinx, m_nbiny, m_k, xy.data(), resp.data(), &n));
    kp.clear();
    kp.reserve(n);
    for (int i = 0; i < n; ++i) {
        KeyPoint k(xy[2 * i], xy[2 * i + 1], (float)(2 * m_radius + 1));   /* viso.cpp:967-969 */
        k.response = resp[i];
        kp.push_back(k);
    }
}

MyFeatureExtractor::MyFeatureExtractor(int descriptor_radius) : m_descriptor_radius(descriptor_radius)
{
    if (descriptor_radius != 5) throw viso_b200_error(VISO_ERR_DOMAIN, "MyFeatureExtractor: only descriptor radius 5 (viso.cpp:1174)");
}

void MyFeatureExtractor::compute(const Mat& image, KeyPoints& kp, Mat& d) const
{
    std::lock_guard<std::mutex> l(G().mu);
    const unsigned char* img = image8(image);
    d = Mat((int)kp.size(), descriptorSize(), cv::DataType<float>::type, Scalar(0));   /* viso.cpp:1008 */
    if (kp.empty()) return;
    const vector<float> a = kp_array(kp);
    ck(viso_extract_descriptors(ctx(), img, image.cols, image.rows, image.cols, a.data(), (int)kp.size(), d.ptr<float>(0)));
}

vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, StereoImageSource& images)
{
    vector<Mat> poses;
    poses.push_back(Mat::eye(4, 4, cv::DataType<double>::type)); /* viso.cpp:1189-1190 */
    vector<image_pair> pairs;
    image_pair ip;
    while (images.next(ip)) {                                    /* viso.cpp:1205 */
        if (ip.first.rows != ip.second.rows || ip.first.cols != ip.second.cols ||
            (!pairs.empty() && (ip.first.rows != pairs[0].first.rows || ip.first.cols != pairs[0].first.cols)))
            throw viso_b200_error(VISO_ERR_ARG, "sequence_odometry: all images of a sequence must have one size");
        image8(ip.first); image8(ip.second);
        pairs.push_back(image_pair(ip.first.clone(), ip.second.clone()));
    }
    std::lock_guard<std::mutex> l(G().mu);
    Global& g = G();
    const int F = (int)pairs.size();
    if (F == 0) return poses;
    struct param prm;
    prm.base = 0; prm.calib.f = 0; prm.calib.cu = 0; prm.calib.cv = 0;
    viso_seq* seq = nullptr;
    ck(viso_seq_create(ctx(), F, std::max(g_max_features, 1), 121, prm.ransac_iter, &seq));
    try {
        ck(viso_seq_set_calib(seq, p1.ptr<double>(0), p2.ptr<double>(0)));
        ck(viso_seq_set_image_size(seq, pairs[0].first.cols, pairs[0].first.rows));
        ck(viso_seq_set_detector(seq, g_max_features, 24, 5, .04f));   /* detector(5, MAX_FEATURE_NUM), viso.cpp:1173 */
        for (int t = 0; t < F; ++t) ck(viso_seq_upload_frame_raw(seq, t, pairs[t].first.ptr<unsigned char>(0), pairs[t].second.ptr<unsigned char>(0)));
        ck(viso_sync(ctx()));
        vector<uint32_t> seeds((size_t)F * prm.ransac_iter * 3);
        {
            std::mt19937 gen(g.seed);
            for (auto& s : seeds) s = (uint32_t)gen();
        }
        viso_param p = to_c(prm);
        ck(viso_seq_run(seq, &p, seeds.data()));
        vector<viso_record> rec(F);
        ck(viso_seq_download(seq, rec.data()));
        vector<double> chained((size_t)F * 16);
        const int np = viso_chain_poses(rec.data(), F, chained.data()); /* viso.cpp:1313-1321 */
        for (int i = 1; i < np; ++i) {
            Mat pose(4, 4, cv::DataType<double>::type);
            std::memcpy(pose.ptr<double>(0), &chained[(size_t)16 * i], 16 * sizeof(double));
            poses.push_back(pose);
        }
    } catch (...) {
        viso_seq_destroy(seq);
        throw;
    }
    viso_seq_destroy(seq);
    return poses;
}

/* reference src/viso.h:138-139 / src/viso.cpp:1167-1330: the generator is drained (cv::imread per frame, exactly the
 * reference's loop condition at :1205) and the whole sequence goes to the device in one submission.  dbg_dir only
 * receives debug JPEGs in the reference (:1232-1310); nothing is written here. */
vector<Mat> sequence_odometry(const Mat& p1, const Mat& p2, StereoImageGenerator& images, const boost::filesystem::path& dbg_dir)
{
    (void)dbg_dir;
    struct Adapter : StereoImageSource {
        StereoImageGenerator& gen;
        explicit Adapter(StereoImageGenerator& g_) : gen(g_) {}
        bool next(image_pair& out)
        {
            StereoImageGenerator::result_type r = gen();
            if (!r) return false;
            out = *r;
            return true;
        }
    } src(images);
    return sequence_odometry(p1, p2, src);
}
