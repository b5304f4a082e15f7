// C++ drop-in test of libviso_b200/host (the reference's own function signatures over the C-ABI), in the style of
// the reference's test/test.cpp: test_nl_rigid_motion1 (test.cpp:152-168: same calibration, ransac_minimize_reproj
// must return true) plus the per-frame loop of sequence_odometry, every result compared with the CPU oracle
// (oracle/viso_oracle.h -- test infrastructure) on the same inputs.  Needs a GPU; prints "host test OK".
#include "../../libviso_b200/host/viso.h"
#include "../../libviso_b200/host/mvg.h"
#include "../../libviso_b200/host/estimation.h"
#include "../../include/viso_b200.h"
#include "../../oracle/viso_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

static int g_checks = 0;
#define REQUIRE(c)                                                                 \
    do {                                                                           \
        ++g_checks;                                                                \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } \
    } while (0)

static vector<float> kp_arr(const KeyPoints& kp)
{
    vector<float> a(kp.size() * 2);
    for (size_t i = 0; i < kp.size(); ++i) { a[2 * i] = kp[i].pt.x; a[2 * i + 1] = kp[i].pt.y; }
    return a;
}
static vector<int32_t> m_arr(const Matches& m)
{
    vector<int32_t> a(m.size() * 3);
    for (size_t i = 0; i < m.size(); ++i) for (int j = 0; j < 3; ++j) a[3 * i + j] = m[i][j];
    return a;
}
static bool tr_close(const double* a, const double* b)
{
    double tn = std::sqrt(b[3] * b[3] + b[4] * b[4] + b[5] * b[5]);
    if (tn < 1e-3) tn = 1e-3;
    for (int j = 0; j < 3; ++j) if (std::fabs(a[j] - b[j]) > 1e-6) return false;          // rad
    for (int j = 3; j < 6; ++j) if (std::fabs(a[j] - b[j]) > 1e-6 * tn) return false;     // relative translation
    return true;
}

// stereo projection of viso.cpp:1441-1489 used to synthesise observations
static void project(const double tr[6], const struct param& p, const double X[3], double obs[4])
{
    double T[16];
    vo_tr2mat(tr, T);
    const double x = T[0] * X[0] + T[1] * X[1] + T[2] * X[2] + T[3];
    const double y = T[4] * X[0] + T[5] * X[1] + T[6] * X[2] + T[7];
    const double z = T[8] * X[0] + T[9] * X[1] + T[10] * X[2] + T[11];
    obs[0] = p.calib.f * x / z + p.calib.cu;
    obs[1] = p.calib.f * y / z + p.calib.cv;
    obs[2] = p.calib.f * (x - p.base) / z + p.calib.cu;
    obs[3] = obs[1];
}

static vo_param to_vo(const struct param& p)
{
    vo_param v;
    vo_param_default(&v);
    v.base = p.base; v.f = p.calib.f; v.cu = p.calib.cu; v.cv = p.calib.cv;
    v.inlier_threshold = p.inlier_threshold; v.thresh = p.thresh; v.ransac_iter = p.ransac_iter;
    return v;
}

static void test_nl_rigid_motion1()
{
    struct param p; // test.cpp:158-161
    p.base = .5707;
    p.calib.f = 645.24;
    p.calib.cu = 635.96;
    p.calib.cv = 194.13;
    const int N = 600;
    std::mt19937 gen(7);
    std::uniform_real_distribution<> ux(-20, 20), uy(-2, 3), uz(4, 60), u01(0, 1), uo(-50, 50);
    std::normal_distribution<> noise(0, 0.3);
    const double tr_true[6] = {0.01, -0.02, 0.005, 0.05, -0.02, -1.0};
    Mat X(3, N, CV_64F), observe(4, N, CV_64F);
    for (int i = 0; i < N; ++i) {
        const double Xi[3] = {ux(gen), uy(gen), uz(gen)};
        double ob[4];
        project(tr_true, p, Xi, ob);
        const bool outlier = u01(gen) < 0.3;
        for (int r = 0; r < 3; ++r) X.at<double>(r, i) = Xi[r];
        for (int r = 0; r < 4; ++r) observe.at<double>(r, i) = ob[r] + noise(gen) + (outlier ? uo(gen) : 0.0);
    }
    viso_b200::set_ransac_seed(424242);
    vector<double> tr(6, 0.0);
    vector<int> inliers;
    REQUIRE(ransac_minimize_reproj(X, observe, tr, inliers, p) == true); // test.cpp:166
    std::printf("tr: %g %g %g %g %g %g\n", tr[0], tr[1], tr[2], tr[3], tr[4], tr[5]);
    for (int j = 0; j < 3; ++j) REQUIRE(std::fabs(tr[j] - tr_true[j]) < 5e-3);

    // oracle on the same sample table
    vector<int32_t> table(3 * p.ransac_iter);
    vo_randomsample_table(424242, p.ransac_iter, N, table.data());
    const vo_param vp = to_vo(p);
    double otr[6] = {0, 0, 0, 0, 0, 0};
    vector<int32_t> oinl(N);
    int32_t on = 0;
    const int ook = vo_ransac_minimize_reproj(X.ptr<double>(0), observe.ptr<double>(0), N, &vp, table.data(), otr,
                                              oinl.data(), &on, nullptr, nullptr, nullptr, nullptr);
    REQUIRE(ook == 1);
    REQUIRE((int)inliers.size() == on);
    for (int i = 0; i < on; ++i) REQUIRE(inliers[i] == oinl[i]);
    REQUIRE(tr_close(tr.data(), otr));

    // minimize_reproj (test.cpp:104) and get_inliers on the support set
    vector<double> tr2(6, 0.0);
    REQUIRE(minimize_reproj(X, observe, tr2, p, inliers) == true);
    double otr2[6] = {0, 0, 0, 0, 0, 0};
    REQUIRE(vo_minimize_reproj(X.ptr<double>(0), observe.ptr<double>(0), N, otr2, &vp, oinl.data(), on, nullptr) == 1);
    REQUIRE(tr_close(tr2.data(), otr2));
    pair<vector<int>, double> gi = get_inliers(X, observe, tr, p);
    vector<int32_t> oi(N);
    const int oc = vo_get_inliers(X.ptr<double>(0), observe.ptr<double>(0), N, tr.data(), &vp, oi.data(), nullptr, nullptr);
    REQUIRE((int)gi.first.size() == oc);
    for (int i = 0; i < oc; ++i) REQUIRE(gi.first[i] == oi[i]);

    // fewer than 6 supporters => false (viso.cpp:1571)
    Mat X5(3, 5, CV_64F), o5(4, 5, CV_64F);
    for (int i = 0; i < 5; ++i) {
        for (int r = 0; r < 3; ++r) X5.at<double>(r, i) = X.at<double>(r, i);
        for (int r = 0; r < 4; ++r) o5.at<double>(r, i) = observe.at<double>(r, i);
    }
    vector<double> tr5(6, 0.0);
    vector<int> in5;
    REQUIRE(ransac_minimize_reproj(X5, o5, tr5, in5, p) == false);

    Mat T;
    tr2mat(tr, T);
    double To[16];
    vo_tr2mat(tr.data(), To);
    for (int i = 0; i < 16; ++i) REQUIRE(T.ptr<double>(0)[i] == To[i]);
}

// a small world of 3-D points with fixed random descriptors, seen from a camera moving forward
struct World : FeatureSequence {
    vector<FrameFeatures> fr;
    size_t size() const override { return fr.size(); }
    const FrameFeatures& frame(size_t t) override { return fr[t]; }
};

static void build_world(World& w, const Mat& P1, const Mat& P2, int n_frames, int n_points)
{
    std::mt19937 gen(11);
    std::uniform_real_distribution<> ux(-25, 25), uy(-3, 2), uz(6, 70);
    std::uniform_int_distribution<> ud(-1020, 1020), un(-6, 6);
    vector<double> pts(3 * n_points);
    vector<float> desc((size_t)n_points * 121);
    for (int i = 0; i < n_points; ++i) { pts[3 * i] = ux(gen); pts[3 * i + 1] = uy(gen); pts[3 * i + 2] = uz(gen); }
    for (auto& d : desc) d = (float)ud(gen);
    const double f = P1.at<double>(0, 0), cu = P1.at<double>(0, 2), cv = P1.at<double>(1, 2);
    const double base = std::fabs(P2.at<double>(0, 3) / P2.at<double>(0, 0));
    for (int t = 0; t < n_frames; ++t) {
        FrameFeatures ff;
        vector<float> dl, dr;
        const double zc = 1.0 * t, yaw = 0.01 * t;
        for (int i = 0; i < n_points; ++i) {
            const double X0 = pts[3 * i], Y = pts[3 * i + 1], Z0 = pts[3 * i + 2] - zc;
            const double X = std::cos(yaw) * X0 - std::sin(yaw) * Z0, Z = std::sin(yaw) * X0 + std::cos(yaw) * Z0;
            if (Z < 2) continue;
            const float u1 = (float)std::floor(f * X / Z + cu), v = (float)std::floor(f * Y / Z + cv);
            const float u2 = (float)std::floor(f * (X - base) / Z + cu);
            if (u1 < 6 || u1 > 1234 || u2 < 6 || u2 > 1234 || v < 6 || v > 369 || u1 - u2 < 1) continue;
            ff.kp1.push_back(KeyPoint(u1, v, 11.f));
            ff.kp2.push_back(KeyPoint(u2, v, 11.f));
            for (int k = 0; k < 121; ++k) {
                float a = desc[(size_t)i * 121 + k] + (float)un(gen), b = desc[(size_t)i * 121 + k] + (float)un(gen);
                dl.push_back(std::max(-1020.f, std::min(1020.f, a)));
                dr.push_back(std::max(-1020.f, std::min(1020.f, b)));
            }
        }
        ff.d1.create((int)ff.kp1.size(), 121, CV_32F);
        ff.d2.create((int)ff.kp2.size(), 121, CV_32F);
        std::memcpy(ff.d1.ptr<float>(0), dl.data(), dl.size() * 4);
        std::memcpy(ff.d2.ptr<float>(0), dr.data(), dr.size() * 4);
        w.fr.push_back(ff);
    }
}

static void test_frame_loop()
{
    // KITTI-00 calibration, test.cpp:56-65
    Mat P1 = Mat::zeros(3, 4, CV_64F), P2;
    P1.at<double>(0, 0) = 718.856; P1.at<double>(0, 2) = 607.1928;
    P1.at<double>(1, 1) = 718.856; P1.at<double>(1, 2) = 185.2157;
    P1.at<double>(2, 2) = 1;
    P2 = P1.clone();
    P2.at<double>(0, 3) = -386.1448;
    World w;
    build_world(w, P1, P2, 4, 900);

    // ---- the reference's per-frame sequence, function by function (viso.cpp:1240-1313), vs the oracle
    Mat F = F_from_P<double>(P1, P2);
    if (F.at<double>(2, 2) > 2.2250738585072014e-308) { // viso.cpp:1177-1180, the reference's own statement
        F /= F.at<double>(2, 2);
    }
    double Fo[9];
    vo_F_from_P(P1.ptr<double>(0), P2.ptr<double>(0), 1, Fo);
    for (int i = 0; i < 9; ++i) REQUIRE(F.ptr<double>(0)[i] == Fo[i]);
    struct param prm;
    prm.base = std::fabs(P2.at<double>(0, 3) / P2.at<double>(0, 0));
    prm.calib.f = P1.at<double>(0, 0); prm.calib.cu = P1.at<double>(0, 2); prm.calib.cv = P1.at<double>(1, 2);

    const FrameFeatures &f0 = w.fr[0], &f1 = w.fr[1];
    auto check_match = [&](const KeyPoints& ka, const KeyPoints& kb, const Mat& da, const Mat& db, const MatchParams& sp,
                           Matches& out) {
        match_desc(ka, kb, da, db, out, sp);
        vo_match_params vp;
        if (sp.enforce_epipolar) vo_match_params_stereo(&vp, Fo); else vo_match_params_temporal(&vp);
        vector<int32_t> om(ka.size() * 3 + 3);
        int32_t on = 0;
        const vector<float> a = kp_arr(ka), b = kp_arr(kb);
        REQUIRE(vo_match_desc(a.data(), (int)ka.size(), b.data(), (int)kb.size(), da.ptr<float>(0), db.ptr<float>(0), 121,
                              &vp, om.data(), &on, nullptr, nullptr, nullptr, nullptr, nullptr) == 0);
        REQUIRE((int)out.size() == on);
        for (int i = 0; i < on; ++i) for (int j = 0; j < 3; ++j) REQUIRE(out[i][j] == om[3 * i + j]);
    };
    Matches lr0, lr1, m11, m22;
    check_match(f0.kp1, f0.kp2, f0.d1, f0.d2, MatchParams(F), lr0); // viso.cpp:1240
    check_match(f1.kp1, f1.kp2, f1.d1, f1.d2, MatchParams(F), lr1);
    check_match(f1.kp1, f0.kp1, f1.d1, f0.d1, MatchParams(), m11);  // viso.cpp:1264
    check_match(f1.kp2, f0.kp2, f1.d2, f0.d2, MatchParams(), m22);  // viso.cpp:1275
    REQUIRE(lr1.size() > 100 && m11.size() > 100);

    Mat x0, x1;
    collect_matches(f0.kp1, f0.kp2, lr0, x0);
    collect_matches(f1.kp1, f1.kp2, lr1, x1);
    Mat X0 = triangulate_rectified<double>(x0, prm); // viso.cpp:1247
    {
        const vector<float> a = kp_arr(f0.kp1), b = kp_arr(f0.kp2);
        const vector<int32_t> m = m_arr(lr0);
        vector<double> ox(4 * lr0.size()), oX(3 * lr0.size());
        vo_collect_matches(a.data(), b.data(), m.data(), (int)lr0.size(), ox.data());
        vo_triangulate_rectified_f64(ox.data(), (int)lr0.size(), prm.calib.f, prm.base, prm.calib.cu, prm.calib.cv, oX.data());
        for (size_t i = 0; i < ox.size(); ++i) REQUIRE(x0.ptr<double>(0)[i] == ox[i]);
        for (size_t i = 0; i < oX.size(); ++i) REQUIRE(X0.ptr<double>(0)[i] == oX[i]);
    }
    vector<Vec4i> circ;
    Matches pcl;
    match_circle(lr1, lr0, m11, m22, circ, pcl); // viso.cpp:1282
    {
        const vector<int32_t> a = m_arr(lr1), b = m_arr(lr0), c = m_arr(m11), d = m_arr(m22);
        vector<int32_t> oc(lr1.size() * 4 + 4), op(lr1.size() * 3 + 3);
        const int n = vo_match_circle(a.data(), (int)lr1.size(), b.data(), (int)lr0.size(), c.data(), (int)m11.size(),
                                      d.data(), (int)m22.size(), oc.data(), op.data());
        REQUIRE((int)circ.size() == n && (int)pcl.size() == n && n > 50);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < 4; ++j) REQUIRE(circ[i][j] == oc[4 * i + j]);
            REQUIRE(pcl[i][0] == op[3 * i] && pcl[i][1] == op[3 * i + 1]);
        }
    }
    // gather (viso.cpp:1291-1305) and solve
    const int C = (int)circ.size();
    Mat Xp_c(3, C, CV_64F), x_c(4, C, CV_64F);
    for (int i = 0; i < C; ++i) {
        for (int r = 0; r < 4; ++r) x_c.at<double>(r, i) = x1.at<double>(r, pcl[i][0]);
        for (int r = 0; r < 3; ++r) Xp_c.at<double>(r, i) = X0.at<double>(r, pcl[i][1]);
    }
    viso_b200::set_ransac_seed(99);
    vector<double> tr(6, 0.0);
    vector<int> inl;
    REQUIRE(ransac_minimize_reproj(Xp_c, x_c, tr, inl, prm) == true); // viso.cpp:1313
    std::printf("frame 1 motion: %g %g %g %g %g %g (%zu inliers of %d)\n", tr[0], tr[1], tr[2], tr[3], tr[4], tr[5], inl.size(), C);
    REQUIRE(std::fabs(tr[5] + 1.0) < 0.1 && std::fabs(tr[1] + 0.01) < 5e-3); // camera moved 1 m forward, yaw 0.01

    // ---- the whole loop in one batched submission vs the oracle's sequential loop, same seeds
    viso_b200::set_ransac_seed(5);
    vector<Mat> poses = sequence_odometry(P1, P2, w);
    const int nF = (int)w.size(), H = prm.ransac_iter;
    vector<int32_t> nL(nF), nR(nF);
    vector<int64_t> offL(nF), offR(nF);
    vector<float> kpL, kpR, dL, dR;
    for (int t = 0; t < nF; ++t) {
        const FrameFeatures& f = w.fr[t];
        nL[t] = (int)f.kp1.size(); nR[t] = (int)f.kp2.size();
        offL[t] = (int64_t)kpL.size() / 2; offR[t] = (int64_t)kpR.size() / 2;
        const vector<float> a = kp_arr(f.kp1), b = kp_arr(f.kp2);
        kpL.insert(kpL.end(), a.begin(), a.end()); kpR.insert(kpR.end(), b.begin(), b.end());
        dL.insert(dL.end(), f.d1.ptr<float>(0), f.d1.ptr<float>(0) + (size_t)nL[t] * 121);
        dR.insert(dR.end(), f.d2.ptr<float>(0), f.d2.ptr<float>(0) + (size_t)nR[t] * 121);
    }
    vector<uint32_t> seeds((size_t)nF * H * 3);
    {
        std::mt19937 gen(5);
        for (auto& s : seeds) s = (uint32_t)gen();
    }
    vo_param vp;
    vo_param_default(&vp);
    vp.ransac_iter = H;
    vector<vo_record> rec(nF);
    vector<double> oposes((size_t)(nF + 1) * 16);
    int32_t onp = 0;
    REQUIRE(vo_sequence(nF, nL.data(), nR.data(), offL.data(), offR.data(), kpL.data(), kpR.data(), dL.data(), dR.data(),
                        121, P1.ptr<double>(0), P2.ptr<double>(0), &vp, seeds.data(), rec.data(), nullptr, nullptr,
                        nullptr, nullptr, nullptr, nullptr, oposes.data(), &onp) == 0);
    REQUIRE((int)poses.size() == onp && onp == nF);
    for (int i = 0; i < onp; ++i)
        for (int k = 0; k < 16; ++k) REQUIRE(std::fabs(poses[i].ptr<double>(0)[k] - oposes[(size_t)16 * i + k]) < 1e-6);
    std::printf("final pose z: %g (truth ~ %g)\n", poses.back().at<double>(2, 3), 1.0 * (nF - 1));
}

// test.cpp:170-206 (solveRigidMotion) and :9-39 (triangulate_dlt), both disabled in the reference
static void test_mvg_estimation()
{
    const float pts[9] = {0, 0, 1, 0, 1, 0, 1, 0, 0}; // columns (0,0,1), (0,1,0), (1,0,0): test.cpp:175-181
    Mat X1(3, 3, CV_32F), X2(3, 3, CV_32F);
    const double T1[12] = {1, 0, 0, 1, 0, 0, -1, 2, 0, 1, 0, 3}; // R = Rx(pi/2), t = (1,2,3)
    for (int c = 0; c < 3; ++c) {
        for (int r = 0; r < 3; ++r) X1.at<float>(r, c) = pts[3 * c + r];
        for (int r = 0; r < 3; ++r)
            X2.at<float>(r, c) = (float)(T1[4 * r] * pts[3 * c] + T1[4 * r + 1] * pts[3 * c + 1] + T1[4 * r + 2] * pts[3 * c + 2] + T1[4 * r + 3]);
    }
    Mat T;
    solveRigidMotion(X2, X1, T); // test.cpp:199
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) REQUIRE(std::fabs(T.at<float>(r, c) - T1[4 * r + c]) < 1e-5);

    Mat P1 = Mat::zeros(3, 4, CV_64F), P2;
    P1.at<double>(0, 0) = 718.856; P1.at<double>(0, 2) = 607.1928; P1.at<double>(1, 1) = 718.856; P1.at<double>(1, 2) = 185.2157;
    P1.at<double>(2, 2) = 1;
    P2 = P1.clone();
    P2.at<double>(0, 3) = -386.1448;
    Mat x1(2, 1, CV_32F), x2(2, 1, CV_32F); // X = (0, 0, 1) -> test.cpp:12-33
    x1.at<float>(0, 0) = (float)P1.at<double>(0, 2); x1.at<float>(1, 0) = (float)P1.at<double>(1, 2);
    x2.at<float>(0, 0) = (float)(P1.at<double>(0, 2) - 386.1448); x2.at<float>(1, 0) = (float)P1.at<double>(1, 2);
    Mat Xd = triangulate_dlt(x1, x2, P1, P2);
    REQUIRE(std::fabs(Xd.at<float>(0, 0)) < 1e-2 && std::fabs(Xd.at<float>(1, 0)) < 1e-2 && std::fabs(Xd.at<float>(2, 0) - 1) < 1e-2);
    Mat Xr = triangulate_rectified(x1, x2, 718.856, 386.1448 / 718.856, 607.1928, 185.2157);
    REQUIRE(std::fabs(Xr.at<float>(2, 0) - 1) < 1e-3);
}


// ---- front end: HarrisBinnedFeatureDetector, MyFeatureExtractor and the image-driven sequence_odometry
// (viso.cpp:911-1025, :1167-1330) against the oracle.  Images: a band-limited random texture seen at a constant
// disparity (8 px), shifted 3 px per frame.
struct ShiftedTexture : StereoImageSource {
    int W = 1241, H = 376, n = 4, t = 0;
    vector<unsigned char> tex;   // H x (W + 64)
    ShiftedTexture()
    {
        const int TW = W + 64;
        std::mt19937 gen(21);
        std::uniform_int_distribution<> u(0, 255);
        vector<float> a((size_t)H * TW), b((size_t)H * TW);
        for (auto& v : a) v = (float)u(gen);
        for (int pass = 0; pass < 2; ++pass) {   // two 3 x 3 box blurs
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < TW; ++x) {
                    float sum = 0; int cnt = 0;
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int yy = y + dy, xx = x + dx;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= TW) continue;
                            sum += a[(size_t)yy * TW + xx]; ++cnt;
                        }
                    b[(size_t)y * TW + x] = sum / cnt;
                }
            a.swap(b);
        }
        tex.resize(a.size());
        for (size_t i = 0; i < a.size(); ++i) tex[i] = (unsigned char)std::min(255.f, std::max(0.f, (a[i] - 127.5f) * 4.f + 127.5f));
    }
    Mat view(int x0) const
    {
        Mat m(H, W, CV_8U);
        for (int y = 0; y < H; ++y) std::memcpy(m.ptr<unsigned char>(y), &tex[(size_t)y * (W + 64) + x0], W);
        return m;
    }
    image_pair pair_at(int i) const { return image_pair(view(20 + 3 * i), view(28 + 3 * i)); }
    bool next(image_pair& out) override
    {
        if (t >= n) return false;
        out = pair_at(t++);
        return true;
    }
};

static void test_front_end()
{
    Mat P1 = Mat::zeros(3, 4, CV_64F), P2;
    P1.at<double>(0, 0) = 718.856; P1.at<double>(0, 2) = 607.1928;
    P1.at<double>(1, 1) = 718.856; P1.at<double>(1, 2) = 185.2157;
    P1.at<double>(2, 2) = 1;
    P2 = P1.clone();
    P2.at<double>(0, 3) = -386.1448;
    ShiftedTexture src;
    const int W = src.W, H = src.H, NF = 1200;   // MAX_FEATURE_NUM, viso.cpp:1172

    // detector + extractor, one image (viso.cpp:1226-1231)
    HarrisBinnedFeatureDetector detector(5, NF);
    MyFeatureExtractor extractor(5);
    const image_pair p0 = src.pair_at(0);
    KeyPoints kp;
    detector.detect(p0.first, kp);
    vector<float> okp((size_t)NF * 2), oresp(NF);
    const int on = vo_detect_harris_binned(p0.first.ptr<unsigned char>(0), H, W, NF, 24, 5, .04f, 1, okp.data(), oresp.data());
    REQUIRE(on == NF && (int)kp.size() == on);
    for (int i = 0; i < on; ++i) {
        REQUIRE(kp[i].pt.x == okp[2 * i] && kp[i].pt.y == okp[2 * i + 1] && kp[i].response == oresp[i]);
        REQUIRE(kp[i].size == 11.f);
    }
    Mat d;
    extractor.compute(p0.first, kp, d);
    REQUIRE(d.rows == on && d.cols == 121 && extractor.descriptorSize() == 121);
    vector<float> sob((size_t)W * H), od((size_t)on * 121);
    vo_sobel_x(p0.first.ptr<unsigned char>(0), H, W, sob.data());
    vo_extract_descriptors(sob.data(), H, W, okp.data(), on, 5, od.data());
    REQUIRE(std::memcmp(d.ptr<float>(0), od.data(), od.size() * 4) == 0);

    // images -> poses in one submission vs the oracle's front end + sequential loop, same seeds
    viso_b200::set_ransac_seed(7);
    vector<Mat> poses = sequence_odometry(P1, P2, src);
    const int nF = src.n, Hy = 50;
    vector<int32_t> nL(nF), nR(nF);
    vector<int64_t> offL(nF), offR(nF);
    vector<float> kpL, kpR, dL, dR;
    for (int t = 0; t < nF; ++t) {
        const image_pair ip = src.pair_at(t);
        for (int side = 0; side < 2; ++side) {
            const Mat& im = side ? ip.second : ip.first;
            vector<float> k2((size_t)NF * 2);
            const int n = vo_detect_harris_binned(im.ptr<unsigned char>(0), H, W, NF, 24, 5, .04f, 1, k2.data(), nullptr);
            vector<float> dd((size_t)n * 121);
            vo_sobel_x(im.ptr<unsigned char>(0), H, W, sob.data());
            vo_extract_descriptors(sob.data(), H, W, k2.data(), n, 5, dd.data());
            vector<float>& K = side ? kpR : kpL; vector<float>& D = side ? dR : dL;
            (side ? nR : nL)[t] = n; (side ? offR : offL)[t] = (int64_t)K.size() / 2;
            K.insert(K.end(), k2.begin(), k2.begin() + 2 * n); D.insert(D.end(), dd.begin(), dd.end());
        }
    }
    vector<uint32_t> seeds((size_t)nF * Hy * 3);
    {
        std::mt19937 gen(7);
        for (auto& s : seeds) s = (uint32_t)gen();
    }
    vo_param vp;
    vo_param_default(&vp);
    vp.ransac_iter = Hy;
    vector<vo_record> rec(nF);
    vector<double> oposes((size_t)(nF + 1) * 16);
    int32_t onp = 0;
    REQUIRE(vo_sequence(nF, nL.data(), nR.data(), offL.data(), offR.data(), kpL.data(), kpR.data(), dL.data(), dR.data(),
                        121, P1.ptr<double>(0), P2.ptr<double>(0), &vp, seeds.data(), rec.data(), nullptr, nullptr,
                        nullptr, nullptr, nullptr, nullptr, oposes.data(), &onp) == 0);
    REQUIRE((int)poses.size() == onp);
    for (int i = 0; i < onp; ++i)
        for (int k = 0; k < 16; ++k) REQUIRE(std::fabs(poses[i].ptr<double>(0)[k] - oposes[(size_t)16 * i + k]) < 1e-6);
    int okc = 0;
    for (int t = 1; t < nF; ++t) okc += rec[t].ok;
    std::printf("front end: %d keypoints/image, %d poses, %d of %d frame pairs ok, n_circ %d\n", on, onp, okc, nF - 1, rec[1].n_circ);
    REQUIRE(onp >= 2 && rec[1].n_circ > 100);
}

int main()
{
    test_front_end();
    test_mvg_estimation();
    test_nl_rigid_motion1();
    test_frame_loop();
    REQUIRE(viso_b200::kernel_launches() > 0);
    std::printf("host test OK (%d checks, %lld kernel launches)\n", g_checks, viso_b200::kernel_launches());
    return 0;
}
