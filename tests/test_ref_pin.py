"""The oracle pinned against the REFERENCE'S OWN code.

oracle/_ref/libviso_ref.so holds the reference's hot-path functions compiled UNCHANGED from /root/reference/src (the
live line ranges of viso.cpp extracted at build time, mvg.cpp / misc.cpp / estimation.cpp as they lie: oracle/Makefile
target `ref`) against header stand-ins for OpenCV / Boost / Eigen (compat/, oracle/shim/).  These tests run the
restated oracle (oracle/viso_oracle.cpp) and that library on the same inputs and require identical outputs -- every
match, distance, order, circular match, Jacobian entry, inlier, step of Gauss-Newton.  The third-party routines inside
the shim are themselves pinned to OpenCV 4.13 by the golden vectors (the `shim_*` tests below).

CPU only; nothing here reads /root/reference at run time (the library is built by __graft_entry__.build()).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, make_seeds

from oracle import ref as refmod

pytestmark = pytest.mark.skipif(not refmod.available(), reason="oracle/_ref/libviso_ref.so is not built and the reference tree is absent")


@pytest.fixture(scope="module")
def ref():
    refmod.lib()
    return refmod


def random_features(rng, n, w=1241, h=376, dlen=121, integer=True):
    if integer:
        flat = rng.choice(w * h, n, replace=False)
        kp = np.stack([flat % w, flat // w], 1).astype(np.float32)
    else:
        kp = (rng.random((n, 2)) * [w, h]).astype(np.float32)
    d = rng.integers(-1020, 1021, size=(n, dlen)).astype(np.float32)
    return kp, d


# ------------------------------------------------------------------------------------------------ match_desc

def test_match_desc_synthetic_frames(oracle, ref, small_sequence):
    """viso.cpp:668-726 on the config-1 style synthetic frames: stereo (epipolar gate) and temporal (ratio test)"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    F = oracle.F_from_P(*synth.kitti_calib())
    for t in (1, 2, 4):
        f, fp = frames[t], frames[t - 1]
        s = oracle.match_params_stereo(F)
        o = oracle.match_desc(f["kpL"], f["kpR"], f["dL"], f["dR"], s)["matches"]
        assert np.array_equal(o, ref.match_desc(f["kpL"], f["kpR"], f["dL"], f["dR"], s)) and len(o) > 100
        s = oracle.match_params_temporal()
        for a, b in ((("kpL", "dL"), ("kpL", "dL")), (("kpR", "dR"), ("kpR", "dR"))):
            o = oracle.match_desc(f[a[0]], fp[b[0]], f[a[1]], fp[b[1]], s)["matches"]
            assert np.array_equal(o, ref.match_desc(f[a[0]], fp[b[0]], f[a[1]], fp[b[1]], s)) and len(o) > 50


@pytest.mark.parametrize("n,w,h,K,radius,integer", [
    (1500, 400, 300, 250, 80.0, True),    # the top-max_neighbors truncation binds (viso.cpp:180-186), many L1 ties
    (1500, 400, 300, 16, 80.0, True),
    (1200, 640, 300, 200, 80.0, False),   # float coordinates
    (500, 1241, 376, 5, 300.0, True),     # radius beyond the image height
    (300, 200, 120, 250, 0.0, True),      # radius 0
    (400, 300, 200, 1, 40.0, True),
])
def test_match_desc_random(oracle, ref, n, w, h, K, radius, integer):
    rng = np.random.default_rng(n * 3 + K)
    kp1, d1 = random_features(rng, n, w, h, integer=integer)
    kp2, d2 = random_features(rng, n + 11, w, h, integer=integer)
    if radius == 0.0:
        kp2[5:150] = kp1[5:150]
    d2[rng.integers(0, len(d2), 300)] = d2[rng.integers(0, len(d2), 300)]  # SAD ties
    for second in (0, 1):
        s = oracle.match_params_temporal()
        s.max_neighbors = K; s.radius = radius; s.enforce_2nd_best = second
        o = oracle.match_desc(kp1, kp2, d1, d2, s)["matches"]
        assert np.array_equal(o, ref.match_desc(kp1, kp2, d1, d2, s))


def test_match_desc_quirks(oracle, ref):
    """index 0 ends the scan (viso.cpp:693), coincident keypoints, identical descriptors (ties -> last scanned)"""
    rng = np.random.default_rng(5)
    kp1, d1 = random_features(rng, 900, 300, 200)
    base, d = random_features(rng, 150, 300, 200)
    kp2 = np.repeat(base, 6, axis=0); d2 = np.repeat(d, 6, axis=0)
    kp2[0] = (150, 100)
    d1[:50] = d[:50]
    for K in (7, 64, 250):
        s = oracle.match_params_temporal()
        s.max_neighbors = K
        o = oracle.match_desc(kp1, kp2, d1, d2, s)
        assert (o["idx"] == 0).sum() == 0
        assert np.array_equal(o["matches"], ref.match_desc(kp1, kp2, d1, d2, s))


def test_match_desc_general_F_and_mono_configuration(oracle, ref):
    rng = np.random.default_rng(8)
    kp1, d1 = random_features(rng, 800, 640, 300)
    kp2, d2 = random_features(rng, 800, 640, 300)
    F = rng.standard_normal((3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-2], [1e-3, 1e-2, 1.0]])
    s = oracle.match_params_stereo(F)
    s.sampson_thresh = 4.0
    o = oracle.match_desc(kp1, kp2, d1, d2, s)["matches"]
    assert np.array_equal(o, ref.match_desc(kp1, kp2, d1, d2, s)) and len(o) > 20
    s.radius = 10; s.enforce_2nd_best = 1; s.ratio_2nd_best = 0.9   # calibratedSFM's parameters, viso.cpp:1364-1367
    assert np.array_equal(oracle.match_desc(kp1, kp2, d1, d2, s)["matches"], ref.match_desc(kp1, kp2, d1, d2, s))
    for _ in range(200):
        a, b = rng.random(2) * [640, 300], rng.random(2) * [640, 300]
        x, y = oracle.sampson_distance(F, a, b), ref.sampson_distance(F, a, b)
        assert x == y or (np.isnan(x) and np.isnan(y))


def test_golden_path_inputs(oracle, ref):
    """the stress inputs of tests/golden/path_cv2.npz (the cv2-executed transcription) through the reference's own code"""
    g = np.load(os.path.join(GOLDEN, "path_cv2.npz"))
    stereo, temporal = oracle.match_params_stereo(g["F"]), oracle.match_params_temporal()
    for t in range(3):
        m = ref.match_desc(g[f"f{t}_kpL"], g[f"f{t}_kpR"], g[f"f{t}_dL"], g[f"f{t}_dR"], stereo)
        assert np.array_equal(m, oracle.sort_matches(g[f"lr{t}_push"]))
    def sp(enforce_epipolar, Fm, sampson, second, ratio, K, radius):
        m = oracle.match_params_temporal()
        m.enforce_epipolar, m.enforce_2nd_best, m.max_neighbors = int(enforce_epipolar), int(second), K
        m.radius, m.sampson_thresh, m.ratio_2nd_best = radius, sampson, ratio
        if Fm is not None:
            for i, v in enumerate(np.asarray(Fm).reshape(9)):
                m.F[i] = v
        return m
    # truncation (found > K), the mono configuration (general F + ratio test), no ratio test: tools/make_golden_path.py
    cases = {"trunc": sp(False, None, 0.0, True, .9, 16, 80.0), "mono": sp(True, g["s_Fg"], 400.0, True, .9, 250, 25.0),
             "plain": sp(False, None, 0.0, False, .9, 250, 80.0)}
    for name, m in cases.items():
        got = ref.match_desc(g["s_kpa"], g["s_kpb"], g["s_da"], g["s_db"], m)
        assert np.array_equal(got, oracle.sort_matches(g[f"s_{name}_push"])), name


# ------------------------------------------------------------------------------------------------ circle, geometry

def pipeline_state(oracle, frames, P1, P2):
    from libviso_b200 import synth
    F = oracle.F_from_P(P1, P2)
    stereo, temporal = oracle.match_params_stereo(F), oracle.match_params_temporal()
    lr = [oracle.match_desc(f["kpL"], f["kpR"], f["dL"], f["dR"], stereo)["matches"] for f in frames]
    out = []
    for t in range(1, len(frames)):
        f, fp = frames[t], frames[t - 1]
        m11 = oracle.match_desc(f["kpL"], fp["kpL"], f["dL"], fp["dL"], temporal)["matches"]
        m22 = oracle.match_desc(f["kpR"], fp["kpR"], f["dR"], fp["dR"], temporal)["matches"]
        out.append((lr[t], lr[t - 1], m11, m22))
    return lr, out


def test_match_circle_collect_triangulate(oracle, ref, small_sequence):
    """viso.cpp:206-243, 501-514, 1137-1162"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    lr, quads = pipeline_state(oracle, frames[:4], P1, P2)
    for (a, b, c, d) in quads:
        co, po = oracle.match_circle(a, b, c, d)
        cr, pr = ref.match_circle(a, b, c, d)
        assert np.array_equal(co, cr) and np.array_equal(po[:, :2], pr[:, :2]) and len(co) > 20
    f = frames[1]
    xo = oracle.collect_matches(f["kpL"], f["kpR"], lr[1])
    Xo = oracle.triangulate_rectified_f64(xo, synth.F_PX, synth.BASE, synth.CU, synth.CV)
    xr, Xr = ref.collect_triangulate(f["kpL"], f["kpR"], lr[1], synth.F_PX, synth.BASE, synth.CU, synth.CV)
    assert np.array_equal(xo, xr) and np.array_equal(Xo, Xr, equal_nan=True)
    # zero and negative disparities propagate as IEEE inf / negative depth (no clamp, viso.cpp:1147-1151)
    kp = np.array([[10, 5], [20, 7], [30, 9]], np.float32)
    kq = np.array([[10, 5], [25, 7], [29, 9]], np.float32)
    m = np.array([[0, 0, 1], [1, 1, 2], [2, 2, 3]], np.int32)
    xo = oracle.collect_matches(kp, kq, m)
    with np.errstate(all="ignore"):
        Xo = oracle.triangulate_rectified_f64(xo, 700.0, 0.5, 600.0, 180.0)
    xr, Xr = ref.collect_triangulate(kp, kq, m, 700.0, 0.5, 600.0, 180.0)
    assert np.array_equal(Xo, Xr, equal_nan=True)


def test_tr2mat_and_F_from_P(oracle, ref):
    """viso.cpp:109-133; mvg.h:41-66 + the normalisation of viso.cpp:1176-1180"""
    from libviso_b200 import synth
    rng = np.random.default_rng(3)
    for _ in range(50):
        tr = rng.standard_normal(6) * [0.3, 0.3, 0.3, 2, 2, 2]
        assert np.array_equal(oracle.tr2mat(tr), ref.tr2mat(tr))
    P1, P2 = synth.kitti_calib()
    assert np.array_equal(oracle.F_from_P(P1, P2), ref.F_from_P(P1, P2))
    assert np.array_equal(oracle.F_from_P(P1, P2, False), ref.F_from_P(P1, P2, False))
    for _ in range(20):
        A, B = rng.standard_normal((3, 4)) * 100, rng.standard_normal((3, 4)) * 100
        assert np.array_equal(oracle.F_from_P(A, B), ref.F_from_P(A, B))


# ------------------------------------------------------------------------------------------------ estimation

def test_compute_J_get_inliers_minimize_reproj(oracle, ref):
    """viso.cpp:1401-1497, 1509-1537, 1583-1623: bit for bit, including the weight-column and signed-convergence quirks"""
    from libviso_b200 import synth
    X, obs, tr_true = synth.make_ransac_problem(n=600, seed=11)
    p = oracle.param_default(base=synth.BASE, f=synth.F_PX, cu=synth.CU, cv=synth.CV)
    rng = np.random.default_rng(2)
    for trial in range(6):
        tr = tr_true + rng.standard_normal(6) * 0.01 * trial
        active = np.sort(rng.choice(600, 3 if trial % 2 else 40, replace=False)).astype(np.int32)
        Jo, po, ro = oracle.compute_J(X, obs, tr, p, active)
        Jr, pr, rr = ref.compute_J(X, obs, tr, p, active)
        assert np.array_equal(Jo, Jr) and np.array_equal(po, pr) and np.array_equal(np.ravel(ro), rr)
        assert np.array_equal(oracle.get_inliers(X, obs, tr, p)[0], ref.get_inliers(X, obs, tr, p))
        oko, tro, _ = oracle.minimize_reproj(X, obs, np.zeros(6), p, active)
        okr, trr = ref.minimize_reproj(X, obs, np.zeros(6), p, active)
        assert oko == okr
        assert np.array_equal(tro, trr)
    # degenerate sample (three identical points): singular normal equations -> false (viso.cpp:1603-1606)
    Xd = np.repeat(X[:, :1], 3, axis=1); od = np.repeat(obs[:, :1], 3, axis=1)
    okr, _ = ref.minimize_reproj(Xd, od, np.zeros(6), p, np.arange(3, dtype=np.int32))
    assert okr == oracle.minimize_reproj(Xd, od, np.zeros(6), p, np.arange(3, dtype=np.int32))[0]


@pytest.mark.parametrize("n,H", [(300, 50), (2000, 300)])
def test_ransac_minimize_reproj(oracle, ref, n, H):
    """viso.cpp:1543-1580 with the same host-supplied sample table: identical inlier set, tr to 1e-9 (the oracle sums
    J^T r in row order, cv::gemm's order is build dependent; the shim's product is row order as well)"""
    from libviso_b200 import synth
    X, obs, _ = synth.make_ransac_problem(n=n, seed=n)
    p = oracle.param_default(base=synth.BASE, f=synth.F_PX, cu=synth.CU, cv=synth.CV, ransac_iter=H)
    table = oracle.randomsample_table(7, H, n)
    o = oracle.ransac_minimize_reproj(X, obs, p, table)
    okr, trr, inlr = ref.ransac_minimize_reproj(X, obs, p, table)
    assert bool(o["ok"]) == okr and okr
    assert np.array_equal(o["inliers"], inlr)
    assert np.array_equal(np.asarray(o["tr"]), trr)
    # fewer than 6 inliers -> false, best_tr as left by the loop (viso.cpp:1571-1573)
    Xb, ob = X[:, :5].copy(), obs[:, :5].copy()
    tb = oracle.randomsample_table(1, 20, 5)
    p.ransac_iter = 20
    ob2 = oracle.ransac_minimize_reproj(Xb, ob, p, tb)
    okr, trr, inlr = ref.ransac_minimize_reproj(Xb, ob, p, tb)
    assert bool(ob2["ok"]) == okr and not okr


# ------------------------------------------------------------------------------------------------ front end + sequence

def test_front_end(oracle, ref, small_sequence):
    """HarrisBinnedFeatureDetector (viso.cpp:911-979; the reference's std::nth_element order = the oracle's order_rule 0,
    on the shared canonical cornerHarris) and MyFeatureExtractor (viso.cpp:981-1025)"""
    frames, _ = small_sequence
    img = frames[0]["imL"]
    for n in (1200, 600, 2040):
        ko, ro = oracle.detect_harris_binned(img, n, order_rule=0, with_response=True)
        kr, rr = ref.detect(img, n)
        assert np.array_equal(ko, kr) and np.array_equal(ro, rr) and len(ko) > n // 2
    H, W = img.shape
    extra = np.array([[0, 0], [1, 1], [W - 1, H - 1], [W - 2, 3], [5, H - 1], [0.5, 1.5], [2.5, 3.5], [100.49, 50.51]], np.float32)
    kp = np.concatenate([ko, extra])
    assert np.array_equal(oracle.extract_descriptors(oracle.sobel_x(img), kp), ref.extract(img, kp))


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


def test_sequence_odometry_from_image_files(oracle, ref, small_sequence, tmp_path):
    """the reference's sequence_odometry (viso.cpp:1167-1330: MAX_FEATURE_NUM 1200, 50 hypotheses) on image files against
    the oracle's pipeline fed with the oracle's front end: same poses.  Images are written as PGM under the names the
    kitti driver uses (image_0/%06d.png): the stand-in cv::imread decodes by content."""
    from libviso_b200 import synth
    frames, _ = small_sequence
    frames = frames[:4]
    P1, P2 = synth.kitti_calib()
    os.makedirs(tmp_path / "image_0"); os.makedirs(tmp_path / "image_1")
    for t, f in enumerate(frames):
        write_pgm(tmp_path / "image_0" / ("%06d.png" % t), f["imL"])
        write_pgm(tmp_path / "image_1" / ("%06d.png" % t), f["imR"])
    seeds1 = make_seeds(1, 50)[0]
    poses = ref.sequence_odometry(P1, P2, str(tmp_path / "image_0" / "%06d.png"), str(tmp_path / "image_1" / "%06d.png"), 0, 10 ** 6, seeds1)
    ofr = []
    for f in frames:
        kl = oracle.detect_harris_binned(f["imL"], 1200, order_rule=0); kr = oracle.detect_harris_binned(f["imR"], 1200, order_rule=0)
        ofr.append(dict(kpL=kl, kpR=kr, dL=oracle.extract_descriptors(oracle.sobel_x(f["imL"]), kl),
                        dR=oracle.extract_descriptors(oracle.sobel_x(f["imR"]), kr)))
    o = oracle.sequence(ofr, P1, P2, oracle.param_default(ransac_iter=50), np.repeat(seeds1[None], len(frames), 0))
    assert len(poses) == len(o["poses"]) == len(frames)
    assert np.abs(poses - o["poses"]).max() < 1e-9


def test_pipeline_on_the_config1_sequence(oracle, ref, small_sequence):
    """BASELINE configs[0]-style sequence (synthetic KITTI-shaped frames, given features, 50 hypotheses per frame pair):
    the reference's own functions in the reference's loop order (ref_abi.inc vr_pipeline) against the oracle's
    vo_sequence -- per frame pair the same ok / n_inliers / n_circ, tr and chained poses"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    H = 50
    seeds = make_seeds(len(frames), H)
    o = oracle.sequence(frames, P1, P2, oracle.param_default(ransac_iter=H), seeds)
    r = ref.pipeline(frames, P1, P2, H, seeds)
    for k in ("ok", "n_inliers", "n_circ"):
        assert np.array_equal(o["records"][k], r["records"][k]), k
    assert o["records"]["ok"][1:].all() and (o["records"]["n_circ"][1:] > 50).all()
    assert np.array_equal(o["records"]["tr"], r["records"]["tr"])
    assert len(o["poses"]) == len(r["poses"]) and np.abs(o["poses"] - r["poses"]).max() < 1e-12
    # degenerate frames: no features at all in one frame -> that pair and the next are skipped (n_circ < 3), like :1283-1288
    empty = dict(kpL=np.zeros((0, 2), np.float32), kpR=np.zeros((0, 2), np.float32), dL=np.zeros((0, 121), np.float32),
                 dR=np.zeros((0, 121), np.float32))
    fr2 = [frames[0], frames[1], empty, frames[2], frames[3]]
    seeds2 = make_seeds(len(fr2), H)
    o = oracle.sequence(fr2, P1, P2, oracle.param_default(ransac_iter=H), seeds2)
    r = ref.pipeline(fr2, P1, P2, H, seeds2)
    for k in ("ok", "n_inliers", "n_circ"):
        assert np.array_equal(o["records"][k], r["records"][k]), k
    assert list(o["records"]["ok"]) == [0, 1, 0, 0, 1]


# ------------------------------------------------------------------------------------------------ mvg / estimation.cpp

def test_mvg_and_rigid_motion(oracle, ref):
    """mvg.cpp:124-192 and estimation.cpp:29-51 compiled as they lie (SVD-based: tolerance)"""
    from libviso_b200 import synth
    rng = np.random.default_rng(4)
    P1, P2 = synth.kitti_calib()
    X = np.stack([rng.uniform(-10, 10, 40), rng.uniform(-2, 2, 40), rng.uniform(5, 50, 40)])
    Xh = np.vstack([X, np.ones(40)])
    x1 = (P1 @ Xh); x1 = (x1[:2] / x1[2]).astype(np.float32)
    x2 = (P2 @ Xh); x2 = (x2[:2] / x2[2]).astype(np.float32)
    a, b = oracle.triangulate_dlt(x1, x2, P1, P2), ref.triangulate_dlt(x1, x2, P1, P2)
    assert np.abs(a - b).max() <= 2e-4 * np.abs(b).max()
    a = oracle.triangulate_rectified_f32(x1, x2, synth.F_PX, synth.BASE, synth.CU, synth.CV)
    assert np.array_equal(a, ref.triangulate_rectified_f32(x1, x2, synth.F_PX, synth.BASE, synth.CU, synth.CV))
    A = rng.standard_normal((3, 12)).astype(np.float32)
    c, s = np.cos(0.4), np.sin(0.4)
    R = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float32)
    B = (R.T @ (A - np.array([[1], [2], [3]], np.float32))).astype(np.float32)
    To, Tr = oracle.solve_rigid_motion(A, B), ref.solve_rigid_motion(A, B)
    assert np.abs(To - Tr).max() < 1e-4


# ------------------------------------------------------------------------------------------------ the shim's own pins

@pytest.mark.parametrize("case", ["int_dense", "int_sparse", "float", "dup"])
def test_shim_radius_search_matches_cvflann(ref, case):
    """oracle/shim/opencv2/flann/flann.hpp against OpenCV 4.13's cvflann linear index (tests/golden/flann_radius.npz)"""
    g = np.load(os.path.join(GOLDEN, "flann_radius.npz"))
    pts, qs = g[case + "_pts"], g[case + "_q"]
    radius, K = float(g[case + "_radius"]), int(g[case + "_K"])
    for q, found, idx, dist in zip(qs, g[case + "_found"], g[case + "_idx"], g[case + "_dist"]):
        total, nb, d = ref.shim_radius_search(q, pts, radius, K)
        n = min(total, K)
        assert total == found and np.array_equal(nb[:n], idx[:n]) and np.array_equal(d[:n], dist[:n]) and (nb[n:] == -1).all()


def test_shim_linear_algebra_matches_opencv(ref):
    """cv::mulTransposed, cv::solve(DECOMP_LU), Mat::inv, cv::determinant of the stand-ins against OpenCV 4.13"""
    g = np.load(os.path.join(GOLDEN, "linalg.npz"))
    for i in range(3):
        assert np.array_equal(ref.shim_mul_transposed(g[f"mt_J{i}"]), g[f"mt_JtJ{i}"])
    for A, b, x, ok in zip(g["lu_A"], g["lu_b"], g["lu_x"], g["lu_ok"]):
        ok_s, x_s = ref.shim_solve(A, b)
        assert ok_s == bool(ok)
        if ok:
            assert np.array_equal(x_s, x)
    for A, Ai in zip(g["inv_A"], g["inv_Ai"]):
        ok, got = ref.shim_invert(A)
        assert ok and np.array_equal(got, Ai)
    for A, det in zip(g["det_A"], g["det"]):
        assert ref.shim_determinant(A) == det
    assert np.array_equal(ref.F_from_P(g["F_P1"], g["F_P2"], False), g["F_raw"])


def test_shim_sobel_matches_opencv(ref):
    g = np.load(os.path.join(GOLDEN, "sobel.npz"))
    assert np.array_equal(ref.shim_sobel(g["img"]), g["sob"])
