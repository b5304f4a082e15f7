"""Pin the CPU oracle (test infrastructure) before trusting it.

The reference ships no golden vectors and cannot be compiled here (no OpenCV / Eigen / Boost headers),
so the oracle is pinned two ways:
  * tests/golden/*.npz -- outputs of the OpenCV primitives the reference calls on the path (cvflann linear L1
    radiusSearch viso.cpp:181,684; cv::mulTransposed :1599; cv::solve LU :1602; Mat::inv :1319; cv::determinant
    mvg.h:62-64; cv::Sobel viso.cpp:1010), generated with cv2 4.13 by tools/make_golden.py.  Bit equality unless noted.
  * the known-answer recipes of the reference's own (disabled) tests: test/test.cpp:51-114 (Gauss-Newton recovers
    tr0), :170-206 (Kabsch), :9-39 (DLT), mvg.cpp:73-89 (F_from_P), plus std::sort order against the real libstdc++.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def load(name):
    return np.load(os.path.join(GOLDEN, name))


# ------------------------------------------------------------------------------------------------ cv2-pinned vectors

@pytest.mark.parametrize("case", ["int_dense", "int_sparse", "float", "dup"])
def test_radius_search_matches_cvflann(oracle, case):
    g = load("flann_radius.npz")
    pts, qs = g[case + "_pts"], g[case + "_q"]
    radius, K = float(g[case + "_radius"]), int(g[case + "_K"])
    for q, found, idx, dist in zip(qs, g[case + "_found"], g[case + "_idx"], g[case + "_dist"]):
        total, nb, d = oracle.radius_search(q, pts, radius, K)
        assert total == found
        n = min(total, K)
        assert np.array_equal(nb[:n], idx[:n])
        assert np.array_equal(d[:n], dist[:n])
        assert (nb[n:] == -1).all()  # the reference pre-fills -1 (viso.cpp:680-681)


def test_mul_transposed_and_lu_match_opencv(oracle):
    g = load("linalg.npz")
    for i in range(3):
        assert np.array_equal(oracle.mul_transposed(g[f"mt_J{i}"]), g[f"mt_JtJ{i}"])
    for A, b, x, ok in zip(g["lu_A"], g["lu_b"], g["lu_x"], g["lu_ok"]):
        ok_o, x_o = oracle.solve_lu(A, b)
        assert ok_o == bool(ok)
        if ok:
            assert np.array_equal(x_o, x)
    for A, Ai in zip(g["inv_A"], g["inv_Ai"]):
        ok, got = oracle.invert_lu(A)
        assert ok and np.array_equal(got, Ai)
    for A, det in zip(g["det_A"], g["det"]):
        assert oracle.determinant(A) == det
    assert np.array_equal(oracle.F_from_P(g["F_P1"], g["F_P2"], False), g["F_raw"])


def test_pose_chain_matches_opencv(oracle):
    """pose = pose * inv(T) (viso.cpp:1319); cv2.gemm's accumulation order is build dependent => 1e-14"""
    g = load("linalg.npz")
    pose = np.eye(4)
    for i in range(10):
        ok, Ti = oracle.invert_lu(g["inv_A"][i])
        assert ok
        pose = pose @ Ti
        assert np.abs(pose - g["pose_chain"][i]).max() < 1e-14 * max(1.0, np.abs(pose).max())


def test_descriptor_domain_from_sobel(oracle):
    """descriptors are 11x11 patches of cv::Sobel-x: integer valued, |v| <= 1020 (the kernel's u16 domain)"""
    g = load("sobel.npz")
    sob = g["sob"]
    assert np.array_equal(sob, np.rint(sob)) and np.abs(sob).max() <= 1020
    kp = np.array([[0, 0], [5, 5], [63, 47], [20.6, 30.4], [62, 3]], np.float32)
    d = oracle.extract_descriptors(sob, kp)
    assert d.shape == (5, 121)
    # viso.cpp:1013-1019: Point2i p = kp.pt (rounds), out-of-image and row/col 0 samples are 0
    x, y = 21, 30
    want = np.array([[sob[y + v, x + u] for u in range(-5, 6)] for v in range(-5, 6)], np.float32)
    got = d[3].reshape(11, 11)
    assert np.array_equal(got, want) or np.array_equal(got.T, want)
    assert (d[0] == 0).sum() >= 121 - 25  # corner: only the strictly-inside quadrant survives the '> 0' rule


# ------------------------------------------------------------------------------------------------ reference recipes

def kitti_param(oracle, H=50):
    from libviso_b200 import synth
    P1, P2 = synth.kitti_calib()
    return oracle.param_default(base=abs(P2[0, 3] / P2[0, 0]), f=P1[0, 0], cu=P1[0, 2], cv=P1[1, 2], ransac_iter=H)


def stereo_project(X, tr, oracle, p):
    T = oracle.tr2mat(tr)
    Xc = T[:3, :3] @ X + T[:3, 3:4]
    u1 = p.f * Xc[0] / Xc[2] + p.cu
    v1 = p.f * Xc[1] / Xc[2] + p.cv
    u2 = p.f * (Xc[0] - p.base) / Xc[2] + p.cu
    return np.stack([u1, v1, u2, v1])


def test_gauss_newton_recovers_motion(oracle):
    """test/test.cpp:51-114: KITTI-00 calibration, 10 random points, tr0 = (0,0,0,1,0,0), start from 0"""
    p = kitti_param(oracle)
    rng = np.random.default_rng(0)
    X = rng.random((3, 10)) * 1000
    X[2] += 5
    tr0 = np.array([0, 0, 0, 1.0, 0, 0])
    obs = stereo_project(X, tr0, oracle, p)
    ok, tr, iters = oracle.minimize_reproj(X, obs, np.zeros(6), p, np.arange(10))
    assert ok and iters < 100
    assert np.abs(tr - tr0).sum() < 1e-4


def test_compute_J_is_the_derivative_of_the_prediction(oracle):
    p = kitti_param(oracle)
    rng = np.random.default_rng(1)
    X = np.stack([rng.uniform(-20, 20, 8), rng.uniform(-2, 3, 8), rng.uniform(4, 60, 8)])
    tr = np.array([0.01, -0.02, 0.005, 0.05, -0.02, -1.0])
    obs = stereo_project(X, tr, oracle, p) + rng.normal(0, 0.3, (4, 8))
    active = np.arange(8)
    J, pred, res = oracle.compute_J(X, obs, tr, p, active)
    assert np.allclose(pred, stereo_project(X, tr, oracle, p), rtol=0, atol=1e-9)
    w = 1.0 / (np.abs(obs[0] - p.cu) / abs(p.cu) + 0.05)  # viso.cpp:1449
    assert np.allclose(res.reshape(8, 4).T, w * (obs - pred), rtol=0, atol=1e-12)
    eps = 1e-6
    for j in range(6):
        d = np.zeros(6); d[j] = eps
        num = (stereo_project(X, tr + d, oracle, p) - stereo_project(X, tr - d, oracle, p)) / (2 * eps)
        assert np.allclose(J[:, j].reshape(8, 4).T, w * num, rtol=1e-5, atol=1e-5)


def test_convergence_quirk_signed_threshold(oracle):
    """viso.cpp:1610 tests `fabs(p > thresh)`: a large NEGATIVE step counts as converged and is not applied"""
    p = kitti_param(oracle)
    rng = np.random.default_rng(2)
    X = np.stack([rng.uniform(-20, 20, 12), rng.uniform(-2, 3, 12), rng.uniform(4, 60, 12)])
    tr_true = np.array([-0.02, -0.03, -0.01, -0.5, -0.2, -1.5])  # every component negative
    obs = stereo_project(X, tr_true, oracle, p)
    ok, tr, iters = oracle.minimize_reproj(X, obs, np.zeros(6), p, np.arange(12))
    assert ok and iters == 1 and np.array_equal(tr, np.zeros(6))  # returned inside the first iteration


def test_ransac_first_best_and_min_inliers(oracle):
    from libviso_b200 import synth
    X, obs, tr_true = synth.make_ransac_problem(n=400, seed=5)
    p = kitti_param(oracle, H=64)
    table = oracle.randomsample_table(424242, 64, 400)
    r = oracle.ransac_minimize_reproj(X, obs, p, table)
    assert r["ok"] and len(r["inliers"]) >= 6
    cnt = np.where(r["hyp_ok"] == 1, r["hyp_count"], -1)
    assert r["best_hyp"] == int(np.argmax(cnt))  # argmax returns the FIRST maximum: strict '>' at viso.cpp:1564
    assert np.array_equal(r["inliers"], np.sort(r["inliers"]))
    assert np.abs(r["tr"][:3] - tr_true[:3]).max() < 5e-3
    # fewer than 6 supporters => false (viso.cpp:1571)
    p5 = kitti_param(oracle, H=8)
    r5 = oracle.ransac_minimize_reproj(X[:, :5], obs[:, :5], p5, oracle.randomsample_table(1, 8, 5))
    assert not r5["ok"]


def test_randomsample_table_is_algorithm_s(oracle):
    t = oracle.randomsample_table(424242, 4096, 10000)
    assert t.shape == (4096, 3) and t.min() >= 0 and t.max() < 10000
    assert (t[:, 0] < t[:, 1]).all() and (t[:, 1] < t[:, 2]).all()  # ascending distinct (viso.cpp:87-107)
    t3 = oracle.randomsample_table(7, 16, 3)
    assert np.array_equal(t3, np.tile([0, 1, 2], (16, 1)))


def test_solve_rigid_motion_recipe(oracle):
    """test/test.cpp:170-206: R = Rx(pi/2), t = (1,2,3); T maps B -> A"""
    X1 = np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0], [1, 1, 1]], np.float32).T
    c, s = np.cos(np.pi / 2), np.sin(np.pi / 2)
    T1 = np.array([[1, 0, 0, 1], [0, c, -s, 2], [0, s, c, 3], [0, 0, 0, 1]], np.float64)
    X2 = (T1[:3, :3] @ X1 + T1[:3, 3:4]).astype(np.float32)
    T = oracle.solve_rigid_motion(X2, X1)
    assert np.abs(T - T1).max() < 1e-5
    assert np.abs(X2 - (T[:3, :3] @ X1 + T[:3, 3:4])).max() < 1e-5


def test_triangulate_recipes(oracle):
    """test/test.cpp:9-39 (DLT: X=(0,0,1)) and both rectified variants (viso.cpp:1137-1162, mvg.cpp:172-192)"""
    from libviso_b200 import synth
    P1, P2 = synth.kitti_calib()
    X = np.array([[0.0, 1.5, -2.0], [0.0, 0.3, 1.0], [1.0, 12.0, 30.0]])
    x1 = oracle.project_points(X, P1).astype(np.float32)
    x2 = oracle.project_points(X, P2).astype(np.float32)
    Xd = oracle.triangulate_dlt(x1, x2, P1, P2)
    assert np.abs(Xd - X).max() < 1e-2
    f, cu, cv, base = P1[0, 0], P1[0, 2], P1[1, 2], abs(P2[0, 3] / P2[0, 0])
    x = np.vstack([x1, x2]).astype(np.float64)
    Xr = oracle.triangulate_rectified_f64(x, f, base, cu, cv)
    assert np.allclose(Xr, X, rtol=1e-3, atol=1e-3)
    Xf = oracle.triangulate_rectified_f32(x1, x2, f, base, cu, cv)
    assert np.allclose(Xf, X, rtol=1e-3, atol=1e-3)
    # no clamp in the pipeline version (viso.cpp:1146): zero disparity -> inf; the mvg version clamps at 1e-4
    x0 = np.array([[700.0], [200.0], [700.0], [200.0]])
    assert np.isinf(oracle.triangulate_rectified_f64(x0, f, base, cu, cv)[2, 0])
    assert np.isfinite(oracle.triangulate_rectified_f32(x0[:2].astype(np.float32), x0[2:].astype(np.float32),
                                                        f, base, cu, cv)).all()
    with pytest.raises(OverflowError):
        oracle.project_points(np.array([[1.0], [1.0], [0.0]]), P1)  # misc.h:118-119


def test_F_from_P_recipe(oracle):
    """mvg.cpp:73-89: P1=[I|0], P2=[I|(1,0,0)^T] => F = [0 0 0; 0 0 1; 0 -1 0] exactly"""
    P1 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = np.hstack([np.eye(3), np.array([[1.0], [0], [0]])])
    F = oracle.F_from_P(P1, P2, False)
    assert np.array_equal(F, np.array([[0, 0, 0], [0, 0, 1.0], [0, -1.0, 0]]))


def test_sampson_gate_is_one_pixel_for_rectified_kitti(oracle):
    """with the rectified KITTI F, Sampson <= 1 <=> |dy| <= 1 px (0, 0.5, 2.0 for dy = 0, 1, 2)"""
    from libviso_b200 import synth
    F = oracle.F_from_P(*synth.kitti_calib())
    for dy, want in ((0, 0.0), (1, 0.5), (2, 2.0)):
        assert abs(oracle.sampson_distance(F, (300, 100), (280, 100 + dy)) - want) < 1e-6


def test_sort_matches_is_libstdcxx_std_sort(oracle, tmp_path):
    """viso.cpp:724 uses std::sort, whose order of equal distances is a property of libstdc++'s introsort; the oracle's
    restatement (and the device copy of it) must give the same permutation as the real thing"""
    import ctypes as C
    src = os.path.join(ROOT, "tests", "introsort_check.cpp")
    so = str(tmp_path / "introsort_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so])
    chk = C.CDLL(so)
    rng = np.random.default_rng(9)
    for n, hi in ((0, 5), (1, 5), (16, 3), (17, 3), (700, 40), (2000, 100000), (5000, 7), (3000, 1), (40000, 2000)):
        m = np.ascontiguousarray(
            np.stack([np.arange(n), rng.integers(0, 10 ** 6, n), rng.integers(0, hi, n)], 1).astype(np.int32))
        restated = np.zeros_like(m)
        same = chk.introsort_matches_std(m.ctypes.data_as(C.c_void_p), n, restated.ctypes.data_as(C.c_void_p))
        assert same == 1, (n, hi)                      # device header == real std::sort
        assert np.array_equal(oracle.sort_matches(m), restated), (n, hi)  # oracle == both
    # adversarial: organ-pipe and sorted inputs exercise the depth limit / heapsort branch
    for m3 in (np.r_[np.arange(3000), np.arange(3000)[::-1]], np.arange(5000), np.arange(5000)[::-1] // 3):
        n = len(m3)
        m = np.ascontiguousarray(np.stack([np.arange(n), np.arange(n), m3], 1).astype(np.int32))
        restated = np.zeros_like(m)
        assert chk.introsort_matches_std(m.ctypes.data_as(C.c_void_p), n, restated.ctypes.data_as(C.c_void_p)) == 1
        assert np.array_equal(oracle.sort_matches(m), restated)


# ------------------------------------------------------------------------------------------------ whole path

def test_sequence_recovers_ground_truth_motion(oracle, small_sequence):
    """BASELINE configs[0] in miniature: synthetic KITTI-layout sequence through the whole CPU path"""
    from conftest import make_seeds
    from libviso_b200 import synth
    frames, gt = small_sequence
    P1, P2 = synth.kitti_calib()
    seeds = make_seeds(len(frames), 50)
    out = oracle.sequence(frames, P1, P2, oracle.param_default(ransac_iter=50), seeds, dump=True)
    rec = out["records"]
    assert rec["ok"][1:].all() and (rec["n_circ"][1:] > 50).all()
    for t in range(1, len(frames)):
        assert np.abs(oracle.tr2mat(rec["tr"][t]) - gt[t]).max() < 0.05
        lr = out["lr_matches"][t]
        assert (np.diff(lr[:, 2]) >= 0).all() and len(np.unique(lr[:, 0])) == len(lr)
    assert len(out["poses"]) == len(frames)


# ------------------------------------------------------------------------------------------------ cv2 transcription

def _dense_matches(o, g):
    want = np.stack([o["idx"], o["d1"], o["d2"], o["valid"]], 1).astype(np.int64)
    assert np.array_equal(want, g)


def test_path_against_cv2_transcription(oracle):
    """tests/golden/path_cv2.npz (tools/make_golden_path.py): the reference's control flow transcribed literally with the
    REAL OpenCV calls at every third-party call site (cvflann radiusSearch, cv::norm, mulTransposed, gemm, solve,
    invert).  Integer results must be identical, tr / poses agree to 1e-9 (gemm's accumulation order)."""
    g = load("path_cv2.npz")
    F, P1, P2 = g["F"], g["P1"], g["P2"]
    assert np.array_equal(oracle.F_from_P(P1, P2), F)
    stereo, temporal = oracle.match_params_stereo(F), oracle.match_params_temporal()
    p = oracle.param_default(base=abs(P2[0, 3] / P2[0, 0]), f=P1[0, 0], cu=P1[0, 2], cv=P1[1, 2], ransac_iter=12)
    fr = [{k: g[f"f{t}_{k}"] for k in ("kpL", "kpR", "dL", "dR")} for t in range(3)]
    lr, xs, Xs = [], [], []
    for t, f in enumerate(fr):
        o = oracle.match_desc(f["kpL"], f["kpR"], f["dL"], f["dR"], stereo)
        _dense_matches(o, g[f"lr{t}_dense"])
        assert np.array_equal(o["matches"], oracle.sort_matches(g[f"lr{t}_push"]))   # push order -> std::sort order
        lr.append(o["matches"])
        x = oracle.collect_matches(f["kpL"], f["kpR"], o["matches"])
        X = oracle.triangulate_rectified_f64(x, p.f, p.base, p.cu, p.cv)
        assert np.array_equal(x, g[f"x{t}"]) and np.array_equal(X, g[f"X{t}"])
        xs.append(x); Xs.append(X)
    pose = np.eye(4)
    n_pose = 1
    for t in (1, 2):
        f, fp = fr[t], fr[t - 1]
        o11 = oracle.match_desc(f["kpL"], fp["kpL"], f["dL"], fp["dL"], temporal)
        o22 = oracle.match_desc(f["kpR"], fp["kpR"], f["dR"], fp["dR"], temporal)
        _dense_matches(o11, g[f"m11_{t}_dense"])
        _dense_matches(o22, g[f"m22_{t}_dense"])
        circ, pcl = oracle.match_circle(lr[t], lr[t - 1], o11["matches"], o22["matches"])
        assert np.array_equal(circ, g[f"circ{t}"]) and np.array_equal(pcl[:, :2], g[f"pcl{t}"])
        assert len(circ) >= 20
        x_c = np.ascontiguousarray(xs[t][:, pcl[:, 0]]); Xp_c = np.ascontiguousarray(Xs[t - 1][:, pcl[:, 1]])
        table = oracle.samples_from_seeds(g["seeds"][t], len(circ))
        assert np.array_equal(table, g[f"table{t}"])
        r = oracle.ransac_minimize_reproj(Xp_c, x_c, p, table)
        assert r["ok"] == bool(g[f"ok{t}"])
        assert np.array_equal(r["hyp_ok"], g[f"hok{t}"]) and np.array_equal(r["hyp_count"], g[f"hcnt{t}"])
        assert np.array_equal(r["inliers"], g[f"inl{t}"])
        assert np.abs(r["tr"] - g[f"tr{t}"]).max() < 1e-9
        if r["ok"]:
            ok, pose = oracle.pose_update(pose, r["tr"])
            assert ok and np.abs(pose - g["poses"][n_pose]).max() < 1e-9
            n_pose += 1
    assert n_pose == len(g["poses"])
    # stress cases: truncation (found > K), the mono configuration (general F + ratio test), index-0 terminator, ties
    def sp(enforce_epipolar, Fm, sampson, second, ratio, K, radius):
        m = oracle.match_params_temporal()
        m.enforce_epipolar, m.enforce_2nd_best, m.max_neighbors = int(enforce_epipolar), int(second), K
        m.radius, m.sampson_thresh, m.ratio_2nd_best = radius, sampson, ratio
        if Fm is not None:
            for i, v in enumerate(np.asarray(Fm).reshape(9)):
                m.F[i] = v
        return m
    cases = {"trunc": sp(False, None, 0.0, True, .9, 16, 80.0),
             "mono": sp(True, g["s_Fg"], 400.0, True, .9, 250, 25.0),
             "plain": sp(False, None, 0.0, False, .9, 250, 80.0)}
    for name, m in cases.items():
        o = oracle.match_desc(g["s_kpa"], g["s_kpb"], g["s_da"], g["s_db"], m)
        _dense_matches(o, g[f"s_{name}_dense"])
        assert np.array_equal(o["matches"], oracle.sort_matches(g[f"s_{name}_push"]))
        assert (o["idx"] >= 0).sum() > 10, name   # (random descriptors: the ratio test rejects most)


# ------------------------------------------------------------------------------------------------ front end (8f rank 1-2)

def _harris_numpy(img, k=np.float32(0.04)):
    """the canonical float32 evaluation of oracle/viso_oracle.h written with numpy (IEEE, no FMA)"""
    def refl(i, n):
        i = np.where(i < 0, -i, i)
        return np.where(i >= n, 2 * (n - 1) - i, i)
    h, w = img.shape
    p = img.astype(np.float32)
    xi = lambda d: refl(np.arange(w) + d, w)
    yi = lambda d: refl(np.arange(h) + d, h)
    s = 1.0 / (16 * 3 * 255.0)
    f0, f1, f2 = np.float32(6.0 * s), np.float32(4.0 * s), np.float32(1.0 * s)
    r = (p[:, xi(2)] - p[:, xi(-2)]) + np.float32(2) * (p[:, xi(1)] - p[:, xi(-1)])
    dx = f0 * r; dx = dx + f1 * (r[yi(-1)] + r[yi(1)]); dx = dx + f2 * (r[yi(-2)] + r[yi(2)])
    t = f0 * p; t = t + f1 * (p[:, xi(-1)] + p[:, xi(1)]); t = t + f2 * (p[:, xi(-2)] + p[:, xi(2)])
    dy = np.float32(2) * (t[yi(1)] - t[yi(-1)]); dy = dy + (t[yi(2)] - t[yi(-2)])
    def box(c):
        rs = (c[:, xi(-1)] + c) + c[:, xi(1)]
        return (rs[yi(-1)] + rs) + rs[yi(1)]
    a, b, c = box(dx * dx), box(dx * dy), box(dy * dy)
    tr = a + c
    return (a * c - b * b) - (k * tr) * tr


@pytest.mark.parametrize("case", ["crop", "noise"])
def test_harris_response_and_detector_against_opencv(oracle, case):
    """tests/golden/harris.npz: cv2.cornerHarris (not bit-reproducible, see oracle/viso_oracle.h) within 2e-6 of the
    image maximum; the canonical evaluation bit-exactly against its numpy statement; the binned detector run on the
    oracle's response picks the keypoints cv2's response picks"""
    g = load("harris.npz")
    img, ref = g[case + "_img"], g[case + "_resp"]
    nx, ny, n = (int(v) for v in g[case + "_bins"])
    r = oracle.harris_response(img, 0.04)
    assert np.abs(r - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.array_equal(r, _harris_numpy(img))
    kp, resp = oracle.detect_harris_binned(img, n, nx, ny, 0.04, order_rule=1, with_response=True)
    assert len(kp) == n
    assert np.array_equal(resp, np.abs(r)[kp[:, 1].astype(int), kp[:, 0].astype(int)])
    want = set(map(tuple, g[case + "_kp"]))
    got = set(map(tuple, kp))
    assert len(got & want) >= 0.95 * len(want)
    # the reference's literal std::nth_element (order_rule 0) keeps the same set; only the order inside a bin differs
    kp0 = oracle.detect_harris_binned(img, n, nx, ny, 0.04, order_rule=0)
    assert set(map(tuple, kp0)) == got
    per = n // (nx * ny)
    for b in range(nx * ny):
        blk, blk0 = kp[b * per:(b + 1) * per], kp0[b * per:(b + 1) * per]
        assert set(map(tuple, blk)) == set(map(tuple, blk0))
        sx, sy = img.shape[1] // nx, img.shape[0] // ny
        assert (blk[:, 0] // sx == b // ny).all() and (blk[:, 1] // sy == b % ny).all()   # binx outer, biny inner
        assert (np.diff(resp[b * per:(b + 1) * per]) >= 0).all()                          # ascending inside a bin


def test_harris_detector_edge_cases(oracle):
    """flat image: every response is 0 and is skipped (viso.cpp:956); bins with fewer candidates than the quota"""
    flat = np.full((40, 48), 100, np.uint8)
    assert len(oracle.detect_harris_binned(flat, 24, 4, 2)) == 0
    img = flat.copy()
    img[10:14, 5:9] = 255           # one corner blob in the first bin
    kp = oracle.detect_harris_binned(img, 8 * 400, 4, 2)     # quota 400 per bin, far more than non-zero responses
    assert 0 < len(kp) < 8 * 400 and (kp[:, 0] < 24).all() and (kp[:, 1] < 24).all()
    r = np.abs(oracle.harris_response(img))
    assert len(kp) == int((r[:40, :48] != 0).sum())


def test_sobel_x_matches_opencv(oracle):
    g = load("sobel.npz")
    assert np.array_equal(oracle.sobel_x(g["img"]), g["sob"])
