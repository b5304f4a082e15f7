"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Integer / index work (match indices, SAD distances, match order, circular matches, inlier sets, per-hypothesis
support counts) must be bit-exact.  Floating point: tr vectors of the RANSAC hypotheses and of the refined pose are
compared with TR_TOL (rotations, rad) and REL_T_TOL (relative translation) -- device sin/cos may differ from glibc
in the last place, everything else is evaluated with the reference's operation order (-fmad=false).
"""
import numpy as np
import pytest

from conftest import make_seeds

pytestmark = pytest.mark.gpu

TR_TOL = 1e-6      # rad, north_star
REL_T_TOL = 1e-6   # relative translation, north_star


def assert_tr_close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert np.all(np.abs(a[..., :3] - b[..., :3]) <= TR_TOL), (a, b)
    scale = np.maximum(np.linalg.norm(b[..., 3:], axis=-1, keepdims=True), 1e-3)
    assert np.all(np.abs(a[..., 3:] - b[..., 3:]) <= REL_T_TOL * scale), (a, b)


def dense_equal(g, o):
    for k in ("idx", "d1", "d2", "valid"):
        assert np.array_equal(g[k], o[k]), k


def random_features(rng, n, w=1241, h=376, dlen=121, integer=True):
    if integer:
        flat = rng.choice(w * h, n, replace=False)
        kp = np.stack([flat % w, flat // w], 1).astype(np.float32)
    else:
        kp = (rng.random((n, 2)) * [w, h]).astype(np.float32)
    d = rng.integers(-1020, 1021, size=(n, dlen)).astype(np.float32)
    return kp, d


# ------------------------------------------------------------------------------------------------ match_desc

def test_sort_matches_is_std_sort_order(ctx, oracle):
    """viso.cpp:724: the parallel introsort reproduces libstdc++'s permutation of equal distances (the oracle's
    restatement is checked against the real std::sort in tests/test_oracle_golden.py)"""
    rng = np.random.default_rng(9)
    cases = [(0, 5), (1, 5), (2, 1), (16, 3), (17, 3), (33, 2), (700, 40), (2040, 100000), (2040, 300), (5000, 7),
             (3000, 1), (12000, 50), (15001, 9), (40000, 2000)]
    for n, hi in cases:
        m = np.stack([rng.permutation(n), rng.integers(0, 10 ** 6, n), rng.integers(0, hi, n)], 1).astype(np.int32)
        assert np.array_equal(ctx.sort_matches(m), oracle.sort_matches(m)), (n, hi)
    # adversarial shapes: organ pipe / sorted / reverse sorted runs exercise the depth limit (heapsort branch)
    for d in (np.r_[np.arange(3000), np.arange(3000)[::-1]], np.arange(5000), np.arange(5000)[::-1] // 3,
              np.r_[np.arange(1, 4001, 2), np.arange(0, 4000, 2)]):
        n = len(d)
        m = np.stack([np.arange(n), np.arange(n)[::-1], d], 1).astype(np.int32)
        assert np.array_equal(ctx.sort_matches(m), oracle.sort_matches(m)), n


@pytest.mark.parametrize("mode", ["stereo", "temporal"])
def test_match_desc_synthetic_frames(ctx, api, oracle, small_sequence, mode):
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    F = oracle.F_from_P(P1, P2)
    for t in (1, 3):
        f, fp = frames[t], frames[t - 1]
        if mode == "stereo":
            args = (f["kpL"], f["kpR"], f["dL"], f["dR"])
            so, sg = oracle.match_params_stereo(F), api.match_params_stereo(F)
        else:
            args = (f["kpL"], fp["kpL"], f["dL"], fp["dL"])
            so, sg = oracle.match_params_temporal(), api.match_params_temporal()
        o = oracle.match_desc(*args, so)
        dense_equal(ctx.match_desc_dense(*args, sg), o)
        assert np.array_equal(ctx.match_desc(*args, sg), o["matches"])
        assert len(o["matches"]) > 50


@pytest.mark.parametrize("n,K,radius,integer", [
    (3000, 250, 80.0, True),     # top-K truncation binds for most queries (dense), many L1 ties
    (3000, 16, 80.0, True),      # tiny K: threshold refinement inside a distance bin
    (1500, 200, 80.0, False),    # float coordinates: real (dist, idx) selection
    (800, 5, 300.0, True),       # radius larger than the image height
    (500, 250, 0.0, True),       # radius 0: only coincident points
    (400, 1, 40.0, True),
])
def test_match_desc_random(ctx, api, oracle, n, K, radius, integer):
    rng = np.random.default_rng(n * 7 + K)
    w, h = (400, 300) if n >= 1500 else (1241, 376)
    kp1, d1 = random_features(rng, n, w, h, integer=integer)
    kp2, d2 = random_features(rng, n + 17, w, h, integer=integer)
    if radius == 0.0:
        kp2[5:200] = kp1[5:200]
    # make SAD ties likely: copy some descriptors
    d2[rng.integers(0, len(d2), 300)] = d2[rng.integers(0, len(d2), 300)]
    for second in (0, 1):
        so = oracle.match_params_temporal(); sg = api.match_params_temporal()
        for s in (so, sg):
            s.max_neighbors = K; s.radius = radius; s.enforce_2nd_best = second
        o = oracle.match_desc(kp1, kp2, d1, d2, so)
        dense_equal(ctx.match_desc_dense(kp1, kp2, d1, d2, sg), o)
        assert np.array_equal(ctx.match_desc(kp1, kp2, d1, d2, sg), o["matches"])


def test_match_desc_index0_terminator(ctx, api, oracle):
    """target index 0 ends the reference's candidate scan (viso.cpp:693): put it in the middle of the image"""
    rng = np.random.default_rng(5)
    kp1, d1 = random_features(rng, 1200, 300, 200)
    kp2, d2 = random_features(rng, 1200, 300, 200)
    kp2[0] = (150, 100)
    so, sg = oracle.match_params_temporal(), api.match_params_temporal()
    o = oracle.match_desc(kp1, kp2, d1, d2, so)
    assert (o["idx"] == -1).sum() > 0 and (o["idx"] == 0).sum() == 0
    dense_equal(ctx.match_desc_dense(kp1, kp2, d1, d2, sg), o)


def test_match_desc_duplicate_points_and_identical_descriptors(ctx, api, oracle):
    """coincident keypoints (L1 ties broken by index) and identical descriptors (SAD ties -> last in scan order)"""
    rng = np.random.default_rng(6)
    base, d = random_features(rng, 60, 200, 120)
    kp2 = np.repeat(base, 8, axis=0); d2 = np.repeat(d, 8, axis=0)
    kp1, d1 = random_features(rng, 300, 200, 120)
    d1[:50] = d[:50]
    for K in (7, 64, 250):
        so, sg = oracle.match_params_temporal(), api.match_params_temporal()
        so.max_neighbors = sg.max_neighbors = K
        o = oracle.match_desc(kp1, kp2, d1, d2, so)
        dense_equal(ctx.match_desc_dense(kp1, kp2, d1, d2, sg), o)
        assert np.array_equal(ctx.match_desc(kp1, kp2, d1, d2, sg), o["matches"])


def test_match_desc_epipolar_gate_general_F(ctx, api, oracle):
    rng = np.random.default_rng(8)
    kp1, d1 = random_features(rng, 1500, 640, 300)
    kp2, d2 = random_features(rng, 1500, 640, 300)
    F = np.array([[1e-7, 2e-6, -3e-4], [-1e-6, 2e-7, 4e-2], [2e-4, -4.1e-2, 1.0]])
    so, sg = oracle.match_params_stereo(F), api.match_params_stereo(F)
    so.sampson_thresh = sg.sampson_thresh = 4.0
    o = oracle.match_desc(kp1, kp2, d1, d2, so)
    assert o["valid"].sum() > 100
    dense_equal(ctx.match_desc_dense(kp1, kp2, d1, d2, sg), o)
    # the mono configuration of calibratedSFM (viso.cpp:1384-1390): general F, ratio test .9, radius 10
    so, sg = oracle.match_params_stereo(F), api.match_params_stereo(F)
    for s_ in (so, sg):
        s_.sampson_thresh = 4.0; s_.enforce_2nd_best = 1; s_.ratio_2nd_best = .9; s_.radius = 10.0
    kp2b = kp1 + rng.integers(-3, 4, size=kp1.shape).astype(np.float32)     # targets near the queries
    d2b = np.clip(d1 + rng.integers(-40, 41, size=d1.shape), -1020, 1020).astype(np.float32)
    o = oracle.match_desc(kp1, kp2b, d1, d2b, so)
    dense_equal(ctx.match_desc_dense(kp1, kp2b, d1, d2b, sg), o)
    assert np.array_equal(ctx.match_desc(kp1, kp2b, d1, d2b, sg), o["matches"])
    # degenerate F: every Sampson distance is NaN -> nothing matches
    so, sg = oracle.match_params_stereo(np.zeros((3, 3))), api.match_params_stereo(np.zeros((3, 3)))
    o = oracle.match_desc(kp1, kp2, d1, d2, so)
    assert o["valid"].sum() == 0
    dense_equal(ctx.match_desc_dense(kp1, kp2, d1, d2, sg), o)


def test_match_desc_empty_and_ragged(ctx, api, oracle):
    rng = np.random.default_rng(9)
    kp, d = random_features(rng, 40)
    sg = api.match_params_temporal(); so = oracle.match_params_temporal()
    e_kp = np.zeros((0, 2), np.float32); e_d = np.zeros((0, 121), np.float32)
    assert len(ctx.match_desc(e_kp, kp, e_d, d, sg)) == 0
    g = ctx.match_desc_dense(kp, e_kp, d, e_d, sg)
    assert (g["idx"] == -1).all() and (g["valid"] == 0).all()
    assert len(ctx.match_desc(kp, e_kp, d, e_d, sg)) == 0
    # one target only: it is index 0, the terminator, so never matched
    g = ctx.match_desc_dense(kp, kp[:1], d, d[:1], sg)
    dense_equal(g, oracle.match_desc(kp, kp[:1], d, d[:1], so))
    # short descriptors and keypoints far outside the default image extent
    kp1, d1 = random_features(rng, 300, 3000, 2000, dlen=9)
    kp2, d2 = random_features(rng, 333, 3000, 2000, dlen=9)
    kp1[:10] -= 500
    dense_equal(ctx.match_desc_dense(kp1, kp2, d1, d2, sg), oracle.match_desc(kp1, kp2, d1, d2, so))


def test_match_desc_domain_errors(ctx, api):
    rng = np.random.default_rng(10)
    kp, d = random_features(rng, 40)
    sg = api.match_params_temporal()
    bad = d.copy(); bad[3, 7] = 0.5
    with pytest.raises(api.VisoError) as e:
        ctx.match_desc(kp, kp, bad, d, sg)
    assert e.value.code == -3
    bad = d.copy(); bad[3, 7] = 5000
    with pytest.raises(api.VisoError):
        ctx.match_desc(kp, kp, d, bad, sg)
    sg.max_neighbors = 0
    with pytest.raises(api.VisoError):
        ctx.match_desc(kp, kp, d, d, sg)


def test_match_desc_20k_properties(ctx, api, oracle):
    """BASELINE config 3 size (20k keypoints / image).  Full oracle comparison on a query subset (the oracle's scan
    is per query, so a subset of queries against the full target set is exact), plus size-independent properties."""
    from libviso_b200 import synth
    pair = synth.make_dense_pair(20000, seed=2000)
    P1, P2 = synth.kitti_calib()
    F = oracle.F_from_P(P1, P2)
    for name, so, sg in (("stereo", oracle.match_params_stereo(F), api.match_params_stereo(F)),
                         ("temporal", oracle.match_params_temporal(), api.match_params_temporal())):
        g = ctx.match_desc_dense(pair["kpL"], pair["kpR"], pair["dL"], pair["dR"], sg)
        sub = np.random.default_rng(1).choice(20000, 400, replace=False)
        o = oracle.match_desc(pair["kpL"][sub], pair["kpR"], pair["dL"][sub], pair["dR"], so)
        for k in ("idx", "d1", "d2", "valid"):
            assert np.array_equal(g[k][sub], o[k]), (name, k)
        v = g["valid"] == 1
        # properties: the reported distance is the SAD of the reported pair; the pair is within the radius; index 0
        # is never matched; d1 <= d2
        sad = np.abs(pair["dL"][v] - pair["dR"][g["idx"][v]]).sum(1).astype(np.int64)
        assert np.array_equal(sad, g["d1"][v])
        l1 = np.abs(pair["kpL"][v] - pair["kpR"][g["idx"][v]]).sum(1)
        assert (l1 <= 80).all() and (g["idx"][v] > 0).all() and (g["d1"] <= g["d2"]).all()
        m = ctx.match_desc(pair["kpL"], pair["kpR"], pair["dL"], pair["dR"], sg)
        assert len(m) == v.sum() and (np.diff(m[:, 2]) >= 0).all()
        assert np.array_equal(np.sort(m[:, 0]), np.nonzero(v)[0])
        assert np.array_equal(m, oracle.sort_matches(np.stack([np.nonzero(v)[0], g["idx"][v], g["d1"][v]], 1)))


# ------------------------------------------------------------------------------------------------ circle / geometry

def test_match_circle(ctx, oracle, small_sequence):
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    F = oracle.F_from_P(P1, P2)
    f, fp = frames[2], frames[1]
    mlr = oracle.match_desc(f["kpL"], f["kpR"], f["dL"], f["dR"], oracle.match_params_stereo(F))["matches"]
    mlrp = oracle.match_desc(fp["kpL"], fp["kpR"], fp["dL"], fp["dR"], oracle.match_params_stereo(F))["matches"]
    m11 = oracle.match_desc(f["kpL"], fp["kpL"], f["dL"], fp["dL"], oracle.match_params_temporal())["matches"]
    m22 = oracle.match_desc(f["kpR"], fp["kpR"], f["dR"], fp["dR"], oracle.match_params_temporal())["matches"]
    co, po = oracle.match_circle(mlr, mlrp, m11, m22)
    cg, pg = ctx.match_circle(mlr, mlrp, m11, m22)
    assert len(co) > 20
    assert np.array_equal(cg, co) and np.array_equal(pg, po)
    e = np.zeros((0, 3), np.int32)
    assert len(ctx.match_circle(e, mlrp, m11, m22)[0]) == 0
    assert len(ctx.match_circle(mlr, mlrp, e, m22)[0]) == 0


def test_match_circle_rejects_duplicates(ctx, api):
    m = np.array([[0, 1, 5], [1, 2, 6]], np.int32)
    dup = np.array([[0, 1, 5], [0, 2, 6]], np.int32)
    with pytest.raises(api.VisoError) as e:
        ctx.match_circle(m, m, dup, m)
    assert e.value.code == -5


def test_triangulate_and_project(ctx, oracle):
    from libviso_b200 import synth
    rng = np.random.default_rng(11)
    m = 5000
    x = np.stack([rng.uniform(0, 1241, m), rng.uniform(0, 376, m), rng.uniform(0, 1241, m), rng.uniform(0, 376, m)])
    x[2, :10] = x[0, :10]            # zero disparity -> inf/nan exactly like the reference (no clamp)
    x = np.round(x)                  # integer pixels as produced by the detector
    Xg = ctx.triangulate_rectified_f64(x, synth.F_PX, synth.BASE, synth.CU, synth.CV)
    Xo = oracle.triangulate_rectified_f64(x, synth.F_PX, synth.BASE, synth.CU, synth.CV)
    assert np.array_equal(Xg, Xo, equal_nan=True)
    x1 = x[:2].astype(np.float32); x2 = x[2:].astype(np.float32)
    Xg = ctx.triangulate_rectified_f32(x1, x2, synth.F_PX, synth.BASE, synth.CU, synth.CV)
    Xo = oracle.triangulate_rectified_f32(x1, x2, synth.F_PX, synth.BASE, synth.CU, synth.CV)
    assert np.array_equal(Xg, Xo)
    P1, P2 = synth.kitti_calib()
    X = np.stack([rng.uniform(-20, 20, m), rng.uniform(-2, 3, m), rng.uniform(4, 60, m)])
    assert np.array_equal(ctx.project_points(X, P2), oracle.project_points(X, P2))
    X[2, 7] = 0.0
    with pytest.raises(OverflowError):
        ctx.project_points(X, P1)
    with pytest.raises(OverflowError):
        oracle.project_points(X, P1)
    # collect_matches + triangulate fused
    kp1 = np.round(rng.uniform(0, 1241, (300, 2))).astype(np.float32)
    kp2 = np.round(rng.uniform(0, 1241, (310, 2))).astype(np.float32)
    mt = np.stack([rng.integers(0, 300, 200), rng.integers(0, 310, 200), rng.integers(0, 9999, 200)], 1).astype(np.int32)
    xg, Xg = ctx.collect_triangulate(kp1, kp2, mt, synth.F_PX, synth.BASE, synth.CU, synth.CV)
    xo = oracle.collect_matches(kp1, kp2, mt)
    assert np.array_equal(xg, xo)
    assert np.array_equal(Xg, oracle.triangulate_rectified_f64(xo, synth.F_PX, synth.BASE, synth.CU, synth.CV),
                          equal_nan=True)


# ------------------------------------------------------------------------------------------------ estimation

def _params(api, oracle, H):
    from libviso_b200 import synth
    kw = dict(base=synth.BASE, f=synth.F_PX, cu=synth.CU, cv=synth.CV, ransac_iter=H)
    return api.param_default(**kw), oracle.param_default(**kw)


def test_get_inliers_and_minimize_reproj(ctx, api, oracle):
    from libviso_b200 import synth
    X, obs, tr_true = synth.make_ransac_problem(2000, seed=3001)
    pg, po = _params(api, oracle, 50)
    for tr in (np.zeros(6), tr_true, tr_true + 0.01):
        io, _, margin = oracle.get_inliers(X, obs, tr, po)
        ig = ctx.get_inliers(X, obs, tr, pg)
        assert margin > 1e-9, "fixture is knife-edge"
        assert np.array_equal(ig, io)
    # the reference's disabled known-answer recipe (test/test.cpp:51-114): noise-free points, GN from 0
    rng = np.random.default_rng(1)
    Xk = np.stack([rng.uniform(-10, 10, 10), rng.uniform(-2, 2, 10), rng.uniform(5, 40, 10)])
    tr0 = np.array([0, 0, 0, 1.0, 0, 0])
    T = oracle.tr2mat(tr0)
    Xc = T[:3, :3] @ Xk + T[:3, 3:]
    ob = np.stack([synth.F_PX * Xc[0] / Xc[2] + synth.CU, synth.F_PX * Xc[1] / Xc[2] + synth.CV,
                   synth.F_PX * (Xc[0] - synth.BASE) / Xc[2] + synth.CU, synth.F_PX * Xc[1] / Xc[2] + synth.CV])
    act = np.arange(10, dtype=np.int32)
    ok_o, tr_o, _ = oracle.minimize_reproj(Xk, ob, np.zeros(6), po, act)
    ok_g, tr_g = ctx.minimize_reproj(Xk, ob, np.zeros(6), pg, act)
    assert ok_o and ok_g and np.abs(tr_o - tr0).sum() < 1e-4 and np.abs(tr_g - tr0).sum() < 1e-4
    assert_tr_close(tr_g, tr_o)
    # noisy subset with a non-trivial active list (exercises the observe(0,i) weight quirk, viso.cpp:1449)
    act = np.sort(rng.choice(2000, 700, replace=False)).astype(np.int32)
    io, _, _ = oracle.get_inliers(X, obs, tr_true, po)
    act = np.intersect1d(act, io).astype(np.int32)
    ok_o, tr_o, it = oracle.minimize_reproj(X, obs, tr_true * 0.9, po, act)
    ok_g, tr_g = ctx.minimize_reproj(X, obs, tr_true * 0.9, pg, act)
    assert ok_o == ok_g
    assert_tr_close(tr_g, tr_o)
    # singular: all three points identical -> solve fails -> false, tr untouched
    act = np.array([5, 5, 5], np.int32)
    ok_o, tr_o, _ = oracle.minimize_reproj(X, obs, np.zeros(6), po, act)
    ok_g, tr_g = ctx.minimize_reproj(X, obs, np.zeros(6), pg, act)
    assert ok_o == ok_g
    assert_tr_close(tr_g, tr_o)


def test_device_sincos_is_glibc_bit_for_bit(ctx, tmp_path):
    """libviso_b200/csrc/glibc_sincos.h on the device against the host libm (what the oracle and the reference call,
    viso.cpp:1410-1411): identical bits over every range below glibc's huge-argument reduction"""
    from test_sincos import build_replica, libm_sincos, sincos_arguments
    x = sincos_arguments(seed=1, n=300000)
    x = np.ascontiguousarray(x[np.abs(x) < float.fromhex("0x1.921fbp+26")])   # glibc's test is on the high word: k < 0x419921FB
    s, c = ctx.debug_sincos(x)
    s0, c0 = libm_sincos(build_replica(tmp_path), x)
    assert np.array_equal(s.view(np.int64), s0.view(np.int64))
    assert np.array_equal(c.view(np.int64), c0.view(np.int64))
    assert len(x) > 2_000_000


@pytest.mark.parametrize("n,H", [(400, 50), (10000, 256), (10000, 4096)])   # the last one is BASELINE configs[3]
def test_ransac_minimize_reproj(ctx, api, oracle, n, H):
    from libviso_b200 import synth
    X, obs, tr_true = synth.make_ransac_problem(n, seed=3000 + n)
    pg, po = _params(api, oracle, H)
    table = oracle.randomsample_table(424242, H, n)
    o = oracle.ransac_minimize_reproj(X, obs, po, table)
    g = ctx.ransac_minimize_reproj(X, obs, pg, table)
    assert o["ok"] and g["ok"]
    assert np.array_equal(g["hyp_ok"], o["hyp_ok"])
    assert np.array_equal(g["hyp_count"], o["hyp_count"])
    okm = o["hyp_ok"] == 1
    assert np.array_equal(g["hyp_tr"][okm], o["hyp_tr"][okm])   # same arithmetic, same libm: the hypotheses are identical
    assert g["best_hyp"] == o["best_hyp"]
    assert np.array_equal(g["inliers"], o["inliers"])
    assert_tr_close(g["tr"], o["tr"])
    assert np.abs(g["tr"] - tr_true).max() < 0.05


def test_ransac_hypotheses_do_not_depend_on_the_iteration_cap(ctx, api, oracle):
    """viso_set_hyp_iteration_cap: where a hypothesis moves from the quad kernel to the warp-per-hypothesis kernel changes
    which lanes compute what, never the values -- every output, including the tr of the hypotheses that fail after 100
    iterations, is bit-identical for every cap, and equal to the oracle"""
    from libviso_b200 import synth
    n, H = 400, 3000
    X, obs, _ = synth.make_ransac_problem(n, seed=77, outlier_frac=0.3)
    pg, po = _params(api, oracle, H)
    table = oracle.randomsample_table(99, H, n)
    o = oracle.ransac_minimize_reproj(X, obs, po, table)
    try:
        ctx.set_hyp_iteration_cap(100)
        ref = ctx.ransac_minimize_reproj(X, obs, pg, table)
        assert np.array_equal(ref["hyp_ok"], o["hyp_ok"]) and np.array_equal(ref["hyp_count"], o["hyp_count"])
        assert (o["hyp_ok"] == 0).sum() >= 10            # there ARE hypotheses that run out of iterations
        for cap in (1, 2, 5, 8, 30, 99):
            ctx.set_hyp_iteration_cap(cap)
            g = ctx.ransac_minimize_reproj(X, obs, pg, table)
            for k in ("hyp_ok", "hyp_count", "inliers"):
                assert np.array_equal(g[k], ref[k]), (cap, k)
            assert np.array_equal(g["hyp_tr"].view(np.int64), ref["hyp_tr"].view(np.int64)), cap
            assert g["best_hyp"] == ref["best_hyp"] and np.array_equal(g["tr"].view(np.int64), ref["tr"].view(np.int64))
    finally:
        ctx.set_hyp_iteration_cap(8)


def test_ransac_zero_iterations(ctx, api, oracle):
    """param.ransac_iter == 0: the loop of viso.cpp:1555 never runs -> false, best_tr untouched, no inliers"""
    from libviso_b200 import synth
    X, obs, _ = synth.make_ransac_problem(50, seed=1)
    pg, _ = _params(api, oracle, 0)
    tr0 = np.array([0.1, 0.2, 0.3, 1, 2, 3.0])
    g = ctx.ransac_minimize_reproj(X, obs, pg, np.zeros((0, 3), np.int32), tr0)
    assert not g["ok"] and np.array_equal(g["tr"], tr0) and len(g["inliers"]) == 0 and g["best_hyp"] == -1


def test_ransac_failure_modes(ctx, api, oracle):
    from libviso_b200 import synth
    rng = np.random.default_rng(2)
    pg, po = _params(api, oracle, 20)
    # pure noise: fewer than 6 inliers for every hypothesis -> false, best_tr = winning hypothesis (or the input)
    n = 60
    X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-2, 3, n), rng.uniform(4, 60, n)])
    obs = np.stack([rng.uniform(0, 1241, n), rng.uniform(0, 376, n), rng.uniform(0, 1241, n), rng.uniform(0, 376, n)])
    table = oracle.randomsample_table(7, 20, n)
    tr0 = np.array([0.1, 0.2, 0.3, 1, 2, 3.0])
    o = oracle.ransac_minimize_reproj(X, obs, po, table, tr0)
    g = ctx.ransac_minimize_reproj(X, obs, pg, table, tr0)
    assert not o["ok"] and not g["ok"]
    assert np.array_equal(g["hyp_ok"], o["hyp_ok"]) and np.array_equal(g["hyp_count"], o["hyp_count"])
    assert g["best_hyp"] == o["best_hyp"] and np.array_equal(g["inliers"], o["inliers"])
    assert_tr_close(g["tr"], o["tr"])
    # fewer than 3 correspondences: nothing to sample
    g = ctx.ransac_minimize_reproj(X[:, :2], obs[:, :2], pg, table, tr0)
    assert not g["ok"] and np.array_equal(g["tr"], tr0)


# ------------------------------------------------------------------------------------------------ whole sequence

def test_sequence_pipeline(ctx, api, oracle, small_sequence):
    """viso.cpp:1205-1327 over 6 frames: every intermediate product is compared"""
    from libviso_b200 import synth
    frames, gt = small_sequence
    frames = [dict(f) for f in frames]
    # ragged input: one frame with far fewer keypoints, one empty right image
    frames[4] = {k: v[:150] for k, v in frames[4].items()}
    frames[5]["kpR"] = frames[5]["kpR"][:0]
    frames[5]["dR"] = frames[5]["dR"][:0]
    nF, H = len(frames), 50
    P1, P2 = synth.kitti_calib()
    seeds = make_seeds(nF, H)
    po = oracle.param_default(ransac_iter=H)
    pg = api.param_default(ransac_iter=H)
    o = oracle.sequence(frames, P1, P2, po, seeds, dump=True)
    seq = ctx.sequence(nF, 640, 121, H)
    seq.set_calib(P1, P2)
    seq.upload(frames)
    seq.run(pg, seeds)
    rec = seq.download()
    for t in range(nF):
        assert np.array_equal(seq.get_lr_matches(t), o["lr_matches"][t]), t
        if t == 0:
            continue
        assert np.array_equal(seq.get_dense(1, t), o["m11"][t]), t
        assert np.array_equal(seq.get_dense(2, t), o["m22"][t]), t
        c4, _ = seq.get_circ(t)
        assert np.array_equal(c4, o["circ"][t]), t
        assert np.array_equal(seq.get_inliers(t), o["inliers"][t]), t
    for k in ("ok", "n_inliers", "n_circ", "best_hyp"):
        assert np.array_equal(rec[k], o["records"][k]), k
    assert_tr_close(rec["tr"], o["records"]["tr"])
    assert rec["ok"][1:4].all()
    poses = api.chain_poses(rec)
    assert len(poses) == len(o["poses"])
    assert np.abs(poses - o["poses"]).max() < 1e-6
    # the estimated motion is the ground-truth motion of the synthetic scene (sanity, loose)
    for t in (1, 2, 3):
        assert np.abs(api.tr2mat(rec["tr"][t]) - gt[t]).max() < 0.05
    # idempotence: a second run over the resident inputs reproduces the records bit for bit
    seq.run(pg)
    rec2 = seq.download()
    assert rec2.tobytes() == rec.tobytes()
    mb, pairs, evaluated = seq.stats()
    assert mb > 0 and 0 < evaluated <= pairs
    seq.close()


# ------------------------------------------------------------------------------------------------ device front-end

def test_device_extractor_matches_reference_descriptors(ctx, api, small_sequence):
    """MyFeatureExtractor on the device (viso.cpp:1004-1024): packed rows == the cv::Sobel-based descriptors + 1024,
    including keypoints on / outside the border and non-integral coordinates (Point2i rounding)"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    f = frames[2]
    rng = np.random.default_rng(5)
    H, W = f["imL"].shape
    extra = np.array([[0, 0], [1, 1], [W - 1, H - 1], [W - 2, 3], [5, H - 1], [0.5, 1.5], [2.5, 3.5], [100.49, 50.51],
                      [W + 3, 10], [-4, 7], [300, H + 2]], np.float32)
    kpL = np.concatenate([f["kpL"], extra, (rng.random((40, 2)) * [W, H]).astype(np.float32)])
    kpR = np.concatenate([f["kpR"], extra])
    seq = ctx.sequence(2, len(kpL) + 8, 121, 8)
    seq.set_image_size(W, H)
    seq.upload_frame_images(0, f["imL"], f["imR"], kpL, kpR)
    seq.upload_frame_images(1, f["imL"], f["imR"], kpL, kpR)
    seq.set_calib(*synth.kitti_calib())
    seq.run(api.param_default(ransac_iter=8), make_seeds(2, 8))
    ctx.sync()
    for side, (img, kp) in enumerate(((f["imL"], kpL), (f["imR"], kpR))):
        want = synth.extract_descriptors(img, kp).astype(np.int32) + 1024
        got = seq.get_packed(0, side).astype(np.int32)
        assert got.shape == (len(kp), 128)
        assert np.array_equal(got[:, :121], want)
        assert (got[:, 121:] == 0).all()
    seq.close()


def test_sequence_from_images_and_chunked_ranges(ctx, api, oracle, small_sequence):
    """images + keypoints in, chunked run_range() submissions: records identical to the one-shot f32-descriptor run
    and to the oracle"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    H = 50
    seeds = make_seeds(len(frames), H)
    po = oracle.param_default(ransac_iter=H)
    o = oracle.sequence(frames, P1, P2, po, seeds)
    pg = api.param_default(ransac_iter=H)
    cap = max(max(len(f["kpL"]), len(f["kpR"])) for f in frames)
    seq = ctx.sequence(len(frames), cap, 121, H)
    seq.set_calib(P1, P2)
    seq.set_image_size(synth.W, synth.H)
    seq.set_seeds(seeds, H)
    # chunks of 2 frames; frames 2,3 arrive as f32 descriptors, the others as images
    for t0 in range(0, len(frames), 2):
        for t in range(t0, min(t0 + 2, len(frames))):
            f = frames[t]
            if t in (2, 3):
                seq.upload_frame(t, f["kpL"], f["kpR"], f["dL"], f["dR"])
            else:
                seq.upload_frame_images(t, f["imL"], f["imR"], f["kpL"], f["kpR"])
        seq.run_range(pg, t0, min(t0 + 2, len(frames)))
    rec = seq.download()
    for k in ("ok", "n_inliers", "n_circ", "best_hyp"):
        assert np.array_equal(rec[k], o["records"][k]), k
    assert_tr_close(rec["tr"], o["records"]["tr"])
    # re-upload + re-run without an intervening host sync must wait for the kernels still reading the frames
    for rep in range(2):
        for t, f in enumerate(frames):
            seq.upload_frame_images(t, f["imL"], f["imR"], f["kpL"], f["kpR"])
        seq.run(pg)
    rec2 = seq.download()
    assert rec2.tobytes() == rec.tobytes()
    seq.close()


# ------------------------------------------------------------------------------------------------ mvg / estimation

def test_triangulate_dlt_and_solve_rigid_motion(ctx, oracle):
    """mvg.cpp:124-169 and estimation.cpp:29-51 (test.cpp:9-39, :170-206 recipes).  SVD-based: tolerance, not bits."""
    from libviso_b200 import synth
    P1, P2 = synth.kitti_calib()
    rng = np.random.default_rng(21)
    X = np.stack([rng.uniform(-20, 20, 500), rng.uniform(-2, 3, 500), rng.uniform(4, 60, 500)])
    x1 = oracle.project_points(X, P1).astype(np.float32)
    x2 = oracle.project_points(X, P2).astype(np.float32)
    got = ctx.triangulate_dlt(x1, x2, P1, P2)
    want = oracle.triangulate_dlt(x1, x2, P1, P2)
    assert np.allclose(got, want, rtol=2e-4, atol=1e-4)
    assert np.allclose(got, X, rtol=2e-2, atol=5e-2)      # test.cpp:36: |X - Xt| < 1e-2-ish at float pixel precision
    # Kabsch: R = Rx(pi/2), t = (1,2,3) (test.cpp:170-206) and random motions
    c, s = np.cos(np.pi / 2), np.sin(np.pi / 2)
    T1 = np.array([[1, 0, 0, 1], [0, c, -s, 2], [0, s, c, 3], [0, 0, 0, 1]])
    for T_true, n in ((T1, 3), (T1, 200), (oracle.tr2mat([0.1, -0.2, 0.05, 0.5, -1, 2]), 1000)):
        B = rng.standard_normal((3, n)).astype(np.float32)
        if n == 3:
            B = np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0]], np.float32).T.copy()
        A = (T_true[:3, :3] @ B + T_true[:3, 3:4]).astype(np.float32)
        T = ctx.solve_rigid_motion(A, B)
        assert np.abs(T - oracle.solve_rigid_motion(A, B)).max() < 1e-4
        assert np.abs(T - T_true).max() < 1e-4


def test_long_sequence_properties(ctx, api, oracle, small_sequence):
    """BASELINE configs[1] shape in miniature (240 frames driven back and forth over the rendered ones): the
    size-independent properties -- rerun idempotence, images == descriptors, chunked == one-shot, frame pair (t-1, t)
    of the long sequence == the same pair computed alone, oracle agreement on a prefix."""
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    F, H = 240, 50
    U = len(frames)
    order = [(t % (2 * (U - 1))) if (t % (2 * (U - 1))) < U else 2 * (U - 1) - (t % (2 * (U - 1))) for t in range(F)]
    seeds = make_seeds(F, H)
    pg = api.param_default(ransac_iter=H)
    cap = max(max(len(f["kpL"]), len(f["kpR"])) for f in frames)

    def run(mode):
        seq = ctx.sequence(F, cap, 121, H)
        seq.set_calib(P1, P2)
        seq.set_image_size(synth.W, synth.H)
        seq.set_seeds(seeds, H)
        for t in range(F):
            f = frames[order[t]]
            if mode == "desc":
                seq.upload_frame(t, f["kpL"], f["kpR"], f["dL"], f["dR"])
            else:
                seq.upload_frame_images(t, f["imL"], f["imR"], f["kpL"], f["kpR"])
        if mode == "chunks":
            for t0 in range(0, F, 37):
                seq.run_range(pg, t0, min(F, t0 + 37))
        else:
            seq.run(pg)
        rec = seq.download()
        if mode == "images":
            seq.run(pg)
            assert seq.download().tobytes() == rec.tobytes()      # idempotent
        seq.close()
        return rec

    rec = run("images")
    assert run("desc").tobytes() == rec.tobytes()
    assert run("chunks").tobytes() == rec.tobytes()
    assert rec["ok"][1:].all()
    # a frame pair does not depend on its position in the sequence: pairs (t-1, t) with the same two rendered frames
    # and the same seeds give the same record
    o = oracle.sequence([frames[order[t]] for t in range(12)], P1, P2, oracle.param_default(ransac_iter=H), seeds[:12])
    for k in ("ok", "n_inliers", "n_circ", "best_hyp"):
        assert np.array_equal(rec[k][:12], o["records"][k]), k
    assert_tr_close(rec["tr"][:12], o["records"]["tr"])
    poses = api.chain_poses(rec)
    assert len(poses) == F and np.isfinite(poses).all()


def test_chunk_upload_equals_per_frame_upload(ctx, api, small_sequence):
    """viso_seq_upload_chunk_images (three large copies per chunk, keypoint rows padded to the sequence capacity) gives
    the same records as frame-by-frame uploads"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    F, H = len(frames), 30
    seeds = make_seeds(F, H)
    pg = api.param_default(ransac_iter=H)
    cap = max(max(len(f["kpL"]), len(f["kpR"])) for f in frames)

    def fresh():
        seq = ctx.sequence(F, cap, 121, H)
        seq.set_calib(P1, P2)
        seq.set_image_size(synth.W, synth.H)
        seq.set_seeds(seeds, H)
        return seq

    a = fresh()
    a.upload_images(frames)
    a.run(pg)
    want = a.download()
    a.close()

    b = fresh()
    capq = b.capacity()
    assert capq >= cap
    img = np.zeros((F, 2, synth.H, synth.W), np.uint8)
    kpL = np.full((F, capq, 2), -7.0, np.float32); kpR = np.full((F, capq, 2), -7.0, np.float32)   # padding is ignored
    nL = np.zeros(F, np.int32); nR = np.zeros(F, np.int32)
    for t, f in enumerate(frames):
        img[t, 0], img[t, 1] = f["imL"], f["imR"]
        nL[t], nR[t] = len(f["kpL"]), len(f["kpR"])
        kpL[t, :nL[t]] = f["kpL"]; kpR[t, :nR[t]] = f["kpR"]
    for t0, t1 in ((0, 4), (4, F)):
        b.upload_chunk_images_raw(t0, t1 - t0, img[t0:].ctypes.data, kpL[t0:].ctypes.data, nL[t0:].ctypes.data,
                                  kpR[t0:].ctypes.data, nR[t0:].ctypes.data)
        ctx.sync()          # pageable numpy memory
        b.run_range(pg, t0, t1)
    got = b.download()
    b.close()
    assert got.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------------------------ detector (8f rank 2)

@pytest.mark.gpu
def test_detect_harris_matches_oracle(ctx, oracle, small_sequence):
    """viso_detect_harris against the oracle's canonical cornerHarris + binned top-n: keypoints, their order and their
    responses bit-for-bit.  Cases: KITTI geometry, odd sizes with borders inside bins, quotas above the number of
    columns (no lower bound from the column maxima) and above 256 candidates (bisection), flat regions (bins with
    fewer non-zero responses than the quota), ties (a periodic pattern gives many equal responses)."""
    frames, _ = small_sequence
    rng = np.random.default_rng(5)
    cases = [(frames[0]["imL"], 2040, 24, 5), (frames[1]["imR"], 600, 24, 5), (frames[2]["imL"], 24 * 5 * 166, 24, 5),
             (frames[3]["imL"], 24 * 5 * 300, 24, 5),
             (rng.integers(0, 256, size=(83, 97), dtype=np.uint8), 60, 4, 3),
             (rng.integers(0, 256, size=(64, 200), dtype=np.uint8), 50, 1, 1),      # one 200-px-wide bin: 4 strips
             (rng.integers(0, 256, size=(150, 16), dtype=np.uint8), 30, 1, 2)]
    flat = np.full((90, 120), 90, np.uint8); flat[20:26, 30:37] = 250; flat[60:70, 80:95] = 10
    cases.append((flat, 12 * 40, 4, 3))
    yy, xx = np.mgrid[0:96, 0:128]
    cases.append(((((xx // 8) + (yy // 8)) % 2 * 200 + 20).astype(np.uint8), 8 * 30, 4, 2))   # checkerboard: ties
    for img, n, nx, ny in cases:
        want, wresp = oracle.detect_harris_binned(img, n, nx, ny, 0.04, order_rule=1, with_response=True)
        got, gresp = ctx.detect_harris(img, n, nx, ny, 0.04, with_response=True)
        assert len(got) == len(want), (img.shape, n, len(got), len(want))
        assert np.array_equal(got, want), (img.shape, n)
        assert gresp.tobytes() == wresp.tobytes(), (img.shape, n)


@pytest.mark.gpu
def test_sequence_from_raw_images(ctx, api, oracle, small_sequence):
    """images only: detector + extractor + pipeline on the device against the oracle's front end + pipeline"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    H, nfeat = 50, 600
    seeds = make_seeds(len(frames), H)
    oframes = oracle.frames_from_images([(f["imL"], f["imR"]) for f in frames], nfeat)
    po = oracle.param_default(ransac_iter=H)
    o = oracle.sequence(oframes, P1, P2, po, seeds, dump=True)
    pg = api.param_default(ransac_iter=H)
    seq = ctx.sequence(len(frames), nfeat, 121, H)
    seq.set_calib(P1, P2)
    seq.set_image_size(synth.W, synth.H)
    seq.set_detector(nfeat)
    seq.set_seeds(seeds, H)
    # frames 0-1 in one block, the rest frame by frame, two submissions
    block = np.ascontiguousarray(np.stack([np.stack([frames[t]["imL"], frames[t]["imR"]]) for t in range(2)]))
    seq.upload_chunk_raw(0, 2, block)
    ctx.sync()
    seq.run_range(pg, 0, 2)
    for t in range(2, len(frames)):
        seq.upload_frame_raw_images(t, frames[t]["imL"], frames[t]["imR"])
    seq.run_range(pg, 2, len(frames))
    rec = seq.download()
    for t, f in enumerate(oframes):
        assert np.array_equal(seq.get_keypoints(t, 0), f["kpL"]) and np.array_equal(seq.get_keypoints(t, 1), f["kpR"])
        if t > 0:
            assert np.array_equal(seq.get_dense(1, t), o["m11"][t]), t
    for k in ("ok", "n_inliers", "n_circ", "best_hyp"):
        assert np.array_equal(rec[k], o["records"][k]), k
    assert_tr_close(rec["tr"], o["records"]["tr"])
    assert rec["ok"][1:].all() and (rec["n_inliers"][1:] > 50).all()
    # mixing: frame 3 re-uploaded with host keypoints (the oracle's), same result
    f = oframes[3]
    seq.upload_frame_images(3, f["imL"], f["imR"], f["kpL"], f["kpR"])
    seq.run(pg)
    assert seq.download().tobytes() == rec.tobytes()
    seq.close()


@pytest.mark.gpu
def test_extract_descriptors_standalone(ctx, oracle, small_sequence):
    """viso_extract_descriptors (MyFeatureExtractor::computeImpl, viso.cpp:1004-1024) in the cv::Mat layout: interior,
    border (the > 0 rule at :1018), half-integer coordinates (Point2i rounds half to even)"""
    frames, _ = small_sequence
    img = frames[0]["imL"]
    h, w = img.shape
    rng = np.random.default_rng(9)
    kp = np.concatenate([frames[0]["kpL"][:200],
                         np.array([[0, 0], [w - 1, h - 1], [3, 2], [w - 2, 5], [5, h - 3], [0.5, 1.5], [2.5, 3.5], [10.5, 7.5]], np.float32),
                         np.stack([rng.integers(0, w, 100), rng.integers(0, h, 100)], 1).astype(np.float32)])
    want = oracle.extract_descriptors(oracle.sobel_x(img), kp)
    got = ctx.extract_descriptors(img, kp)
    assert got.tobytes() == want.tobytes()


@pytest.mark.gpu
def test_raw_sequence_with_featureless_frames(ctx, api, oracle, small_sequence):
    """a flat image pair in the middle of a sequence: the detector finds nothing there (every response is 0,
    viso.cpp:956), the two frame pairs around it yield no pose (< 3 circular matches, viso.cpp:1283-1288), the others
    are unaffected; and a pair whose right image is flat (stereo matching finds nothing)"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    H, nfeat = 50, 360
    flat = np.full((synth.H, synth.W), 77, np.uint8)
    imgs = [(frames[0]["imL"], frames[0]["imR"]), (frames[1]["imL"], frames[1]["imR"]), (flat, flat),
            (frames[3]["imL"], frames[3]["imR"]), (frames[4]["imL"], flat), (frames[5]["imL"], frames[5]["imR"])]
    seeds = make_seeds(len(imgs), H)
    oframes = oracle.frames_from_images(imgs, nfeat)
    assert len(oframes[2]["kpL"]) == 0 and len(oframes[4]["kpR"]) == 0
    o = oracle.sequence(oframes, P1, P2, oracle.param_default(ransac_iter=H), seeds)
    seq = ctx.sequence(len(imgs), nfeat, 121, H)
    seq.set_calib(P1, P2)
    seq.set_image_size(synth.W, synth.H)
    seq.set_detector(nfeat)
    for t, (a, b) in enumerate(imgs):
        seq.upload_frame_raw_images(t, a, b)
    seq.run(api.param_default(ransac_iter=H), seeds)
    rec = seq.download()
    for t, f in enumerate(oframes):
        assert np.array_equal(seq.get_keypoints(t, 0), f["kpL"]) and np.array_equal(seq.get_keypoints(t, 1), f["kpR"]), t
    for k in ("ok", "n_inliers", "n_circ", "best_hyp"):
        assert np.array_equal(rec[k], o["records"][k]), k
    assert_tr_close(rec["tr"], o["records"]["tr"])
    assert rec["ok"].tolist() == [0, 1, 0, 0, 0, 0]
    seq.close()


@pytest.mark.gpu
def test_detect_harris_random_geometries(ctx, oracle):
    """40 random image sizes / bin layouts / quotas (tiny bins, one-row bins, bins wider than one 58-column strip,
    widths below the 64-column strip, bins reaching the last row and column) against the oracle, bit for bit"""
    rng = np.random.default_rng(77)
    for case in range(40):
        h, w = int(rng.integers(8, 160)), int(rng.integers(8, 260))
        nbx, nby = int(rng.integers(1, min(7, w) + 1)), int(rng.integers(1, min(7, h) + 1))
        per = int(rng.integers(1, 40))
        n = nbx * nby * per + int(rng.integers(0, nbx * nby))
        kind = case % 4
        if kind == 0:
            img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
        elif kind == 1:   # smooth: blurred noise, like the synthetic frames
            a = rng.random((h + 4, w + 4))
            a = sum(a[i:i + h, j:j + w] for i in range(5) for j in range(5)) / 25
            img = np.clip((a - 0.5) * 900 + 128, 0, 255).astype(np.uint8)
        elif kind == 2:   # piecewise constant: most responses are exactly 0
            img = np.repeat(np.repeat(rng.integers(0, 256, size=((h + 15) // 16, (w + 15) // 16), dtype=np.uint8), 16, 0), 16, 1)[:h, :w]
            img = np.ascontiguousarray(img)
        else:             # two grey levels: many exactly equal responses (ties at the cut)
            img = ((rng.random((h, w)) < 0.08) * 200 + 20).astype(np.uint8)
        want, wresp = oracle.detect_harris_binned(img, n, nbx, nby, 0.04, order_rule=1, with_response=True)
        got, gresp = ctx.detect_harris(img, n, nbx, nby, 0.04, with_response=True)
        assert np.array_equal(got, want), (case, h, w, nbx, nby, n)
        assert gresp.tobytes() == wresp.tobytes(), (case, h, w, nbx, nby, n)


@pytest.mark.gpu
def test_detector_argument_errors(ctx, api):
    seq = ctx.sequence(2, 600, 121, 10)
    img = np.zeros((64, 96), np.uint8)
    with pytest.raises(api.VisoError):          # image size first
        seq.set_detector(240)
    seq.set_image_size(96, 64)
    with pytest.raises(api.VisoError):          # raw upload needs a detector
        seq.upload_frame_raw_images(0, img, img)
    with pytest.raises(api.VisoError):          # more keypoints than the sequence object can hold
        seq.set_detector(24 * 5 * 6)
    with pytest.raises(api.VisoError):          # fewer features than bins: nothing could ever be kept
        seq.set_detector(100)
    with pytest.raises(api.VisoError):          # more bins than pixels (viso.cpp:934)
        seq.set_detector(600, 200, 5)
    seq.set_detector(480)
    with pytest.raises(api.VisoError):          # the bin layout is fixed once set
        seq.set_detector(240)
    seq.close()
    with pytest.raises(api.VisoError):          # standalone: image too small
        ctx.detect_harris(np.zeros((4, 4), np.uint8), 10, 1, 1)
    big = np.zeros((600, 600), np.uint8)
    with pytest.raises(api.VisoError) as e:     # one bin of 360 000 pixels does not fit in shared memory
        ctx.detect_harris(big, 10, 1, 1)
    assert e.value.code == -3
    assert len(ctx.detect_harris(big, 0, 24, 5)) == 0    # n_features 0: nothing to do


@pytest.mark.gpu
def test_sequence_with_dense_clusters_uses_the_pending_list(ctx, api, oracle, small_sequence):
    """a dense blob of keypoints in every frame overflows the tile kernel's per-query candidate lists; those queries
    are handed to the generic kernel through the pending list -- same results as the oracle"""
    from libviso_b200 import synth
    frames, _ = small_sequence
    rng = np.random.default_rng(31)
    P1, P2 = synth.kitti_calib()
    H = 50
    dense = []
    for t, f in enumerate(frames[:4]):
        cells = rng.choice(24 * 24, 160, replace=False)
        blob = np.stack([600 + cells % 24 + 2 * t, 150 + cells // 24], 1).astype(np.float32)
        g = dict(f)
        for side, im in (("L", f["imL"]), ("R", f["imR"])):
            kp = np.concatenate([f["kp" + side], blob - (np.array([9, 0], np.float32) if side == "R" else 0)])
            g["kp" + side] = kp
            g["d" + side] = oracle.extract_descriptors(oracle.sobel_x(im), kp)
        dense.append(g)
    seeds = make_seeds(len(dense), H)
    o = oracle.sequence(dense, P1, P2, oracle.param_default(ransac_iter=H), seeds, dump=True)
    cap = max(max(len(f["kpL"]), len(f["kpR"])) for f in dense)
    seq = ctx.sequence(len(dense), cap, 121, H)
    seq.set_calib(P1, P2)
    seq.upload(dense)
    seq.run(api.param_default(ransac_iter=H), seeds)
    rec = seq.download()
    assert seq.last_pending() > 50
    for t in range(1, len(dense)):
        assert np.array_equal(seq.get_dense(1, t), o["m11"][t]), t
        assert np.array_equal(seq.get_dense(2, t), o["m22"][t]), t
    for k in ("ok", "n_inliers", "n_circ", "best_hyp"):
        assert np.array_equal(rec[k], o["records"][k]), k
    assert_tr_close(rec["tr"], o["records"]["tr"])
    seq.close()


@pytest.mark.gpu
def test_sequence_pending_list_overflow_falls_back_to_the_scan(ctx, api, oracle):
    """9 frames of 3000 keypoints packed into 400 x 300 px: every query has more than max_neighbors points in range, so
    all 75 000 queries of the 25 match jobs are left to the generic kernel -- more than the pending list holds, which
    switches the generic kernel to its scan over all queries.  Matches against the oracle, bit for bit."""
    from libviso_b200 import synth
    rng = np.random.default_rng(41)
    P1, P2 = synth.kitti_calib()
    H, F, n = 10, 9, 3000
    frames = []
    for t in range(F):
        kl, dl = random_features(rng, n, 400, 300)
        kr, dr = random_features(rng, n, 400, 300)
        frames.append(dict(kpL=kl, kpR=kr, dL=dl, dR=dr))
    seeds = make_seeds(F, H)
    o = oracle.sequence(frames, P1, P2, oracle.param_default(ransac_iter=H), seeds, dump=True)
    seq = ctx.sequence(F, n, 121, H)
    seq.set_calib(P1, P2)
    seq.upload(frames)
    seq.run(api.param_default(ransac_iter=H), seeds)
    rec = seq.download()
    assert seq.last_pending() > 65536
    for t in range(F):
        assert np.array_equal(seq.get_lr_matches(t), o["lr_matches"][t]), t
        if t:
            assert np.array_equal(seq.get_dense(1, t), o["m11"][t]), t
            assert np.array_equal(seq.get_dense(2, t), o["m22"][t]), t
    for k in ("ok", "n_inliers", "n_circ", "best_hyp"):
        assert np.array_equal(rec[k], o["records"][k]), k
    seq.close()
