"""match_desc through each of its kernel paths (viso_set_match_mode): the staged tile kernel (descriptor rows of a
tile's neighbourhood copied to shared memory by cp.async.bulk, lane = candidate), the gather tile kernel (rows through
L1, eight lanes per row) and the generic per-query kernel must all reproduce the oracle bit for bit -- indices, SAD
distances, second-best distances, validity -- on inputs that exercise the reference's quirks (viso.cpp:668-722)."""
import numpy as np
import pytest

from test_gpu_parity import dense_equal, random_features

pytestmark = pytest.mark.gpu

MODES = ["auto", "staged", "gather", "generic"]


@pytest.fixture(params=MODES)
def mctx(request, ctx):
    ctx.set_match_mode(request.param)
    yield ctx
    ctx.set_match_mode("auto")


def both(oracle, api, **kw):
    so, sg = oracle.match_params_temporal(), api.match_params_temporal()
    for s in (so, sg):
        for k, v in kw.items():
            setattr(s, k, v)
    return so, sg


def test_modes_synthetic_frames(mctx, api, oracle, small_sequence):
    from libviso_b200 import synth
    frames, _ = small_sequence
    F = oracle.F_from_P(*synth.kitti_calib())
    f, fp = frames[2], frames[1]
    o = oracle.match_desc(f["kpL"], f["kpR"], f["dL"], f["dR"], oracle.match_params_stereo(F))
    dense_equal(mctx.match_desc_dense(f["kpL"], f["kpR"], f["dL"], f["dR"], api.match_params_stereo(F)), o)
    o = oracle.match_desc(f["kpL"], fp["kpL"], f["dL"], fp["dL"], oracle.match_params_temporal())
    dense_equal(mctx.match_desc_dense(f["kpL"], fp["kpL"], f["dL"], fp["dL"], api.match_params_temporal()), o)
    assert o["valid"].sum() > 50


@pytest.mark.parametrize("n,w,h,K,radius,integer", [
    (2000, 1241, 376, 250, 80.0, True),    # the pipeline's density: everything stays on the tile kernels
    (2000, 1241, 376, 250, 80.0, False),   # float coordinates
    (3000, 400, 300, 250, 80.0, True),     # dense: the top-K cut binds, queries go through the pending list
    (6000, 640, 200, 250, 80.0, True),     # dense everywhere: the gather kernel runs match_query on the staged records
    (5000, 640, 200, 40, 80.0, False),     # the same with float coordinates and a small K
    (1200, 640, 200, 40, 30.0, True),      # small radius, small K
    (600, 1241, 376, 250, 500.0, True),    # radius beyond the image: neighbourhood = everything
    (300, 200, 120, 250, 0.0, True),       # radius 0
])
def test_modes_random(mctx, api, oracle, n, w, h, K, radius, integer):
    rng = np.random.default_rng(n + K + int(radius))
    kp1, d1 = random_features(rng, n, w, h, integer=integer)
    kp2, d2 = random_features(rng, n + 13, w, h, integer=integer)
    if radius == 0.0:
        kp2[3:150] = kp1[3:150]
    d2[rng.integers(0, len(d2), 400)] = d2[rng.integers(0, len(d2), 400)]  # identical rows: SAD ties
    d1[:100] = d2[rng.integers(0, len(d2), 100)]                          # exact matches (SAD 0)
    for second in (0, 1):
        so, sg = both(oracle, api, max_neighbors=K, radius=radius, enforce_2nd_best=second)
        dense_equal(mctx.match_desc_dense(kp1, kp2, d1, d2, sg), oracle.match_desc(kp1, kp2, d1, d2, so))


def test_modes_index0_terminator_and_duplicates(mctx, api, oracle):
    rng = np.random.default_rng(77)
    kp1, d1 = random_features(rng, 900, 300, 200)
    base, d = random_features(rng, 150, 300, 200)
    kp2 = np.repeat(base, 6, axis=0)
    d2 = np.repeat(d, 6, axis=0)
    kp2[0] = (150, 100)  # index 0 in the middle of the image ends many scans (viso.cpp:693)
    so, sg = both(oracle, api)
    o = oracle.match_desc(kp1, kp2, d1, d2, so)
    assert (o["idx"] == -1).sum() > 0
    dense_equal(mctx.match_desc_dense(kp1, kp2, d1, d2, sg), o)


def test_modes_points_outside_the_grid_extent(mctx, api, oracle):
    """coordinates beyond the context's grid extent (and negative ones) are clamped into border cells"""
    rng = np.random.default_rng(78)
    kp1, d1 = random_features(rng, 700, 1241, 376)
    kp2, d2 = random_features(rng, 700, 1241, 376)
    kp1[:60] += np.array([1300, 0], np.float32)
    kp2[:60] += np.array([1290, 5], np.float32)
    kp1[60:90] -= np.array([1250, 380], np.float32)
    kp2[60:90] -= np.array([1245, 377], np.float32)
    so, sg = both(oracle, api)
    dense_equal(mctx.match_desc_dense(kp1, kp2, d1, d2, sg), oracle.match_desc(kp1, kp2, d1, d2, so))


def test_modes_general_F(mctx, api, oracle):
    rng = np.random.default_rng(8)
    kp1, d1 = random_features(rng, 1200, 640, 300)
    kp2, d2 = random_features(rng, 1200, 640, 300)
    F = rng.standard_normal((3, 3)) * np.array([[1e-6, 1e-5, 1e-3], [1e-5, 1e-6, 1e-2], [1e-3, 1e-2, 1.0]])
    so, sg = oracle.match_params_stereo(F), api.match_params_stereo(F)
    so.sampson_thresh = sg.sampson_thresh = 4.0
    dense_equal(mctx.match_desc_dense(kp1, kp2, d1, d2, sg), oracle.match_desc(kp1, kp2, d1, d2, so))
