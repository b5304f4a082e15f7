"""ctypes binding for oracle/_ref/libviso_ref.so: the REFERENCE'S OWN hot-path functions, compiled unchanged from
/root/reference/src by oracle/Makefile (target `ref`) against the header stand-ins of compat/ and oracle/shim/.

TEST INFRASTRUCTURE ONLY (tests/test_ref_pin.py pins the restated oracle against it; bench.py may time it as the CPU
baseline).  The library is built where /root/reference exists (here; `__graft_entry__.build()` does it) and travels to
the GPU box as a prebuilt file -- nothing reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libviso_ref.so")
REFERENCE = os.environ.get("VISO_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.exists(SO) or os.path.exists(os.path.join(REFERENCE, "src", "viso.cpp"))


def build(force=False):
    """make -C oracle ref (needs the reference tree); returns the path, or None when neither tree nor library exists"""
    if os.path.exists(os.path.join(REFERENCE, "src", "viso.cpp")):
        cmd = ["make", "-C", _HERE, "REF=" + REFERENCE, "ref"] + (["-B"] if force else [])
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return SO if os.path.exists(SO) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        from oracle import oracle
        oracle.lib()  # libviso_ref.so links libviso_oracle.so (cornerHarris forwards to the canonical evaluation)
        if build() is None:
            raise RuntimeError("oracle/_ref/libviso_ref.so is missing and the reference tree is not here to build it")
        _lib = C.CDLL(SO)
        _lib.vr_sampson_distance.restype = C.c_double
        _lib.vr_sampson_distance.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        _lib.vr_shim_determinant.restype = C.c_double
        _lib.vr_shim_radius_search.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(v):
    return C.c_double(float(v))


def match_desc(kp1, kp2, d1, d2, sp):
    """sp: oracle.MatchParams.  Returns the reference's sorted match list [M, 3]."""
    kp1, kp2, d1, d2 = _f32(kp1), _f32(kp2), _f32(d1), _f32(d2)
    n1, n2 = len(kp1), len(kp2)
    dlen = d1.shape[1] if d1.ndim == 2 else d2.shape[1]
    out = np.zeros((max(n1, 1), 3), np.int32)
    F = _f64(np.array(sp.F[:])).reshape(9)
    n = lib().vr_match_desc(_p(kp1), n1, _p(kp2), n2, _p(d1), _p(d2), dlen, int(sp.enforce_epipolar), _p(F),
                            int(sp.max_neighbors), _d(sp.radius), _d(sp.sampson_thresh), int(sp.enforce_2nd_best),
                            _d(sp.ratio_2nd_best), _p(out))
    return out[:n].copy()


def sampson_distance(F, p1, p2):
    Fc = _f64(F).reshape(9)
    return lib().vr_sampson_distance(_p(Fc), float(p1[0]), float(p1[1]), float(p2[0]), float(p2[1]))


def match_circle(mlr, mlrp, m11, m22):
    mlr, mlrp, m11, m22 = _i32(mlr), _i32(mlrp), _i32(m11), _i32(m22)
    circ = np.zeros((max(len(mlr), 1), 4), np.int32); pcl = np.zeros((max(len(mlr), 1), 3), np.int32)
    n = lib().vr_match_circle(_p(mlr), len(mlr), _p(mlrp), len(mlrp), _p(m11), len(m11), _p(m22), len(m22), _p(circ), _p(pcl))
    return circ[:n].copy(), pcl[:n].copy()


def collect_triangulate(kp1, kp2, matches, f, base, cu, cv):
    kp1, kp2, matches = _f32(kp1), _f32(kp2), _i32(matches)
    m = len(matches)
    x = np.zeros((4, m)); X = np.zeros((3, m))
    lib().vr_collect_triangulate(_p(kp1), len(kp1), _p(kp2), len(kp2), _p(matches), m, _d(f), _d(base), _d(cu), _d(cv), _p(x), _p(X))
    return x, X


def compute_J(X, obs, tr, param, active):
    X, obs, tr, active = _f64(X), _f64(obs), _f64(tr), _i32(active)
    n, na = X.shape[1], len(active)
    J = np.zeros((4 * na, 6)); pred = np.zeros((4, na)); res = np.zeros(4 * na)
    lib().vr_compute_J(_p(X), _p(obs), n, _p(tr), _d(param.base), _d(param.f), _d(param.cu), _d(param.cv), _p(active), na,
                       _p(J), _p(pred), _p(res))
    return J, pred, res


def get_inliers(X, obs, tr, param):
    X, obs, tr = _f64(X), _f64(obs), _f64(tr)
    n = X.shape[1]
    inl = np.zeros(max(n, 1), np.int32)
    c = lib().vr_get_inliers(_p(X), _p(obs), n, _p(tr), _d(param.base), _d(param.f), _d(param.cu), _d(param.cv),
                             _d(param.inlier_threshold), _p(inl))
    return inl[:c].copy()


def minimize_reproj(X, obs, tr, param, active):
    X, obs, active = _f64(X), _f64(obs), _i32(active)
    t = _f64(tr).copy()
    ok = lib().vr_minimize_reproj(_p(X), _p(obs), X.shape[1], _p(t), _d(param.base), _d(param.f), _d(param.cu), _d(param.cv),
                                  _d(param.thresh), _p(active), len(active))
    return bool(ok), t


def ransac_minimize_reproj(X, obs, param, table, tr0=None):
    X, obs, table = _f64(X), _f64(obs), _i32(table)
    n = X.shape[1]
    t = np.zeros(6) if tr0 is None else _f64(tr0).copy()
    inl = np.zeros(max(n, 1), np.int32); ni = C.c_int32(0)
    lib().vr_set_samples(_p(table), None, len(table))
    ok = lib().vr_ransac_minimize_reproj(_p(X), _p(obs), n, _p(t), _d(param.base), _d(param.f), _d(param.cu), _d(param.cv),
                                         _d(param.inlier_threshold), _d(param.thresh), int(param.ransac_iter), _p(inl), C.byref(ni))
    lib().vr_set_samples(None, None, 0)
    return bool(ok), t, inl[:ni.value].copy()


def tr2mat(tr):
    T = np.zeros((4, 4))
    t = _f64(tr)
    lib().vr_tr2mat(_p(t), _p(T))
    return T


def F_from_P(P1, P2, normalise=True):
    P1c, P2c = _f64(P1).reshape(12), _f64(P2).reshape(12)
    F = np.zeros((3, 3))
    lib().vr_F_from_P(_p(P1c), _p(P2c), int(normalise), _p(F))
    return F


def detect(img, n_features, k=0.04):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    cap = max(n_features, 1)
    xy = np.zeros((cap, 2), np.float32); rs = np.zeros(cap, np.float32)
    n = lib().vr_detect(_p(img), h, w, int(n_features), C.c_float(k), _p(xy), _p(rs), cap)
    assert 0 <= n <= cap
    return xy[:n].copy(), rs[:n].copy()


def extract(img, kp):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    kp = _f32(kp)
    d = np.zeros((len(kp), 121), np.float32)
    n = lib().vr_extract(_p(img), img.shape[0], img.shape[1], _p(kp), len(kp), _p(d))
    assert n == len(kp)
    return d


def sequence_odometry(P1, P2, mask0, mask1, begin, end, seeds_one_frame, max_poses=4096):
    """the reference's sequence_odometry over image files; every frame pair draws its RANSAC samples from the same
    [50, 3] seed block (mapped to index triples with the number of circular matches, like the product)"""
    P1c, P2c = _f64(P1).reshape(12), _f64(P2).reshape(12)
    seeds = np.ascontiguousarray(seeds_one_frame, dtype=np.uint32)
    poses = np.zeros((max_poses, 4, 4))
    lib().vr_set_samples(None, _p(seeds), len(seeds))
    n = lib().vr_sequence_odometry(_p(P1c), _p(P2c), mask0.encode(), mask1.encode(), int(begin), int(end), _p(poses), max_poses)
    lib().vr_set_samples(None, None, 0)
    return poses[:n].copy()


RECORD_DTYPE = np.dtype([("tr", np.float64, 6), ("ok", np.int32), ("n_inliers", np.int32), ("n_circ", np.int32), ("best_hyp", np.int32)])


def pipeline(frames, P1, P2, ransac_iter, seeds):
    """the per-frame loop of the reference's sequence_odometry on given features (list of dict(kpL, kpR, dL, dR)):
    returns dict(records [n_frames] (best_hyp is not reported by the reference: -1), poses [n, 4, 4])"""
    nF = len(frames)
    nL = np.array([len(f["kpL"]) for f in frames], np.int32); nR = np.array([len(f["kpR"]) for f in frames], np.int32)
    offL = np.zeros(nF, np.int64); offR = np.zeros(nF, np.int64)
    offL[1:] = np.cumsum(nL)[:-1]; offR[1:] = np.cumsum(nR)[:-1]
    kpL = _f32(np.concatenate([f["kpL"] for f in frames])); kpR = _f32(np.concatenate([f["kpR"] for f in frames]))
    dL = _f32(np.concatenate([f["dL"] for f in frames])); dR = _f32(np.concatenate([f["dR"] for f in frames]))
    P1c, P2c = _f64(P1).reshape(12), _f64(P2).reshape(12)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    assert seeds.shape == (nF, ransac_iter, 3)
    rec = np.zeros(nF, RECORD_DTYPE)
    poses = np.zeros((nF + 1, 4, 4)); npz = C.c_int32(0)
    rc = lib().vr_pipeline(nF, _p(nL), _p(nR), _p(offL), _p(offR), _p(kpL), _p(kpR), _p(dL), _p(dR), dL.shape[1], _p(P1c), _p(P2c),
                           int(ransac_iter), _p(seeds), _p(rec), _p(poses), C.byref(npz))
    assert rc == 0
    return dict(records=rec, poses=poses[:npz.value].copy())


def triangulate_dlt(x1, x2, P1, P2):
    x1, x2 = _f32(x1), _f32(x2)
    m = x1.shape[1]
    X = np.zeros((3, m), np.float32)
    P1c, P2c = _f64(P1).reshape(12), _f64(P2).reshape(12)
    lib().vr_triangulate_dlt(_p(x1), _p(x2), m, _p(P1c), _p(P2c), _p(X))
    return X


def triangulate_rectified_f32(x1, x2, f, base, c1u, c1v):
    x1, x2 = _f32(x1), _f32(x2)
    m = x1.shape[1]
    X = np.zeros((3, m), np.float32)
    lib().vr_triangulate_rectified_f32(_p(x1), _p(x2), m, _d(f), _d(base), _d(c1u), _d(c1v), _p(X))
    return X


def solve_rigid_motion(A, B):
    A, B = _f32(A), _f32(B)
    T = np.zeros((4, 4), np.float32)
    lib().vr_solve_rigid_motion(_p(A), _p(B), A.shape[1], _p(T))
    return T


# ---- the OpenCV routines restated in oracle/shim (pinned to the OpenCV-generated golden vectors) ----

def shim_radius_search(q, kp2, radius, K):
    kp2 = _f32(kp2)
    nb = np.empty(K, np.int32); d = np.empty(K, np.float32)
    total = lib().vr_shim_radius_search(float(q[0]), float(q[1]), _p(kp2), len(kp2), float(radius), K, _p(nb), _p(d))
    return total, nb, d


def shim_mul_transposed(J):
    J = _f64(J); out = np.zeros((6, 6))
    lib().vr_shim_mul_transposed(_p(J), J.shape[0], _p(out))
    return out


def shim_solve(A, b):
    A, b = _f64(A), _f64(b).reshape(-1); n = A.shape[0]; x = np.zeros(n)
    ok = lib().vr_shim_solve(_p(A), _p(b), n, _p(x))
    return bool(ok), x


def shim_invert(A):
    A = _f64(A); n = A.shape[0]; Ai = np.zeros((n, n))
    ok = lib().vr_shim_invert(_p(A), n, _p(Ai))
    return bool(ok), Ai


def shim_determinant(A):
    A = _f64(A)
    return lib().vr_shim_determinant(_p(A), A.shape[0])


def shim_sobel(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    out = np.zeros(img.shape, np.float32)
    lib().vr_shim_sobel(_p(img), img.shape[0], img.shape[1], _p(out))
    return out
