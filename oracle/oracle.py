"""ctypes binding for the CPU oracle (oracle/viso_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (libviso_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("VISO_ORACLE_SO") or os.path.join(_HERE, "libviso_oracle.so")  # override: tools/cpu_baseline.py (-O0 build)


def build(force=False):
    src = os.path.join(_HERE, "viso_oracle.cpp")
    hdr = os.path.join(_HERE, "viso_oracle.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B", os.path.basename(_SO)], stdout=subprocess.DEVNULL)
    return _SO


class MatchParams(C.Structure):
    _fields_ = [("enforce_epipolar", C.c_int32), ("enforce_2nd_best", C.c_int32),
                ("max_neighbors", C.c_int32), ("_pad", C.c_int32),
                ("radius", C.c_double), ("sampson_thresh", C.c_double),
                ("ratio_2nd_best", C.c_double), ("F", C.c_double * 9)]


class Param(C.Structure):
    _fields_ = [("base", C.c_double), ("f", C.c_double), ("cu", C.c_double), ("cv", C.c_double),
                ("inlier_threshold", C.c_double), ("thresh", C.c_double),
                ("ransac_iter", C.c_int32), ("_pad", C.c_int32)]


class Record(C.Structure):
    _fields_ = [("tr", C.c_double * 6), ("ok", C.c_int32), ("n_inliers", C.c_int32),
                ("n_circ", C.c_int32), ("best_hyp", C.c_int32)]


RECORD_DTYPE = np.dtype([("tr", np.float64, 6), ("ok", np.int32), ("n_inliers", np.int32),
                         ("n_circ", np.int32), ("best_hyp", np.int32)])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.vo_sampson_distance.restype = C.c_double
        _lib.vo_sampson_distance.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
        _lib.vo_determinant.restype = C.c_double
        _lib.vo_radius_search.argtypes = [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_int,
                                          C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def match_params_stereo(F):
    p = MatchParams()
    Fc = _f64(F).reshape(9)
    lib().vo_match_params_stereo(C.byref(p), _p(Fc))
    return p


def match_params_temporal():
    p = MatchParams()
    lib().vo_match_params_temporal(C.byref(p))
    return p


def param_default(base=0.0, f=0.0, cu=0.0, cv=0.0, ransac_iter=50):
    p = Param()
    lib().vo_param_default(C.byref(p))
    p.base, p.f, p.cu, p.cv, p.ransac_iter = base, f, cu, cv, ransac_iter
    return p


def sampson_distance(F, p1, p2):
    Fc = _f64(F).reshape(9)
    return lib().vo_sampson_distance(_p(Fc), float(p1[0]), float(p1[1]), float(p2[0]), float(p2[1]))


def radius_search(q, kp2, radius, K):
    kp2 = _f32(kp2)
    nb = np.empty(K, np.int32)
    d = np.empty(K, np.float32)
    total = lib().vo_radius_search(float(q[0]), float(q[1]), _p(kp2), len(kp2), float(radius), K, _p(nb), _p(d))
    return total, nb, d


def match_desc(kp1, kp2, d1, d2, sp):
    """returns dict(matches [M,3] sorted like the reference, idx/d1/d2/valid dense per query, n_sad)"""
    kp1, kp2, d1, d2 = _f32(kp1), _f32(kp2), _f32(d1), _f32(d2)
    n1, n2 = len(kp1), len(kp2)
    dlen = d1.shape[1] if d1.ndim == 2 else d2.shape[1]
    matches = np.zeros((max(n1, 1), 3), np.int32)
    nm = C.c_int32(0)
    di = np.zeros(max(n1, 1), np.int32); dd1 = np.zeros(max(n1, 1), np.int32)
    dd2 = np.zeros(max(n1, 1), np.int32); dv = np.zeros(max(n1, 1), np.int32)
    ns = C.c_int64(0)
    rc = lib().vo_match_desc(_p(kp1), n1, _p(kp2), n2, _p(d1), _p(d2), dlen, C.byref(sp),
                             _p(matches), C.byref(nm), _p(di), _p(dd1), _p(dd2), _p(dv), C.byref(ns))
    assert rc == 0
    return dict(matches=matches[:nm.value].copy(), idx=di[:n1], d1=dd1[:n1], d2=dd2[:n1], valid=dv[:n1],
                n_sad=ns.value)


def sort_matches(m):
    m = _i32(m).copy()
    lib().vo_sort_matches(_p(m), len(m))
    return m


def match_circle(mlr, mlrp, m11, m22):
    mlr, mlrp, m11, m22 = _i32(mlr), _i32(mlrp), _i32(m11), _i32(m22)
    cap = max(len(mlr), 1)
    circ = np.zeros((cap, 4), np.int32)
    pcl = np.zeros((cap, 3), np.int32)
    c = lib().vo_match_circle(_p(mlr), len(mlr), _p(mlrp), len(mlrp), _p(m11), len(m11), _p(m22), len(m22),
                              _p(circ), _p(pcl))
    return circ[:c].copy(), pcl[:c].copy()


def collect_matches(kp1, kp2, matches):
    kp1, kp2, matches = _f32(kp1), _f32(kp2), _i32(matches)
    m = len(matches)
    x = np.zeros((4, m), np.float64)
    lib().vo_collect_matches(_p(kp1), _p(kp2), _p(matches), m, _p(x))
    return x


def triangulate_rectified_f64(x, f, base, cu, cv):
    x = _f64(x)
    m = x.shape[1]
    X = np.zeros((3, m), np.float64)
    lib().vo_triangulate_rectified_f64(_p(x), m, C.c_double(f), C.c_double(base), C.c_double(cu), C.c_double(cv), _p(X))
    return X


def triangulate_rectified_f32(x1, x2, f, base, c1u, c1v):
    x1, x2 = _f32(x1), _f32(x2)
    m = x1.shape[1]
    X = np.zeros((3, m), np.float32)
    lib().vo_triangulate_rectified_f32(_p(x1), _p(x2), m, C.c_double(f), C.c_double(base), C.c_double(c1u),
                                       C.c_double(c1v), _p(X))
    return X


def compute_J(X, obs, tr, param, active):
    X, obs, tr, active = _f64(X), _f64(obs), _f64(tr), _i32(active)
    n, na = X.shape[1], len(active)
    J = np.zeros((4 * na, 6)); pred = np.zeros((4, na)); res = np.zeros(4 * na)
    lib().vo_compute_J(_p(X), _p(obs), n, _p(tr), C.byref(param), _p(active), na, _p(J), _p(pred), _p(res))
    return J, pred, res


def get_inliers(X, obs, tr, param):
    X, obs, tr = _f64(X), _f64(obs), _f64(tr)
    n = X.shape[1]
    inl = np.zeros(max(n, 1), np.int32)
    rms = C.c_double(0); mg = C.c_double(0)
    c = lib().vo_get_inliers(_p(X), _p(obs), n, _p(tr), C.byref(param), _p(inl), C.byref(rms), C.byref(mg))
    return inl[:c].copy(), rms.value, mg.value


def mul_transposed(J):
    J = _f64(J); out = np.zeros((6, 6))
    lib().vo_mul_transposed(_p(J), J.shape[0], _p(out))
    return out


def Jt_times_r(J, r):
    J, r = _f64(J), _f64(r); out = np.zeros(6)
    lib().vo_Jt_times_r(_p(J), _p(r), J.shape[0], _p(out))
    return out


def minimize_reproj(X, obs, tr, param, active):
    X, obs, active = _f64(X), _f64(obs), _i32(active)
    tr = _f64(tr).copy()
    it = C.c_int32(0)
    ok = lib().vo_minimize_reproj(_p(X), _p(obs), X.shape[1], _p(tr), C.byref(param), _p(active), len(active),
                                  C.byref(it))
    return bool(ok), tr, it.value


def ransac_minimize_reproj(X, obs, param, table, tr0=None):
    X, obs, table = _f64(X), _f64(obs), _i32(table)
    n = X.shape[1]
    H = param.ransac_iter
    assert table.shape == (H, 3)
    tr = np.zeros(6) if tr0 is None else _f64(tr0).copy()
    inl = np.zeros(max(n, 1), np.int32)
    nb = C.c_int32(0); bh = C.c_int32(-1)
    htr = np.zeros((H, 6)); hok = np.zeros(H, np.int32); hc = np.zeros(H, np.int32)
    ok = lib().vo_ransac_minimize_reproj(_p(X), _p(obs), n, C.byref(param), _p(table), _p(tr), _p(inl),
                                         C.byref(nb), _p(htr), _p(hok), _p(hc), C.byref(bh))
    return dict(ok=bool(ok), tr=tr, inliers=inl[:nb.value].copy(), hyp_tr=htr, hyp_ok=hok, hyp_count=hc,
                best_hyp=bh.value)


def randomsample_table(seed, H, N):
    t = np.zeros((H, 3), np.int32)
    lib().vo_randomsample_table(C.c_uint32(seed), H, N, _p(t))
    return t


def samples_from_seeds(seeds, N):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    H = seeds.shape[0]
    t = np.zeros((H, 3), np.int32)
    lib().vo_samples_from_seeds(_p(seeds), H, N, _p(t))
    return t


def tr2mat(tr):
    tr = _f64(tr); T = np.zeros((4, 4))
    lib().vo_tr2mat(_p(tr), _p(T))
    return T


def invert_lu(A):
    A = _f64(A); n = A.shape[0]; Ai = np.zeros((n, n))
    ok = lib().vo_invert_lu(_p(A), n, _p(Ai))
    return bool(ok), Ai


def solve_lu(A, b):
    A, b = _f64(A), _f64(b).reshape(-1); n = A.shape[0]; x = np.zeros(n)
    ok = lib().vo_solve_lu(_p(A), _p(b), n, _p(x))
    return bool(ok), x


def determinant(A):
    A = _f64(A)
    return lib().vo_determinant(_p(A), A.shape[0])


def F_from_P(P1, P2, normalise=True):
    P1, P2 = _f64(P1).reshape(12), _f64(P2).reshape(12); F = np.zeros((3, 3))
    lib().vo_F_from_P(_p(P1), _p(P2), int(normalise), _p(F))
    return F


def pose_update(pose, tr):
    pose, tr = _f64(pose), _f64(tr); out = np.zeros((4, 4))
    ok = lib().vo_pose_update(_p(pose), _p(tr), _p(out))
    return bool(ok), out


def triangulate_dlt(x1, x2, P1, P2):
    x1, x2, P1, P2 = _f32(x1), _f32(x2), _f64(P1).reshape(12), _f64(P2).reshape(12)
    m = x1.shape[1]; X = np.zeros((3, m), np.float32)
    lib().vo_triangulate_dlt(_p(x1), _p(x2), m, _p(P1), _p(P2), _p(X))
    return X


def solve_rigid_motion(A, B):
    A, B = _f32(A), _f32(B); T = np.zeros((4, 4), np.float32)
    lib().vo_solve_rigid_motion(_p(A), _p(B), A.shape[1], _p(T))
    return T


def project_points(X, P):
    X, P = _f64(X), _f64(P).reshape(12); n = X.shape[1]; x = np.zeros((2, n))
    rc = lib().vo_project_points(_p(X), n, _p(P), _p(x))
    if rc != 0:
        raise OverflowError("divide by zero in h2e")
    return x


def extract_descriptors(sob, kp, radius=5):
    sob, kp = _f32(sob), _f32(kp)
    h, w = sob.shape; n = len(kp); side = 2 * radius + 1
    d = np.zeros((n, side * side), np.float32)
    lib().vo_extract_descriptors(_p(sob), h, w, _p(kp), n, radius, _p(d))
    return d


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def sobel_x(img):
    img = _u8(img); h, w = img.shape
    out = np.zeros((h, w), np.float32)
    lib().vo_sobel_x(_p(img), h, w, _p(out))
    return out


def harris_response(img, k=0.04):
    img = _u8(img); h, w = img.shape
    out = np.zeros((h, w), np.float32)
    lib().vo_harris_response(_p(img), h, w, C.c_float(k), _p(out))
    return out


def detect_harris_binned(img, n, nbinx=24, nbiny=5, k=0.04, order_rule=1, with_response=False):
    img = _u8(img); h, w = img.shape
    xy = np.zeros((max(n, 1), 2), np.float32); rs = np.zeros(max(n, 1), np.float32)
    cnt = lib().vo_detect_harris_binned(_p(img), h, w, n, nbinx, nbiny, C.c_float(k), order_rule, _p(xy), _p(rs))
    assert cnt >= 0
    return (xy[:cnt].copy(), rs[:cnt].copy()) if with_response else xy[:cnt].copy()


def frames_from_images(images, n_features=2040, k=0.04):
    """images: list of (imL, imR) uint8.  The front end of viso.cpp:1208-1222: detect, then describe."""
    frames = []
    for imL, imR in images:
        kpL = detect_harris_binned(imL, n_features, k=k); kpR = detect_harris_binned(imR, n_features, k=k)
        frames.append(dict(kpL=kpL, kpR=kpR, dL=extract_descriptors(sobel_x(imL), kpL),
                           dR=extract_descriptors(sobel_x(imR), kpR), imL=imL, imR=imR))
    return frames


def sequence(frames, P1, P2, param, seeds, dump=False):
    """frames: list of dict(kpL, kpR, dL, dR).  seeds: [n_frames, H, 3] uint32.
    Returns dict(records, poses, and (dump=True) lr_matches, lr_count, m11, m22, circ, inliers as per-frame lists)."""
    nF = len(frames)
    nL = np.array([len(f["kpL"]) for f in frames], np.int32)
    nR = np.array([len(f["kpR"]) for f in frames], np.int32)
    offL = np.zeros(nF, np.int64); offR = np.zeros(nF, np.int64)
    offL[1:] = np.cumsum(nL)[:-1]; offR[1:] = np.cumsum(nR)[:-1]
    kpL = _f32(np.concatenate([f["kpL"] for f in frames])); kpR = _f32(np.concatenate([f["kpR"] for f in frames]))
    dL = _f32(np.concatenate([f["dL"] for f in frames])); dR = _f32(np.concatenate([f["dR"] for f in frames]))
    dlen = dL.shape[1]
    P1c, P2c = _f64(P1).reshape(12), _f64(P2).reshape(12)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    assert seeds.shape == (nF, param.ransac_iter, 3)
    rec = np.zeros(nF, RECORD_DTYPE)
    poses = np.zeros((nF + 1, 4, 4)); npz = C.c_int32(0)
    totL, totR = int(nL.sum()), int(nR.sum())
    if dump:
        lrm = np.zeros((max(totL, 1), 3), np.int32); lrc = np.zeros(nF, np.int32)
        m11 = np.zeros((max(totL, 1), 4), np.int32); m22 = np.zeros((max(totR, 1), 4), np.int32)
        circ = np.zeros((max(totL, 1), 4), np.int32); inl = np.zeros(max(totL, 1), np.int32)
    else:
        lrm = lrc = m11 = m22 = circ = inl = None
    rc = lib().vo_sequence(nF, _p(nL), _p(nR), _p(offL), _p(offR), _p(kpL), _p(kpR), _p(dL), _p(dR), dlen,
                           _p(P1c), _p(P2c), C.byref(param), _p(seeds), _p(rec),
                           _p(lrm), _p(lrc), _p(m11), _p(m22), _p(circ), _p(inl), _p(poses), C.byref(npz))
    assert rc == 0
    out = dict(records=rec, poses=poses[:npz.value].copy())
    if dump:
        out["lr_matches"] = [lrm[offL[t]:offL[t] + lrc[t]].copy() for t in range(nF)]
        out["m11"] = [m11[offL[t]:offL[t] + nL[t]].copy() for t in range(nF)]
        out["m22"] = [m22[offR[t]:offR[t] + nR[t]].copy() for t in range(nF)]
        out["circ"] = [circ[offL[t]:offL[t] + rec["n_circ"][t]].copy() for t in range(nF)]
        out["inliers"] = [inl[offL[t]:offL[t] + rec["n_inliers"][t]].copy() for t in range(nF)]
    return out
