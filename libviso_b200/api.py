"""ctypes binding of the C-ABI in include/viso_b200.h (libviso_b200/libviso_b200.so).

This is the Python face of the product: the parity tests, smoke() and bench.py call the CUDA path through it.
Function names and argument meaning mirror the reference's free functions (src/viso.h, src/mvg.h):
match_desc, match_circle, triangulate_rectified, minimize_reproj, ransac_minimize_reproj, tr2mat, ...
There is NO CPU fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("VISO_B200_LIB") or os.path.join(_HERE, "libviso_b200.so")   # override: tuning builds only

VISO_OK = 0
ERR_NAMES = {-1: "VISO_ERR_CUDA", -2: "VISO_ERR_ARG", -3: "VISO_ERR_DOMAIN", -4: "VISO_ERR_DIV0",
             -5: "VISO_ERR_DUPLICATE", -6: "VISO_ERR_NOMEM"}


class VisoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class MatchParams(C.Structure):
    """MatchParams, reference src/viso.cpp:48-75"""
    _fields_ = [("enforce_epipolar", C.c_int32), ("enforce_2nd_best", C.c_int32),
                ("max_neighbors", C.c_int32), ("_pad", C.c_int32),
                ("radius", C.c_double), ("sampson_thresh", C.c_double),
                ("ratio_2nd_best", C.c_double), ("F", C.c_double * 9)]


class Param(C.Structure):
    """struct param, reference src/viso.h:58-72"""
    _fields_ = [("base", C.c_double), ("f", C.c_double), ("cu", C.c_double), ("cv", C.c_double),
                ("inlier_threshold", C.c_double), ("thresh", C.c_double),
                ("ransac_iter", C.c_int32), ("_pad", C.c_int32)]


RECORD_DTYPE = np.dtype([("tr", np.float64, 6), ("ok", np.int32), ("n_inliers", np.int32),
                         ("n_circ", np.int32), ("best_hyp", np.int32)])

_lib = None


def lib():
    """load the C-ABI library; fails loudly when it has not been built (no fallback)"""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -m libviso_b200.build` (needs nvcc). "
                               "libviso_b200 has no CPU fallback.")
        _lib = C.CDLL(SO_PATH)
        _lib.viso_last_error.restype = C.c_char_p
        _lib.viso_last_error.argtypes = [C.c_void_p]
        _lib.viso_stream.restype = C.c_void_p
        _lib.viso_stream.argtypes = [C.c_void_p]
        _lib.viso_launch_count.restype = C.c_int64
        _lib.viso_launch_count.argtypes = [C.c_void_p]
        _lib.viso_destroy.restype = None
        _lib.viso_destroy.argtypes = [C.c_void_p]
        _lib.viso_seq_destroy.restype = None
        _lib.viso_seq_destroy.argtypes = [C.c_void_p]
        _lib.viso_seq_records_device.restype = C.c_void_p
        _lib.viso_seq_records_device.argtypes = [C.c_void_p]
        _lib.viso_seq_device_bytes.restype = C.c_int64
        _lib.viso_seq_device_bytes.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw address (e.g. a pinned torch tensor's data_ptr())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(v):
    return C.c_double(float(v))


def match_params_stereo(F):
    p = MatchParams()
    Fc = _f64(F).reshape(9)
    lib().viso_match_params_stereo(C.byref(p), _p(Fc))
    return p


def match_params_temporal():
    p = MatchParams()
    lib().viso_match_params_temporal(C.byref(p))
    return p


def param_default(base=0.0, f=0.0, cu=0.0, cv=0.0, ransac_iter=50):
    p = Param()
    lib().viso_param_default(C.byref(p))
    p.base, p.f, p.cu, p.cv, p.ransac_iter = base, f, cu, cv, ransac_iter
    return p


# ---- host bookkeeping (no device needed) ----

def tr2mat(tr):
    tr = _f64(tr); T = np.zeros((4, 4))
    lib().viso_tr2mat(_p(tr), _p(T))
    return T


def F_from_P(P1, P2, normalise=True):
    P1, P2 = _f64(P1).reshape(12), _f64(P2).reshape(12); F = np.zeros((3, 3))
    lib().viso_F_from_P(_p(P1), _p(P2), int(normalise), _p(F))
    return F


def pose_update(pose, tr):
    pose, tr = _f64(pose), _f64(tr); out = np.zeros((4, 4))
    rc = lib().viso_pose_update(_p(pose), _p(tr), _p(out))
    return rc == 0, out


def randomsample_table(seed, H, N):
    t = np.zeros((H, 3), np.int32)
    lib().viso_randomsample_table(C.c_uint32(seed), H, N, _p(t))
    return t


def samples_from_seeds(seeds, N):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    t = np.zeros((seeds.shape[0], 3), np.int32)
    lib().viso_samples_from_seeds(_p(seeds), seeds.shape[0], N, _p(t))
    return t


def chain_poses(records):
    rec = np.ascontiguousarray(records, dtype=RECORD_DTYPE)
    poses = np.zeros((len(rec) + 1, 4, 4))
    n = lib().viso_chain_poses(_p(rec), len(rec), _p(poses))
    return poses[:n].copy()


class Context:
    """viso_ctx: one CUDA device + stream.  Raises when no GPU is present (there is no CPU path)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib().viso_create(C.byref(h), int(device))
        if rc != 0:
            raise VisoError(rc, "viso_create failed: no usable CUDA device (libviso_b200 has no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib().viso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise VisoError(rc, lib().viso_last_error(self.h).decode())

    def share_copy_stream(self, owner):
        """uploads of this context's sequences are queued behind `owner`'s (first-in first-out over one PCIe link)"""
        self._ck(lib().viso_share_copy_stream(self.h, owner.h if owner is not None else None))

    def sync(self):
        self._ck(lib().viso_sync(self.h))

    def stream_ptr(self):
        """the cudaStream_t every kernel of this context is launched on"""
        return int(lib().viso_stream(self.h) or 0)

    def set_image_extent(self, w, h):
        self._ck(lib().viso_set_image_extent(self.h, int(w), int(h)))

    MATCH_MODES = {"auto": 0, "generic": 1, "gather": 2, "staged": 3}

    def set_match_mode(self, mode):
        """force one kernel path of match_desc ("auto", "generic", "gather", "staged"); results are identical"""
        self._ck(lib().viso_set_match_mode(self.h, self.MATCH_MODES[mode] if isinstance(mode, str) else int(mode)))

    def launch_count(self):
        return int(lib().viso_launch_count(self.h))

    def timer_begin(self):
        self._ck(lib().viso_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float(0)
        self._ck(lib().viso_timer_end(self.h, C.byref(ms)))
        return ms.value

    # ---- match_desc, viso.cpp:668-726 ----
    def match_desc_dense(self, kp1, kp2, d1, d2, sp):
        """dense per-query (best_idx, best_d1, best_d2, valid) before compaction / sort"""
        kp1, kp2, d1, d2 = _f32(kp1).reshape(-1, 2), _f32(kp2).reshape(-1, 2), _f32(d1), _f32(d2)
        n1, n2 = len(kp1), len(kp2)
        dlen = d1.shape[1] if (d1.ndim == 2 and n1) else d2.shape[1]
        idx = np.zeros(n1, np.int32); b1 = np.zeros(n1, np.int32); b2 = np.zeros(n1, np.int32); v = np.zeros(n1, np.int32)
        self._ck(lib().viso_match_desc(self.h, _p(kp1), n1, _p(kp2), n2, _p(d1), _p(d2), dlen, C.byref(sp),
                                       _p(idx), _p(b1), _p(b2), _p(v)))
        return dict(idx=idx, d1=b1, d2=b2, valid=v)

    def match_desc(self, kp1, kp2, d1, d2, sp):
        """Matches [M,3] = (i1, i2, dist) in the reference's std::sort order (viso.cpp:724)"""
        kp1, kp2, d1, d2 = _f32(kp1).reshape(-1, 2), _f32(kp2).reshape(-1, 2), _f32(d1), _f32(d2)
        n1, n2 = len(kp1), len(kp2)
        dlen = d1.shape[1] if (d1.ndim == 2 and n1) else d2.shape[1]
        m = np.zeros((max(n1, 1), 3), np.int32)
        nm = C.c_int32(0)
        self._ck(lib().viso_match_desc_sorted(self.h, _p(kp1), n1, _p(kp2), n2, _p(d1), _p(d2), dlen, C.byref(sp),
                                              _p(m), C.byref(nm)))
        return m[:nm.value].copy()

    # ---- std::sort by dist, viso.cpp:724 ----
    def sort_matches(self, m):
        m = np.ascontiguousarray(_i32(m).reshape(-1, 3)).copy()
        self._ck(lib().viso_sort_matches(self.h, _p(m), len(m)))
        return m

    # ---- triangulate_dlt (mvg.cpp:124-169), solveRigidMotion (estimation.cpp:29-51) ----
    def triangulate_dlt(self, x1, x2, P1, P2):
        x1, x2, P1, P2 = _f32(x1), _f32(x2), _f64(P1).reshape(12), _f64(P2).reshape(12)
        m = x1.shape[1]; X = np.zeros((3, m), np.float32)
        self._ck(lib().viso_triangulate_dlt(self.h, _p(x1), _p(x2), m, _p(P1), _p(P2), _p(X)))
        return X

    def solve_rigid_motion(self, A, B):
        A, B = _f32(A), _f32(B); T = np.zeros((4, 4), np.float32)
        self._ck(lib().viso_solve_rigid_motion(self.h, _p(A), _p(B), A.shape[1], _p(T)))
        return T

    # ---- HarrisBinnedFeatureDetector::detectImpl, viso.cpp:925-976 ----
    def detect_harris(self, img, n_features, nbinx=24, nbiny=5, k=0.04, with_response=False):
        img = np.ascontiguousarray(img, dtype=np.uint8); h, w = img.shape
        xy = np.zeros((max(n_features, 1), 2), np.float32); rs = np.zeros(max(n_features, 1), np.float32)
        n = C.c_int32(0)
        self._ck(lib().viso_detect_harris(self.h, _p(img), w, h, w, int(n_features), int(nbinx), int(nbiny), C.c_float(k),
                                          _p(xy), _p(rs), C.byref(n)))
        return (xy[:n.value].copy(), rs[:n.value].copy()) if with_response else xy[:n.value].copy()

    # ---- MyFeatureExtractor::computeImpl, viso.cpp:1004-1024 ----
    def extract_descriptors(self, img, kp):
        img = np.ascontiguousarray(img, dtype=np.uint8); h, w = img.shape
        kp = _f32(kp).reshape(-1, 2); d = np.zeros((len(kp), 121), np.float32)
        self._ck(lib().viso_extract_descriptors(self.h, _p(img), w, h, w, _p(kp), len(kp), _p(d)))
        return d

    # ---- match_circle, viso.cpp:206-243 ----
    def match_circle(self, mlr, mlrp, m11, m22):
        mlr, mlrp, m11, m22 = (_i32(a).reshape(-1, 3) for a in (mlr, mlrp, m11, m22))
        cap = max(len(mlr), 1)
        circ = np.zeros((cap, 4), np.int32); pcl = np.zeros((cap, 3), np.int32)
        c = C.c_int32(0)
        self._ck(lib().viso_match_circle(self.h, _p(mlr), len(mlr), _p(mlrp), len(mlrp), _p(m11), len(m11),
                                         _p(m22), len(m22), _p(circ), _p(pcl), C.byref(c)))
        return circ[:c.value].copy(), pcl[:c.value].copy()

    # ---- collect_matches + triangulate_rectified<double>, viso.cpp:501-514, 1137-1162 ----
    def collect_triangulate(self, kp1, kp2, matches, f, base, cu, cv):
        kp1, kp2, matches = _f32(kp1).reshape(-1, 2), _f32(kp2).reshape(-1, 2), _i32(matches).reshape(-1, 3)
        m = len(matches)
        x = np.zeros((4, m)); X = np.zeros((3, m))
        self._ck(lib().viso_collect_triangulate(self.h, _p(kp1), len(kp1), _p(kp2), len(kp2), _p(matches), m,
                                                _d(f), _d(base), _d(cu), _d(cv), _p(x), _p(X)))
        return x, X

    def triangulate_rectified_f64(self, x, f, base, cu, cv):
        x = _f64(x); m = x.shape[1]; X = np.zeros((3, m))
        self._ck(lib().viso_triangulate_rectified_f64(self.h, _p(x), m, _d(f), _d(base), _d(cu), _d(cv), _p(X)))
        return X

    def triangulate_rectified_f32(self, x1, x2, f, base, c1u, c1v):
        """mvg.cpp:172-192"""
        x1, x2 = _f32(x1), _f32(x2); m = x1.shape[1]; X = np.zeros((3, m), np.float32)
        self._ck(lib().viso_triangulate_rectified_f32(self.h, _p(x1), _p(x2), m, _d(f), _d(base), _d(c1u), _d(c1v),
                                                      _p(X)))
        return X

    def project_points(self, X, P):
        """viso.cpp:326-333; raises OverflowError like the reference's h2e (misc.h:118-119)"""
        X, P = _f64(X), _f64(P).reshape(12); n = X.shape[1]; x = np.zeros((2, n))
        rc = lib().viso_project_points(self.h, _p(X), n, _p(P), _p(x))
        if rc == -4:
            raise OverflowError("divide by zero in h2e")
        self._ck(rc)
        return x

    # ---- estimation, viso.cpp:1401-1623 ----
    def get_inliers(self, X, obs, tr, param):
        X, obs, tr = _f64(X), _f64(obs), _f64(tr); n = X.shape[1]
        inl = np.zeros(max(n, 1), np.int32); c = C.c_int32(0)
        self._ck(lib().viso_get_inliers(self.h, _p(X), _p(obs), n, _p(tr), C.byref(param), _p(inl), C.byref(c)))
        return inl[:c.value].copy()

    def minimize_reproj(self, X, obs, tr, param, active):
        X, obs, active = _f64(X), _f64(obs), _i32(active)
        tr = _f64(tr).copy(); ok = C.c_int32(0)
        self._ck(lib().viso_minimize_reproj(self.h, _p(X), _p(obs), X.shape[1], _p(tr), C.byref(param), _p(active),
                                            len(active), C.byref(ok)))
        return bool(ok.value), tr

    def ransac_minimize_reproj(self, X, obs, param, table, tr0=None):
        X, obs, table = _f64(X), _f64(obs), _i32(table)
        n = X.shape[1]; H = param.ransac_iter
        assert table.shape == (H, 3)
        tr = np.zeros(6) if tr0 is None else _f64(tr0).copy()
        inl = np.zeros(max(n, 1), np.int32)
        ni = C.c_int32(0); ok = C.c_int32(0); bh = C.c_int32(-1)
        htr = np.zeros((max(H, 1), 6)); hok = np.zeros(max(H, 1), np.int32); hc = np.zeros(max(H, 1), np.int32)
        self._ck(lib().viso_ransac_minimize_reproj(self.h, _p(X), _p(obs), n, C.byref(param), _p(table), _p(tr),
                                                   _p(inl), C.byref(ni), C.byref(ok), _p(htr), _p(hok), _p(hc),
                                                   C.byref(bh)))
        return dict(ok=bool(ok.value), tr=tr, inliers=inl[:ni.value].copy(), hyp_tr=htr[:H], hyp_ok=hok[:H],
                    hyp_count=hc[:H], best_hyp=bh.value)

    def set_hyp_iteration_cap(self, cap):
        self._ck(lib().viso_set_hyp_iteration_cap(self.h, int(cap)))

    def debug_sincos(self, x):
        """sin / cos as the estimation kernels evaluate them (glibc's algorithm on the device)"""
        x = _f64(x).reshape(-1)
        s = np.empty_like(x); c = np.empty_like(x)
        self._ck(lib().viso_debug_sincos(self.h, _p(x), len(x), _p(s), _p(c)))
        return s, c

    def sequence(self, n_frames, max_kp, desc_len=121, max_ransac_iter=50):
        return Sequence(self, n_frames, max_kp, desc_len, max_ransac_iter)


class Sequence:
    """viso_seq: the per-frame loop of sequence_odometry (viso.cpp:1205-1327) for all frames of a sequence at once."""

    def __init__(self, ctx, n_frames, max_kp, desc_len=121, max_ransac_iter=50):
        self.ctx = ctx
        self.n_frames, self.max_kp, self.desc_len, self.max_H = n_frames, max_kp, desc_len, max_ransac_iter
        h = C.c_void_p()
        ctx._ck(lib().viso_seq_create(ctx.h, n_frames, max_kp, desc_len, max_ransac_iter, C.byref(h)))
        self.h = h
        self._keep = []

    def close(self):
        if self.h:
            lib().viso_seq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_calib(self, P1, P2):
        P1, P2 = _f64(P1).reshape(12), _f64(P2).reshape(12)
        self.ctx._ck(lib().viso_seq_set_calib(self.h, _p(P1), _p(P2)))

    def upload_frame(self, t, kpL, kpR, dL, dR):
        kpL, kpR, dL, dR = _f32(kpL).reshape(-1, 2), _f32(kpR).reshape(-1, 2), _f32(dL), _f32(dR)
        self.ctx._ck(lib().viso_seq_upload_frame(self.h, t, _p(kpL), len(kpL), _p(kpR), len(kpR), _p(dL), _p(dR)))
        self.ctx.sync()  # the numpy temporaries are pageable and may die

    def upload_frame_raw(self, t, kpL_ptr, nL, kpR_ptr, nR, dL_ptr, dR_ptr):
        """same with raw host addresses (pinned buffers): the copies are then truly asynchronous"""
        self.ctx._ck(lib().viso_seq_upload_frame(self.h, t, _p(kpL_ptr), nL, _p(kpR_ptr), nR, _p(dL_ptr), _p(dR_ptr)))

    def upload(self, frames):
        for t, f in enumerate(frames):
            self.upload_frame(t, f["kpL"], f["kpR"], f["dL"], f["dR"])

    # ---- device front-end: MyFeatureExtractor (viso.cpp:1004-1024) from the 8-bit images ----
    def set_image_size(self, width, height):
        self.ctx._ck(lib().viso_seq_set_image_size(self.h, int(width), int(height)))

    def upload_frame_images(self, t, imL, imR, kpL, kpR):
        imL, imR = np.ascontiguousarray(imL, dtype=np.uint8), np.ascontiguousarray(imR, dtype=np.uint8)
        kpL, kpR = _f32(kpL).reshape(-1, 2), _f32(kpR).reshape(-1, 2)
        self.ctx._ck(lib().viso_seq_upload_frame_images(self.h, t, _p(imL), _p(imR), _p(kpL), len(kpL), _p(kpR), len(kpR)))
        self.ctx.sync()  # the numpy temporaries are pageable and may die

    def upload_frame_images_raw(self, t, imL_ptr, imR_ptr, kpL_ptr, nL, kpR_ptr, nR):
        self.ctx._ck(lib().viso_seq_upload_frame_images(self.h, t, _p(imL_ptr), _p(imR_ptr), _p(kpL_ptr), nL, _p(kpR_ptr), nR))

    def capacity(self):
        return int(lib().viso_seq_capacity(self.h))

    def records_device_ptr(self):
        """device address of the n_frames 64-byte records (for consumers that stay on the GPU, e.g. an NCCL gather)"""
        return int(lib().viso_seq_records_device(self.h))

    def device_bytes(self):
        return int(lib().viso_seq_device_bytes(self.h))

    def upload_chunk_images_raw(self, t0, count, images_ptr, kpL_ptr, nL_ptr, kpR_ptr, nR_ptr):
        """count frames in three copies from (pinned) host blocks: images [count][2][H][W] u8, kp [count][capacity()][2] f32"""
        self.ctx._ck(lib().viso_seq_upload_chunk_images(self.h, int(t0), int(count), _p(images_ptr), _p(kpL_ptr), _p(nL_ptr),
                                                        _p(kpR_ptr), _p(nR_ptr)))

    def upload_images(self, frames):
        for t, f in enumerate(frames):
            self.upload_frame_images(t, f["imL"], f["imR"], f["kpL"], f["kpR"])

    # ---- device detector + extractor: only the images are uploaded ----
    def set_detector(self, n_features, nbinx=24, nbiny=5, k=0.04):
        self.ctx._ck(lib().viso_seq_set_detector(self.h, int(n_features), int(nbinx), int(nbiny), C.c_float(k)))

    def upload_frame_raw_images(self, t, imL, imR):
        imL, imR = np.ascontiguousarray(imL, dtype=np.uint8), np.ascontiguousarray(imR, dtype=np.uint8)
        self.ctx._ck(lib().viso_seq_upload_frame_raw(self.h, int(t), _p(imL), _p(imR)))
        self.ctx.sync()  # the numpy temporaries are pageable and may die

    def upload_frame_raw_ptr(self, t, imL_ptr, imR_ptr):
        self.ctx._ck(lib().viso_seq_upload_frame_raw(self.h, int(t), _p(imL_ptr), _p(imR_ptr)))

    def upload_chunk_raw(self, t0, count, images_ptr):
        """count frames in one block from (pinned) host memory: images [count][2][H][W] u8"""
        self.ctx._ck(lib().viso_seq_upload_chunk_raw(self.h, int(t0), int(count), _p(images_ptr)))

    def get_keypoints(self, t, side):
        out = np.zeros((self.max_kp + 32, 2), np.float32); n = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_get_keypoints(self.h, int(t), int(side), _p(out), C.byref(n)))
        return out[:n.value].copy()

    def run_range(self, param, t0, t1):
        self.ctx._ck(lib().viso_seq_run_range(self.h, C.byref(param), int(t0), int(t1)))

    def get_packed(self, t, side):
        out = np.zeros((self.max_kp + 32, 128), np.uint16); n = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_get_packed(self.h, t, side, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def set_seeds(self, seeds, ransac_iter):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert seeds.size == self.n_frames * ransac_iter * 3
        self.ctx._ck(lib().viso_seq_set_seeds(self.h, _p(seeds), ransac_iter))

    def run(self, param, seeds=None):
        if seeds is not None:
            self.set_seeds(seeds, param.ransac_iter)
        self.ctx._ck(lib().viso_seq_run_resident(self.h, C.byref(param)))

    def download(self, out=None):
        rec = np.zeros(self.n_frames, RECORD_DTYPE) if out is None else out
        self.ctx._ck(lib().viso_seq_download(self.h, _p(rec)))
        return rec

    def download_raw(self, ptr):
        self.ctx._ck(lib().viso_seq_download(self.h, _p(ptr)))

    def stats(self):
        mb = C.c_int64(0); sp = C.c_int64(0); se = C.c_int64(0)
        self.ctx._ck(lib().viso_seq_stats(self.h, C.byref(mb), C.byref(sp), C.byref(se)))
        return mb.value, sp.value, se.value

    def time_match(self, which):
        """device ms of a launch of only the stereo (0) / temporal (1) match jobs; returns (ms, sad_pairs, n_pending)"""
        ms = C.c_float(0); sp = C.c_int64(0); npend = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_time_match(self.h, int(which), C.byref(ms), C.byref(sp), C.byref(npend)))
        return ms.value, sp.value, npend.value

    def last_pending(self):
        n = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_last_pending(self.h, C.byref(n)))
        return n.value

    def match_ms(self):
        ms = C.c_float(0)
        self.ctx._ck(lib().viso_seq_match_ms(self.h, C.byref(ms)))
        return ms.value

    def get_dense(self, which, t):
        out = np.zeros((self.max_kp + 32, 4), np.int32); n = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_get_dense(self.h, which, t, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def get_lr_matches(self, t):
        out = np.zeros((self.max_kp + 32, 3), np.int32); n = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_get_lr_matches(self.h, t, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def get_circ(self, t):
        c4 = np.zeros((self.max_kp + 32, 4), np.int32); p2 = np.zeros((self.max_kp + 32, 2), np.int32); n = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_get_circ(self.h, t, _p(c4), _p(p2), C.byref(n)))
        return c4[:n.value].copy(), p2[:n.value].copy()

    def get_inliers(self, t):
        out = np.zeros(self.max_kp + 32, np.int32); n = C.c_int32(0)
        self.ctx._ck(lib().viso_seq_get_inliers(self.h, t, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def get_hyp(self, t, H):
        htr = np.zeros((H, 6)); hok = np.zeros(H, np.int32); hc = np.zeros(H, np.int32)
        self.ctx._ck(lib().viso_seq_get_hyp(self.h, t, _p(htr), _p(hok), _p(hc)))
        return htr, hok, hc
