"""Synthetic KITTI-shaped stereo data for tests and bench.

Not part of the hot path: this is the *input* generator.  It renders 1241x376 8-bit gray stereo pairs of a
textured street canyon (ground plane + two side walls + ceiling, multi-octave band-limited noise texture) along a
known smooth trajectory with the KITTI-00 calibration from /root/reference/test/test.cpp:56-65, then runs a
Python/cv2 restatement of the reference front-end to obtain what the hot path consumes:

  * HarrisBinnedFeatureDetector (viso.cpp:911-979): cv::cornerHarris(block 3, aperture 5, k), 24x5 bins, the
    n/120 strongest |response| per bin.  The reference leaves m_k uninitialised (viso.cpp:915-919 vs :978); 0.04
    (the ctor default) is used.  std::nth_element's order inside a bin is implementation defined; the generator
    emits each bin's survivors by ascending response (ties: x, then y), bins in the reference's binx-major order.
  * MyFeatureExtractor (viso.cpp:981-1025): cv::Sobel(dx=1, ksize 3, BORDER_REFLECT_101) -> 11x11 patch ->
    121 float32 (integer valued, |v| <= 1020); pixels with row/col <= 0 or >= size read as 0 (viso.cpp:1018).
"""
import numpy as np

W, H = 1241, 376
F_PX, CU, CV = 718.856, 607.1928, 185.2157
P2_03 = -386.1448
BASE = abs(P2_03 / F_PX)


def kitti_calib():
    """P1, P2 (3x4 float64) -- KITTI-00, reference test/test.cpp:56-65."""
    P1 = np.array([[F_PX, 0, CU, 0], [0, F_PX, CV, 0], [0, 0, 1, 0]], np.float64)
    P2 = P1.copy()
    P2[0, 3] = P2_03
    return P1, P2


def make_texture(seed, size=2048):
    import cv2
    rng = np.random.default_rng(seed)
    tex = np.zeros((size, size), np.float32)
    for sigma, amp in ((1.5, 1.0), (4.0, 1.0), (10.0, 1.2), (25.0, 1.5)):
        n = rng.standard_normal((size, size)).astype(np.float32)
        # wrap-around blur so the texture tiles seamlessly
        b = cv2.GaussianBlur(np.tile(n, (3, 3))[size - 128:2 * size + 128, size - 128:2 * size + 128], (0, 0), sigma)
        b = b[128:128 + size, 128:128 + size]
        tex += amp * b / b.std()
    tex = (tex - tex.mean()) / tex.std()
    return np.clip(127.5 + 55.0 * tex, 0, 255).astype(np.float32)


def trajectory(n_frames, seed):
    """camera-to-world poses (R, p) per frame: forward 0.8-1.2 m/frame, smooth bounded yaw (|rate| < 0.01 rad/frame);
    the lateral offset stays within about +-1.5 m so the car remains inside the canyon."""
    rng = np.random.default_rng(seed)
    ph = rng.random(3) * 2 * np.pi
    t = np.arange(n_frames)
    yaw = 0.06 * np.sin(0.05 * t + ph[0]) + 0.02 * np.sin(0.3 * t + ph[1])
    speed = 1.0 + 0.2 * np.sin(t * 0.05 + ph[2])
    p = np.array([-1.2 * np.cos(ph[0]), 0.0, 0.0])
    poses = []
    for i in range(n_frames):
        yaw[i] -= 0.01 * p[0]  # gentle steering back to the canyon axis
        c, s = np.cos(yaw[i]), np.sin(yaw[i])
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        poses.append((R.copy(), p.copy()))
        p = p + R @ np.array([0.0, 0.0, speed[i]])
    return poses


def gt_motion(pose_prev, pose_cur):
    """4x4 T mapping previous-camera coordinates to current-camera coordinates (what tr2mat(tr) estimates)."""
    Rp, pp = pose_prev
    Rc, pc = pose_cur
    T = np.eye(4)
    T[:3, :3] = Rc.T @ Rp
    T[:3, 3] = Rc.T @ (pp - pc)
    return T


_PIX = None


def _pixel_dirs():
    global _PIX
    if _PIX is None:
        u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        _PIX = np.stack([(u - CU) / F_PX, (v - CV) / F_PX, np.ones_like(u)], 0).reshape(3, -1)
    return _PIX


def render(R, origin, tex, noise_rng=None):
    """ray-cast one 1241x376 view of the canyon from camera centre `origin` with rotation R (camera->world)."""
    import cv2
    size = tex.shape[0]
    d = (R @ _pixel_dirs()).astype(np.float32)
    o = np.asarray(origin, np.float32)
    npx = d.shape[1]
    best_s = np.full(npx, np.inf, np.float32)
    ta = np.zeros(npx, np.float32); tb = np.zeros(npx, np.float32)
    planes = (  # axis, offset, (texture axes), texture shift
        (1, 1.65, (0, 2), 0.0),     # ground  y = +1.65
        (0, -7.0, (2, 1), 300.0),   # left wall
        (0, +7.0, (2, 1), 900.0),   # right wall
        (1, -5.0, (0, 2), 1500.0),  # ceiling
    )
    # texture coordinates are taken modulo the tile size in float64 origin space to keep float32 precise
    period = size / 100.0
    with np.errstate(divide="ignore", invalid="ignore"):
        for axis, off, (a, b), shift in planes:
            s = (np.float32(off) - o[axis]) / d[axis]
            ok = (s > 0) & (s < best_s)
            np.copyto(best_s, s, where=ok)
            oa = np.float32(np.mod(float(origin[a]) + shift / 100.0, period))
            ob = np.float32(np.mod(float(origin[b]) + shift * 0.37 / 100.0, period))
            np.copyto(ta, (oa + s * d[a]) * np.float32(100.0), where=ok)
            np.copyto(tb, (ob + s * d[b]) * np.float32(100.0), where=ok)
    mx = np.mod(ta, np.float32(size)).reshape(H, W)
    my = np.mod(tb, np.float32(size)).reshape(H, W)
    img = cv2.remap(tex, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_WRAP)
    fog = np.clip((best_s.reshape(H, W) - np.float32(25.0)) / np.float32(50.0), 0, 1)
    img = img * (1 - fog) + np.float32(127.5) * fog
    if noise_rng is not None:
        img = img + noise_rng.standard_normal(img.shape, dtype=np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def detect_harris_binned(img, n, nbinx=24, nbiny=5, k=0.04, block_size=3, aperture_size=5):
    """viso.cpp:925-976 restated with cv2 (see module docstring for the order rule). Returns (N,2) float32 x,y."""
    import cv2
    resp = np.abs(cv2.cornerHarris(img, block_size, aperture_size, k, borderType=cv2.BORDER_DEFAULT))
    h, w = img.shape
    stridex, stridey = w // nbinx, h // nbiny
    per = n // (nbinx * nbiny)
    out = []
    for binx in range(nbinx):
        for biny in range(nbiny):
            x0, y0 = binx * stridex, biny * stridey
            blk = resp[y0:min(y0 + stridey, h), x0:min(x0 + stridex, w)]
            bh, bw = blk.shape
            # reference scan order: x outer, y inner
            vals = blk.T.reshape(-1)
            xs = np.repeat(np.arange(bw), bh) + x0
            ys = np.tile(np.arange(bh), bw) + y0
            nz = vals != 0
            vals, xs, ys = vals[nz], xs[nz], ys[nz]
            if len(vals) > per:
                order = np.lexsort((ys, xs, vals))[-per:]
            else:
                order = np.lexsort((ys, xs, vals))
            out.append(np.stack([xs[order], ys[order]], 1))
    return np.concatenate(out).astype(np.float32) if out else np.zeros((0, 2), np.float32)


def sobel_x(img):
    import cv2
    return cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=3, scale=1, delta=0, borderType=cv2.BORDER_REFLECT_101)


def extract_descriptors(img, kp, radius=5):
    """viso.cpp:1004-1024 (vectorised). Returns (N,(2r+1)^2) float32."""
    sob = sobel_x(img)
    h, w = img.shape
    px = np.rint(kp[:, 0]).astype(np.int64); py = np.rint(kp[:, 1]).astype(np.int64)  # Point2i p = kp.pt: cvRound
    offs = np.arange(-radius, radius + 1)
    yy = py[:, None, None] + offs[None, :, None]
    xx = px[:, None, None] + offs[None, None, :]
    ok = (yy > 0) & (yy < h) & (xx > 0) & (xx < w)
    d = np.where(ok, sob[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0))
    return np.ascontiguousarray(d.reshape(len(kp), -1), dtype=np.float32)


def make_features(imL, imR, n_features):
    """the reference's front end on one stereo pair (viso.cpp:1226-1231) with OpenCV's cornerHarris and Sobel"""
    kpL = detect_harris_binned(imL, n_features)
    kpR = detect_harris_binned(imR, n_features)
    return dict(kpL=kpL, kpR=kpR, dL=extract_descriptors(imL, kpL), dR=extract_descriptors(imR, kpR), imL=imL, imR=imR)


def make_frame(pose, tex, n_features, noise_rng):
    R, p = pose
    imL = render(R, p, tex, noise_rng)
    imR = render(R, p + R @ np.array([BASE, 0.0, 0.0]), tex, noise_rng)
    return make_features(imL, imR, n_features)


def _seq_worker(args):
    seed, t0, t1, n_features, n_frames = args
    tex = make_texture(seed)
    poses = trajectory(n_frames, seed + 7)
    out = []
    for t in range(t0, t1):
        out.append(make_frame(poses[t], tex, n_features, np.random.default_rng([seed, t])))
    return out


def make_sequence(n_frames, seed=1000, n_features=2040, workers=1):
    """returns (frames, gt) : frames = list of dict(kpL,kpR,dL,dR); gt[t] = 4x4 prev->cur motion (gt[0] = I)."""
    poses = trajectory(n_frames, seed + 7)
    if workers <= 1 or n_frames < 8:
        frames = _seq_worker((seed, 0, n_frames, n_features, n_frames))
    else:
        import multiprocessing as mp
        chunk = (n_frames + workers - 1) // workers
        jobs = [(seed, s, min(s + chunk, n_frames), n_features, n_frames) for s in range(0, n_frames, chunk)]
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            parts = pool.map(_seq_worker, jobs)
        frames = [f for part in parts for f in part]
    gt = [np.eye(4)] + [gt_motion(poses[t - 1], poses[t]) for t in range(1, n_frames)]
    return frames, gt


def make_dense_pair(n_kp=20000, seed=2000):
    """config 3 (matching-only stress): one rendered stereo pair, n_kp distinct random integer-pixel keypoints per
    image (bypassing the detector), descriptors = 11x11 Sobel patches."""
    tex = make_texture(seed)
    poses = trajectory(1, seed + 7)
    R, p = poses[0]
    rng = np.random.default_rng(seed + 1)
    imL = render(R, p, tex, rng)
    imR = render(R, p + R @ np.array([BASE, 0.0, 0.0]), tex, rng)

    def pick():
        flat = rng.choice(W * H, n_kp, replace=False)
        return np.stack([flat % W, flat // W], 1).astype(np.float32)
    kpL, kpR = pick(), pick()
    return dict(kpL=kpL, kpR=kpR, dL=extract_descriptors(imL, kpL), dR=extract_descriptors(imR, kpR))


def make_ransac_problem(n=10000, seed=3000, outlier_frac=0.3, noise_px=0.3):
    """config 4: n stereo correspondences under a known motion, 30% outliers. Returns X (3,n), obs (4,n), tr_true."""
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-2, 3, n), rng.uniform(4, 60, n)])
    tr = np.array([0.01, -0.02, 0.005, 0.05, -0.02, -1.0])
    rx, ry, rz = tr[:3]
    sx, cx, sy, cy, sz, cz = np.sin(rx), np.cos(rx), np.sin(ry), np.cos(ry), np.sin(rz), np.cos(rz)
    Rm = np.array([[cy * cz, -cy * sz, sy],
                   [sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy],
                   [-cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy]])
    Xc = Rm @ X + tr[3:, None]
    obs = np.stack([F_PX * Xc[0] / Xc[2] + CU, F_PX * Xc[1] / Xc[2] + CV,
                    F_PX * (Xc[0] - BASE) / Xc[2] + CU, F_PX * Xc[1] / Xc[2] + CV])
    obs += rng.standard_normal(obs.shape) * noise_px
    out = rng.random(n) < outlier_frac
    obs[:, out] += rng.uniform(-50, 50, (4, int(out.sum())))
    return np.ascontiguousarray(X), np.ascontiguousarray(obs), tr
