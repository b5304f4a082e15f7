#!/usr/bin/env python
"""Generate tests/golden/*.npz: known-answer vectors that PIN the oracle's third-party arithmetic.

The reference (/root/reference) cannot be compiled here and ships no golden vectors, so
the vectors come from the one implementation of the reference's third-party dependencies that IS
present: OpenCV (python cv2 4.13).  Each block below calls the cv2 primitive the reference calls on the
hot path and stores inputs + outputs.  tests/test_oracle_golden.py replays them through the oracle and
requires bit equality (except where noted).  Run here (needs cv2); the outputs are committed.

  python tools/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def flann_cases():
    """cvflann linear L1 radiusSearch: ordering (dist, idx), inclusive radius, total count (viso.cpp:181,684)."""
    rng = np.random.default_rng(11)
    cases = {}
    specs = [
        ("int_dense", rng.integers(0, 120, size=(900, 2)).astype(np.float32), 80.0, 200),   # truncation + ties
        ("int_sparse", rng.integers(0, 1241, size=(400, 2)).astype(np.float32), 80.0, 250),  # found < K
        ("float", (rng.random((700, 2)) * 150).astype(np.float32), 80.0, 200),               # float L1 sums
        ("dup", np.repeat(rng.integers(0, 60, size=(50, 2)), 6, axis=0).astype(np.float32), 40.0, 64),
    ]
    for name, pts, radius, K in specs:
        index = cv2.flann_Index(pts, dict(algorithm=0), 2)  # FLANN_INDEX_LINEAR, FLANN_DIST_L1
        qs = pts[rng.choice(len(pts), 24, replace=False)].copy()
        qs[:4] += np.float32(0.5)
        founds, idxs, dists = [], [], []
        for q in qs:
            ret, ind, d = index.radiusSearch(q.reshape(1, 2), radius, K)
            n = min(ret, K)
            row = np.full(K, -1, np.int32); drow = np.full(K, -1, np.float32)
            row[:n] = ind[0, :n]; drow[:n] = d[0, :n]
            founds.append(ret); idxs.append(row); dists.append(drow)
        cases[name + "_pts"] = pts
        cases[name + "_q"] = qs
        cases[name + "_radius"] = np.float32(radius)
        cases[name + "_K"] = np.int32(K)
        cases[name + "_found"] = np.array(founds, np.int32)
        cases[name + "_idx"] = np.stack(idxs)
        cases[name + "_dist"] = np.stack(dists)
    np.savez_compressed(os.path.join(OUT, "flann_radius.npz"), **cases)


def linalg_cases():
    """cv::mulTransposed (viso.cpp:1599), cv::solve LU (:1602), Mat::inv (:1319), cv::determinant (mvg.h:62-64)."""
    rng = np.random.default_rng(12)
    out = {}
    Js = [rng.standard_normal((12, 6)) * 50, rng.standard_normal((4000, 6)) * 10, rng.standard_normal((28, 6))]
    for i, J in enumerate(Js):
        out[f"mt_J{i}"] = J
        out[f"mt_JtJ{i}"] = cv2.mulTransposed(J, True)
    As, bs, xs, oks = [], [], [], []
    for i in range(40):
        J = rng.standard_normal((12 if i % 2 else 60, 6)) * (10.0 ** rng.integers(-2, 3))
        A = cv2.mulTransposed(J, True)
        if i % 10 == 9:
            A[:, 3] = A[:, 2]; A[3, :] = A[2, :]  # singular
        b = rng.standard_normal((6, 1)) * 100
        ok, x = cv2.solve(A, b, flags=cv2.DECOMP_LU)
        As.append(A); bs.append(b[:, 0]); xs.append(x[:, 0]); oks.append(int(ok))
    out["lu_A"] = np.stack(As); out["lu_b"] = np.stack(bs); out["lu_x"] = np.stack(xs); out["lu_ok"] = np.array(oks)
    Ts, Tis = [], []
    for i in range(20):
        M = rng.standard_normal((4, 4))
        if i < 10:  # rigid-motion shaped
            a = rng.standard_normal(3) * 0.1
            Rx = cv2.Rodrigues(a)[0]
            M = np.eye(4); M[:3, :3] = Rx; M[:3, 3] = rng.standard_normal(3)
        rv, Mi = cv2.invert(M, flags=cv2.DECOMP_LU)
        Ts.append(M); Tis.append(Mi)
    out["inv_A"] = np.stack(Ts); out["inv_Ai"] = np.stack(Tis)
    Ds, dets = [], []
    for i in range(20):
        M = rng.standard_normal((4, 4)) * 100
        Ds.append(M); dets.append(cv2.determinant(M))
    out["det_A"] = np.stack(Ds); out["det"] = np.array(dets)
    # KITTI-00 projection matrices (test.cpp:56-65) -> F via 9 cv2.determinant calls (mvg.h:41-66)
    P1 = np.array([[718.856, 0, 607.1928, 0], [0, 718.856, 185.2157, 0], [0, 0, 1, 0]], np.float64)
    P2 = P1.copy(); P2[0, 3] = -386.1448
    rows = [(1, 2), (2, 0), (0, 1)]
    F = np.zeros((3, 3))
    for r in range(3):
        for c in range(3):
            M = np.vstack([P1[rows[c][0]], P1[rows[c][1]], P2[rows[r][0]], P2[rows[r][1]]])
            F[r, c] = cv2.determinant(M)
    out["F_P1"] = P1; out["F_P2"] = P2; out["F_raw"] = F
    # 4x4 pose product pose * inv(T) through cv2.gemm (viso.cpp:1319)
    pose = np.eye(4); poses = []
    for i in range(10):
        pose = cv2.gemm(pose, Tis[i], 1.0, None, 0.0)
        poses.append(pose.copy())
    out["pose_chain"] = np.stack(poses)
    np.savez_compressed(os.path.join(OUT, "linalg.npz"), **out)


def sobel_case():
    """cv::Sobel(image, CV_32F, 1, 0, 3, 1, 0, BORDER_REFLECT_101) (viso.cpp:1010) on a small random image."""
    rng = np.random.default_rng(13)
    img = rng.integers(0, 256, size=(48, 64), dtype=np.uint8)
    sob = cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=3, scale=1, delta=0, borderType=cv2.BORDER_REFLECT_101)
    np.savez_compressed(os.path.join(OUT, "sobel.npz"), img=img, sob=sob)


def harris_case():
    """cv::cornerHarris(image, 3, 5, 0.04, BORDER_DEFAULT) (viso.cpp:930) on a crop of a synthetic frame plus a random
    image, and the binned detector (viso.cpp:925-976, canonical order rule) applied to cv2's response.  cornerHarris is
    not bit-reproducible (see oracle/viso_oracle.h): the oracle is compared with a tolerance."""
    from libviso_b200 import synth
    tex = synth.make_texture(77, 1024)
    R = np.eye(3); pos = np.array([0.0, 0.0, 0.0])
    full = synth.render(R, pos, tex, np.random.default_rng(78))
    rng = np.random.default_rng(14)
    out = {}
    for name, img, (nx, ny, n) in (("crop", np.ascontiguousarray(full[100:220, 300:460]), (4, 3, 96)),
                                   ("noise", rng.integers(0, 256, size=(75, 102), dtype=np.uint8), (2, 1, 40))):
        resp = cv2.cornerHarris(img, 3, 5, 0.04, borderType=cv2.BORDER_DEFAULT)
        out[name + "_img"] = img
        out[name + "_resp"] = resp
        out[name + "_bins"] = np.array([nx, ny, n], np.int32)
        a = np.abs(resp); h, w = img.shape; sx, sy = w // nx, h // ny; per = n // (nx * ny)
        kps = []
        for bx in range(nx):
            for by in range(ny):
                blk = a[by * sy:(by + 1) * sy, bx * sx:(bx + 1) * sx]
                vals = blk.T.reshape(-1); xs = np.repeat(np.arange(sx), sy) + bx * sx; ys = np.tile(np.arange(sy), sx) + by * sy
                nz = vals != 0
                vals, xs, ys = vals[nz], xs[nz], ys[nz]
                o = np.lexsort((ys, xs, vals))[-per:]
                kps.append(np.stack([xs[o], ys[o]], 1))
        out[name + "_kp"] = np.concatenate(kps).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "harris.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    flann_cases()
    linalg_cases()
    harris_case()
    sobel_case()
    print("wrote", sorted(os.listdir(OUT)))
