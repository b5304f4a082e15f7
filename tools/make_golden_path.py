#!/usr/bin/env python
"""Generate tests/golden/path_cv2.npz: the reference hot path executed with the REAL OpenCV primitives at every call
site the reference has one, on small synthetic inputs.

The reference (/root/reference) cannot be compiled here (no OpenCV / Eigen / Boost C++ headers), so this is the closest
thing to running it: a literal Python transcription of the control flow of viso.cpp in which each third-party call is
made through python-cv2 4.13 (the same library family), not restated:

    match_desc            viso.cpp:668-722   cv2.flann_Index(LINEAR, L1).radiusSearch (:181,684), cv2.norm(NORM_L1) (:702)
    sampsonDistance       viso.cpp:652-666, 390-407 (float / double mix as written)
    match_circle          viso.cpp:206-243   (four nested scans, literal)
    collect / triangulate viso.cpp:501-514, 1137-1154
    compute_J             viso.cpp:1401-1497
    minimize_reproj       viso.cpp:1583-1623 cv2.mulTransposed (:1599), cv2.gemm J^T r + cv2.solve(DECOMP_LU) (:1602)
    get_inliers           viso.cpp:1509-1537
    ransac_minimize_reproj viso.cpp:1543-1580 (sample table instead of randomsample)
    pose update           viso.cpp:1315-1321 cv2.invert + cv2.gemm

tests/test_oracle_golden.py::test_path_against_cv2_transcription replays the stored inputs through the C++ oracle and
compares: integer results exactly, tr to 1e-9 (cv2.gemm's J^T r accumulation order is build dependent).
std::sort (:724) is NOT covered here (python has no libstdc++): its order is pinned by tests/introsort_check.cpp.

  python tools/make_golden_path.py        (needs cv2; the output is committed)
"""
import math
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "path_cv2.npz")

DBL_MAX = sys.float_info.max


def sampson_distance(F, p1, p2):
    """viso.cpp:652-666 + algebricDistance :390-407"""
    p1x, p1y, p2x, p2y = (np.float32(v) for v in (p1[0], p1[1], p2[0], p2[1]))
    Fx0 = F[0, 0] * float(p1x) + F[0, 1] * float(p1y) + F[0, 2]
    Fx1 = F[1, 0] * float(p1x) + F[1, 1] * float(p1y) + F[1, 2]
    Ftx0 = F[0, 0] * float(p2x) + F[1, 0] * float(p2y) + F[2, 0]
    Ftx1 = F[0, 1] * float(p2x) + F[1, 1] * float(p2y) + F[2, 1]
    a = [float(p1x), float(p1y), 1.0]
    b = [float(p2x), float(p2y), 1.0]
    alg = (b[0] * F[0, 0] * a[0] + b[0] * F[0, 1] * a[1] + b[0] * F[0, 2] * a[2] +
           b[1] * F[1, 0] * a[0] + b[1] * F[1, 1] * a[1] + b[1] * F[1, 2] * a[2] +
           b[2] * F[2, 0] * a[0] + b[2] * F[2, 1] * a[1] + b[2] * F[2, 2] * a[2])
    ad = np.float32(alg)                      # float ad = algebricDistance(...)
    ad2 = np.float32(ad * ad)                 # ad*ad in float
    den = Fx0 * Fx0 + Fx1 * Fx1 + Ftx0 * Ftx0 + Ftx1 * Ftx1
    with np.errstate(all="ignore"):
        return float(np.float64(ad2) / np.float64(den))


def match_desc(kp1, kp2, d1, d2, sp):
    """viso.cpp:668-722 (without the final std::sort).  Returns the Match list in push order and the dense results."""
    kp1m, kp2m = np.ascontiguousarray(kp1, np.float32), np.ascontiguousarray(kp2, np.float32)
    n1, K = len(kp1m), sp["max_neighbors"]
    neighbors = -np.ones((n1, K), np.int32)                                   # :680-681
    if len(kp2m):
        index = cv2.flann_Index(kp2m, dict(algorithm=0), 2)                  # LinearIndexParams, L1 (:684)
        for i in range(n1):                                                   # radiusSearch wrapper :170-186
            found, ind, _ = index.radiusSearch(kp1m[i:i + 1], float(sp["radius"]), K)
            m = min(found, K)
            neighbors[i, :m] = ind[0, :m]
    matches = []
    dense = np.zeros((n1, 4), np.int64)
    for i in range(n1):                                                       # :686
        best_d1 = best_d2 = DBL_MAX
        best_idx = -1
        j = 0
        while j < K and neighbors[i, j] > 0:                                  # :692-693 (index 0 terminates)
            nind = int(neighbors[i, j])
            j += 1
            if sp["enforce_epipolar"]:
                s = sampson_distance(sp["F"], kp1m[i], kp2m[nind])
                if not math.isfinite(s) or s > sp["sampson_thresh"]:
                    continue
            d = cv2.norm(d2[nind:nind + 1] - d1[i:i + 1], cv2.NORM_L1)        # :702
            if d <= best_d1:
                best_d2, best_d1, best_idx = best_d1, d, nind
            elif d <= best_d2:
                best_d2 = d
        valid = 0
        if best_idx >= 0:
            if sp["enforce_2nd_best"]:
                if best_d1 < best_d2 * sp["ratio_2nd_best"]:
                    valid = 1
            else:
                valid = 1
            if valid:
                matches.append((i, best_idx, int(best_d1)))
        dense[i] = (best_idx, int(best_d1) if best_idx >= 0 else 2 ** 31 - 1,
                    int(best_d2) if best_d2 < DBL_MAX else 2 ** 31 - 1, valid)
    return np.array(matches, np.int32).reshape(-1, 3), dense


def match_circle(mlr, mlrp, m11, m22):
    """viso.cpp:206-243, literal nested scans"""
    circ, pcl = [], []
    for i in range(len(mlr)):
        il, ir = mlr[i][0], mlr[i][1]
        for j in range(len(m11)):
            if m11[j][0] != il:
                continue
            ilp = m11[j][1]
            for k in range(len(mlrp)):
                if mlrp[k][0] != ilp:
                    continue
                irp = mlrp[k][1]
                for l in range(len(m22)):
                    if m22[l][1] == irp and m22[l][0] == ir:
                        circ.append((il, ir, ilp, irp))
                        pcl.append((i, k))
    return np.array(circ, np.int32).reshape(-1, 4), np.array(pcl, np.int32).reshape(-1, 2)


def collect_matches(kp1, kp2, m):
    x = np.zeros((4, len(m)))
    for i, (a, b, _) in enumerate(m):
        x[0, i], x[1, i], x[2, i], x[3, i] = float(kp1[a][0]), float(kp1[a][1]), float(kp2[b][0]), float(kp2[b][1])
    return x


def triangulate_rectified(x, p):
    X = np.zeros((3, x.shape[1]))
    with np.errstate(all="ignore"):
        for i in range(x.shape[1]):
            d = np.float64(x[0, i]) - np.float64(x[2, i])
            X[0, i] = np.float64(p["base"]) * (x[0, i] - p["cu"]) / d
            X[1, i] = np.float64(p["base"]) * (x[1, i] - p["cv"]) / d
            X[2, i] = np.float64(p["f"] * p["base"]) / d
    return X


def compute_J(X, obs, tr, p, active):
    """viso.cpp:1401-1497 in IEEE doubles (python floats)"""
    rx, ry, rz, tx, ty, tz = (float(v) for v in tr)
    sx, cx, sy, cy, sz, cz = math.sin(rx), math.cos(rx), math.sin(ry), math.cos(ry), math.sin(rz), math.cos(rz)
    r00 = +cy * cz; r01 = -cy * sz; r02 = +sy
    r10 = +sx * sy * cz + cx * sz; r11 = -sx * sy * sz + cx * cz; r12 = -sx * cy
    r20 = -cx * sy * cz + sx * sz; r21 = +cx * sy * sz + sx * cz; r22 = +cx * cy
    rdrx10 = +cx * sy * cz - sx * sz; rdrx11 = -cx * sy * sz - sx * cz; rdrx12 = -cx * cy
    rdrx20 = +sx * sy * cz + cx * sz; rdrx21 = -sx * sy * sz + cx * cz; rdrx22 = -sx * cy
    rdry00 = -sy * cz; rdry01 = +sy * sz; rdry02 = +cy
    rdry10 = +sx * cy * cz; rdry11 = -sx * cy * sz; rdry12 = +sx * sy
    rdry20 = -cx * cy * cz; rdry21 = +cx * cy * sz; rdry22 = -cx * sy
    rdrz00 = -cy * sz; rdrz01 = -cy * cz
    rdrz10 = -sx * sy * sz + cx * cz; rdrz11 = -sx * sy * cz - cx * sz
    rdrz20 = +cx * sy * sz + sx * cz; rdrz21 = +cx * sy * cz - sx * sz
    na = len(active)
    J = np.zeros((4 * na, 6)); pred = np.zeros((4, na)); res = np.zeros((4 * na, 1))
    f, cu, cv, base = p["f"], p["cu"], p["cv"], p["base"]
    for i in range(na):
        a = active[i]
        X1p, Y1p, Z1p = float(X[0, a]), float(X[1, a]), float(X[2, a])
        X1c = r00 * X1p + r01 * Y1p + r02 * Z1p + tx
        Y1c = r10 * X1p + r11 * Y1p + r12 * Z1p + ty
        Z1c = r20 * X1p + r21 * Y1p + r22 * Z1p + tz
        weight = 1.0 / (abs(float(obs[0, i]) - cu) / abs(cu) + 0.05)          # column i, not active[i] (:1449)
        X2c = X1c - base
        for j in range(6):
            if j == 0:
                X1cd = 0.0; Y1cd = rdrx10 * X1p + rdrx11 * Y1p + rdrx12 * Z1p; Z1cd = rdrx20 * X1p + rdrx21 * Y1p + rdrx22 * Z1p
            elif j == 1:
                X1cd = rdry00 * X1p + rdry01 * Y1p + rdry02 * Z1p
                Y1cd = rdry10 * X1p + rdry11 * Y1p + rdry12 * Z1p
                Z1cd = rdry20 * X1p + rdry21 * Y1p + rdry22 * Z1p
            elif j == 2:
                X1cd = rdrz00 * X1p + rdrz01 * Y1p; Y1cd = rdrz10 * X1p + rdrz11 * Y1p; Z1cd = rdrz20 * X1p + rdrz21 * Y1p
            elif j == 3:
                X1cd, Y1cd, Z1cd = 1.0, 0.0, 0.0
            elif j == 4:
                X1cd, Y1cd, Z1cd = 0.0, 1.0, 0.0
            else:
                X1cd, Y1cd, Z1cd = 0.0, 0.0, 1.0
            J[4 * i + 0, j] = weight * f * (X1cd * Z1c - X1c * Z1cd) / (Z1c * Z1c)
            J[4 * i + 1, j] = weight * f * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c)
            J[4 * i + 2, j] = weight * f * (X1cd * Z1c - X2c * Z1cd) / (Z1c * Z1c)
            J[4 * i + 3, j] = weight * f * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c)
        pred[0, i] = f * X1c / Z1c + cu
        pred[1, i] = f * Y1c / Z1c + cv
        pred[2, i] = f * X2c / Z1c + cu
        pred[3, i] = f * Y1c / Z1c + cv
        for r in range(4):
            res[4 * i + r, 0] = weight * (float(obs[r, a]) - pred[r, i])
    return J, pred, res


def get_inliers(X, obs, tr, p):
    n = X.shape[1]
    _, pred, _ = compute_J(X, obs, tr, p, list(range(n)))
    thr2 = p["inlier_threshold"] * p["inlier_threshold"]
    inl = []
    for i in range(n):
        e = [float(obs[r, i]) - pred[r, i] for r in range(4)]
        err2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3]        # pow(x,2) == x*x
        if err2 < thr2:
            inl.append(i)
    return inl


def minimize_reproj(X, obs, tr, p, active):
    """viso.cpp:1583-1623 with the real cv::mulTransposed / gemm / solve"""
    tr = list(tr)
    for _ in range(100):
        J, _, res = compute_J(X, obs, tr, p, active)
        JtJ = cv2.mulTransposed(J, True)                                       # :1599
        Jtr = cv2.gemm(J, res, 1.0, None, 0.0, flags=cv2.GEMM_1_T)             # J.t()*residual
        ok, p_gn = cv2.solve(JtJ, Jtr, flags=cv2.DECOMP_LU)                    # :1602
        if not ok:
            return False, tr
        converged = True
        for j in range(6):
            if abs(float(p_gn[j, 0] > p["thresh"])):                           # sic, :1610
                converged = False
                break
        if converged:
            return True, tr
        for j in range(6):
            tr[j] = tr[j] + 1.0 * float(p_gn[j, 0])
    return False, tr


def ransac_minimize_reproj(X, obs, p, table, tr0):
    best_inl, best_tr = [], list(tr0)
    hyp_ok, hyp_cnt = [], []
    for h in range(p["ransac_iter"]):
        ok, tr = minimize_reproj(X, obs, [0.0] * 6, p, [int(v) for v in table[h]])
        hyp_ok.append(int(ok))
        if not ok:
            hyp_cnt.append(-1)
            continue
        inl = get_inliers(X, obs, tr, p)
        hyp_cnt.append(len(inl))
        if len(inl) > len(best_inl):
            best_inl, best_tr = inl, tr
    if len(best_inl) < 6:
        return False, best_tr, best_inl, hyp_ok, hyp_cnt
    ok, best_tr = minimize_reproj(X, obs, best_tr, p, best_inl)
    if not ok:
        return False, best_tr, best_inl, hyp_ok, hyp_cnt
    return True, best_tr, get_inliers(X, obs, best_tr, p), hyp_ok, hyp_cnt


def tr2mat(tr):
    rx, ry, rz = tr[:3]
    sx, cx, sy, cy, sz, cz = math.sin(rx), math.cos(rx), math.sin(ry), math.cos(ry), math.sin(rz), math.cos(rz)
    T = np.eye(4)
    T[0, :] = [+cy * cz, -cy * sz, +sy, tr[3]]
    T[1, :] = [+sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy, tr[4]]
    T[2, :] = [-cx * sy * cz + sx * sz, +cx * sy * sz + sx * cz, +cx * cy, tr[5]]
    return T


def main():
    from libviso_b200 import synth
    from oracle import oracle
    out = {}
    P1, P2 = synth.kitti_calib()
    rows = [(1, 2), (2, 0), (0, 1)]
    F = np.zeros((3, 3))
    for r in range(3):                                                         # F_from_P, mvg.h:41-66
        for c in range(3):
            F[r, c] = cv2.determinant(np.vstack([P1[rows[c][0]], P1[rows[c][1]], P2[rows[r][0]], P2[rows[r][1]]]))
    if F[2, 2] > sys.float_info.min:                                           # viso.cpp:1177-1180
        F = F / F[2, 2]
    p = dict(base=abs(P2[0, 3] / P2[0, 0]), f=P1[0, 0], cu=P1[0, 2], cv=P1[1, 2], inlier_threshold=2.0, thresh=1e-4,
             ransac_iter=12)
    stereo = dict(enforce_epipolar=True, F=F, sampson_thresh=1.0, enforce_2nd_best=False, ratio_2nd_best=.8,
                  max_neighbors=200, radius=80.0)
    temporal = dict(enforce_epipolar=False, F=None, sampson_thresh=0.0, enforce_2nd_best=True, ratio_2nd_best=.9,
                    max_neighbors=250, radius=80.0)
    frames, _ = synth.make_sequence(3, seed=1234, n_features=420)
    out["F"] = F
    out["P1"], out["P2"] = P1, P2
    for t, fr in enumerate(frames):
        for k in ("kpL", "kpR", "dL", "dR"):
            out[f"f{t}_{k}"] = fr[k]
    # ---- match_desc: the pipeline's three calls per frame + stress cases
    lr, xs, Xs = [], [], []
    for t, fr in enumerate(frames):
        m, dense = match_desc(fr["kpL"], fr["kpR"], fr["dL"], fr["dR"], stereo)
        out[f"lr{t}_push"], out[f"lr{t}_dense"] = m, dense
        m_sorted = oracle.sort_matches(m) if len(m) else m       # std::sort order: pinned separately (introsort_check)
        lr.append(m_sorted)
        xs.append(collect_matches(fr["kpL"], fr["kpR"], m_sorted))
        Xs.append(triangulate_rectified(xs[-1], p))
        out[f"x{t}"], out[f"X{t}"] = xs[-1], Xs[-1]
    seeds = np.random.default_rng(77).integers(0, 2 ** 32, size=(3, p["ransac_iter"], 3), dtype=np.uint32)
    out["seeds"] = seeds
    pose = np.eye(4)
    poses = [pose.copy()]
    for t in (1, 2):
        f, fp = frames[t], frames[t - 1]
        m11, d11 = match_desc(f["kpL"], fp["kpL"], f["dL"], fp["dL"], temporal)
        m22, d22 = match_desc(f["kpR"], fp["kpR"], f["dR"], fp["dR"], temporal)
        out[f"m11_{t}_dense"], out[f"m22_{t}_dense"] = d11, d22
        m11s, m22s = oracle.sort_matches(m11), oracle.sort_matches(m22)
        circ, pcl = match_circle(lr[t], lr[t - 1], m11s, m22s)
        out[f"circ{t}"], out[f"pcl{t}"] = circ, pcl
        C = len(circ)
        x_c = np.zeros((4, C)); Xp_c = np.zeros((3, C))
        for i in range(C):                                                     # viso.cpp:1291-1305
            x_c[:, i] = xs[t][:, pcl[i][0]]
            Xp_c[:, i] = Xs[t - 1][:, pcl[i][1]]
        table = oracle.samples_from_seeds(seeds[t], C)                         # integer mapping, checked in test_abi
        ok, tr, inl, hok, hcnt = ransac_minimize_reproj(Xp_c, x_c, p, table, [0.0] * 6)
        out[f"table{t}"], out[f"ok{t}"], out[f"tr{t}"] = table, np.int32(ok), np.array(tr)
        out[f"inl{t}"], out[f"hok{t}"], out[f"hcnt{t}"] = np.array(inl, np.int32), np.array(hok, np.int32), np.array(hcnt, np.int32)
        if ok:
            rv, Ti = cv2.invert(tr2mat(tr), flags=cv2.DECOMP_LU)               # :1319
            pose = cv2.gemm(pose, Ti, 1.0, None, 0.0)
            poses.append(pose.copy())
    out["poses"] = np.stack(poses)
    # ---- stress cases for match_desc: index-0 terminator, truncation (found > K), SAD ties, general F + ratio test
    rng = np.random.default_rng(5)
    kpa = rng.integers(0, 120, size=(150, 2)).astype(np.float32)
    kpb = rng.integers(0, 120, size=(160, 2)).astype(np.float32)
    da = rng.integers(-1020, 1021, size=(150, 121)).astype(np.float32)
    db = rng.integers(-1020, 1021, size=(160, 121)).astype(np.float32)
    db[40:60] = db[20:40]                                                      # identical descriptors: ties
    kpb[0] = kpa[3]                                                            # target 0 close to several queries
    Fg = rng.standard_normal((3, 3))
    cases = {
        "trunc": dict(temporal, max_neighbors=16),
        "mono": dict(enforce_epipolar=True, F=Fg, sampson_thresh=400.0, enforce_2nd_best=True, ratio_2nd_best=.9,
                     max_neighbors=250, radius=25.0),
        "plain": dict(temporal, enforce_2nd_best=False),
    }
    out["s_kpa"], out["s_kpb"], out["s_da"], out["s_db"], out["s_Fg"] = kpa, kpb, da, db, Fg
    for name, sp in cases.items():
        m, dense = match_desc(kpa, kpb, da, db, sp)
        out[f"s_{name}_push"], out[f"s_{name}_dense"] = m, dense
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: out[k].shape for k in ("lr1_push", "circ1", "inl1", "poses")}, "ok:", out["ok1"], out["ok2"])


if __name__ == "__main__":
    main()
