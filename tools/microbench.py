#!/usr/bin/env python
"""Micro-benchmarks for BASELINE.json configs[2] (matching-only stress, 20k keypoints per image) and configs[3]
(RANSAC / Gauss-Newton batch: 4096 hypotheses x 10k correspondences).  Run on the GPU box; prints one JSON object.

  config 3: one rendered stereo pair with 20 000 random integer-pixel keypoints per image, uploaded as `--copies`
            identical frames of a sequence so that one launch holds `copies` stereo jobs and 2*(copies-1) temporal
            jobs; sad_match time from CUDA events around the match launches; SAD-match GB/s per SURVEY 8d.
  config 4: viso_ransac_minimize_reproj through the C-ABI (host buffers), device time from the context stopwatch.
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--copies", type=int, default=9)
    ap.add_argument("--kp", type=int, default=20000)
    ap.add_argument("--hyp", type=int, default=4096)
    ap.add_argument("--points", type=int, default=10000)
    ap.add_argument("--check", action="store_true", help="compare with the CPU oracle (slow at full size)")
    args = ap.parse_args()
    from libviso_b200 import api, build, synth
    build.build()
    out = {}
    ctx = api.Context(0)
    ctx.set_image_extent(synth.W, synth.H)

    # ---- config 3
    pair = synth.make_dense_pair(args.kp, seed=2000)
    F = args.copies
    seq = ctx.sequence(F, args.kp, 121, 1)
    P1, P2 = synth.kitti_calib()
    seq.set_calib(P1, P2)
    for t in range(F):
        seq.upload_frame(t, pair["kpL"], pair["kpR"], pair["dL"], pair["dR"])
    seeds = np.zeros((F, 1, 3), np.uint32)
    prm = api.param_default(ransac_iter=1)
    seq.run(prm, seeds)
    ctx.sync()
    ms = []
    for _ in range(3):
        seq.run(prm)
        ms.append(seq.match_ms())
    mbytes, pairs, _ = seq.stats()
    n = args.kp
    jobs = 3 * F - 2
    b_call = (n + n) * (128 * 2 + 8) + 12 * n          # SURVEY 8d, u16 layout (264 B per keypoint, dense int4 is 16 B)
    t = float(np.median(ms)) * 1e-3
    out["config3_match_20k"] = {
        "keypoints_per_image": n, "jobs_per_launch": jobs, "sad_match_ms_per_launch": t * 1e3,
        "ms_per_match_desc_call": t * 1e3 / jobs, "sad_pairs_per_launch": int(pairs),
        "pairs_per_second": pairs / t, "algorithmic_bytes_per_call": b_call,
        "sad_match_GBps": jobs * b_call / t / 1e9, "hbm_frac_of_6553": jobs * b_call / t / 1e9 / 6553.0,
        "note": "every query has > max_neighbors points in range => exact top-K cut => generic kernel"}
    if args.check:
        from oracle import oracle
        F_ = oracle.F_from_P(P1, P2)
        o = oracle.match_desc(pair["kpL"], pair["kpR"], pair["dL"], pair["dR"], oracle.match_params_stereo(F_))
        g = seq.get_dense(0, 0)
        out["config3_match_20k"]["stereo_matches_equal_oracle"] = bool(
            np.array_equal(g[:, 0], o["idx"]) and np.array_equal(g[:, 1], o["d1"]) and np.array_equal(g[:, 3], o["valid"]))
    seq.close()

    # ---- config 4
    X, obs, tr_true = synth.make_ransac_problem(args.points, seed=3000)
    H = args.hyp
    table = api.randomsample_table(424242, H, args.points)
    p = api.param_default(base=synth.BASE, f=synth.F_PX, cu=synth.CU, cv=synth.CV, ransac_iter=H)
    r = ctx.ransac_minimize_reproj(X, obs, p, table)
    times = []
    for _ in range(5):
        ctx.sync(); t0 = time.perf_counter()
        ctx.timer_begin()
        r = ctx.ransac_minimize_reproj(X, obs, p, table)
        dev = ctx.timer_end()
        times.append((dev, 1e3 * (time.perf_counter() - t0)))
    dev_ms = float(np.median([a for a, _ in times])); wall_ms = float(np.median([b for _, b in times]))
    flops = 38.0 * H * args.points
    out["config4_ransac_batch"] = {
        "hypotheses": H, "correspondences": args.points, "device_ms": dev_ms, "call_wall_ms": wall_ms,
        "scoring_GFLOPs_f64": flops / (dev_ms * 1e-3) / 1e9, "ok": bool(r["ok"]), "n_inliers": int(len(r["inliers"])),
        "tr_error_rot": float(np.abs(r["tr"][:3] - tr_true[:3]).max()), "best_hyp": int(r["best_hyp"])}
    if args.check:
        from oracle import oracle
        po = oracle.param_default(base=synth.BASE, f=synth.F_PX, cu=synth.CU, cv=synth.CV, ransac_iter=H)
        t0 = time.perf_counter()
        o = oracle.ransac_minimize_reproj(X, obs, po, table)
        out["config4_ransac_batch"]["cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t0)
        out["config4_ransac_batch"]["inliers_equal_oracle"] = bool(np.array_equal(o["inliers"], r["inliers"]))
        dc = np.abs(o["hyp_count"].astype(np.int64) - r["hyp_count"])
        out["config4_ransac_batch"]["hyp_counts_differing"] = int((dc > 0).sum())   # knife-edge points: device sin/cos vs glibc
        out["config4_ransac_batch"]["hyp_count_max_abs_diff"] = int(dc.max())
        out["config4_ransac_batch"]["hyp_ok_equal_oracle"] = bool(np.array_equal(o["hyp_ok"], r["hyp_ok"]))
        out["config4_ransac_batch"]["best_hyp_equal_oracle"] = bool(o["best_hyp"] == r["best_hyp"])
        out["config4_ransac_batch"]["tr_max_abs_diff"] = float(np.abs(o["tr"] - r["tr"]).max())
    print(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
