#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the libviso hot path (match + circle + triangulate + RANSAC pose) on B200.

A "step" is one pass of the whole per-frame loop of sequence_odometry (reference src/viso.cpp:1205-1327, minus
detection / description / debug output) over every synthetic KITTI-shaped 1241x376 stereo sequence a rank owns.

  1 GPU (default)   BASELINE.json configs[1]: ONE sequence of 1000 frames, ~2k features per image, 50 RANSAC hypotheses
                    per frame pair.
  N > 1 (default)   BASELINE.json configs[4]: EIGHT independent sequences x 1025 frames (8192 frame pairs) sharded over
                    the ranks, sequence s on rank s mod N; the 64-byte per-frame-pair records are gathered to rank 0 with
                    one NCCL gather per step INSIDE both timed regions, straight from the device record buffers.  The
                    total work is fixed for N = 2, 4, 8 ("strong"); --sequences S selects the same mode explicitly
                    (e.g. --gpus 1 --sequences 8 --frames 1025).

  value  device throughput: inputs (8-bit images + keypoints) already resident in HBM when the timed region starts;
         K steps timed with CUDA events on the library's stream (the NCCL gather is enqueued on that stream)
  e2e    the same through the C-ABI with HOST buffers: every step copies all frames' images + keypoints + sample seeds
         from pinned host memory, reads the 64-byte records back, and (N > 1) gathers them
  roofline      the sad_match kernel: algorithmic bytes (per frame pair, u16 layout) / its CUDA-event time
  cpu_baseline  the CPU reference path on a bounded sample of the same sequence, 1 core
  extra         (1 GPU) BASELINE configs[2] (20k-keypoint matching, stereo and temporal mode) and configs[3] (4096
                hypotheses x 10k correspondences) measured through the same library, with their parity booleans

`--impl reference` times the reference's CPU path on the host cores, one process per core over independent frame
ranges: the oracle port ("kind": "port": identical results to the reference's own compiled code, tests/test_ref_pin.py,
and several times faster than that code built against the header stand-ins of compat/ -- the harder baseline);
VISO_CPU_BASELINE=reference times oracle/_ref/libviso_ref.so instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frame-pairs/s (match+RANSAC pose) at 1241x376"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sequences", type=int, default=0,
                    help="independent sequences in total, sharded s mod N over the ranks (0: 1 on one GPU = configs[1], "
                         "8 under torchrun = configs[4])")
    ap.add_argument("--frames", type=int, default=0, help="frames per sequence (0: 1000, or 1025 in the 8-sequence mode)")
    ap.add_argument("--features", type=int, default=2040)
    ap.add_argument("--hyp", type=int, default=50)
    ap.add_argument("--unique", type=int, default=256,
                    help="distinct rendered frames per sequence; the sequence drives back and forth over them (every frame "
                         "has its own HBM copy, so the working set is the full --frames)")
    ap.add_argument("--seed", type=int, default=1000, help="sequence s is rendered with seed --seed + s")
    ap.add_argument("--cpu-pairs", type=int, default=384, help="frame pairs in the cpu_baseline sample (about 16 s of CPU work; every record of the sample is compared with the GPU run)")
    ap.add_argument("--chunk", type=int, default=0,
                    help="frames per upload/compute chunk of the e2e pipeline (0: 125, or 250 with --input raw -- measured)")
    ap.add_argument("--e2e-buffers", type=int, default=2, choices=[1, 2],
                    help="sequence objects (each on its own context / streams) taking alternate e2e work items")
    ap.add_argument("--input", default="images", choices=["images", "descriptors", "raw"],
                    help="what crosses the boundary per frame: the two 8-bit images + keypoints (descriptors extracted "
                         "on the device, viso.cpp:1004-1024), the reference's n x 121 f32 descriptor matrices, or the images only")
    ap.add_argument("--e2e-separate-copy-streams", action="store_true",
                    help="A/B switch: every e2e lane uploads on its own copy stream (pieces of the two lanes interleave)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] / configs[3] measurements")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.mode = "configs[1]"
    if args.sequences <= 0:
        args.sequences = 1 if world == 1 else 8
        if world > 1:
            args.mode = "configs[4]"
    elif args.sequences > 1:
        args.mode = "configs[4]"
    args.sequences = max(args.sequences, world)
    if args.frames <= 0:
        args.frames = 1025 if args.mode == "configs[4]" else 1000
    if args.chunk <= 0:
        args.chunk = 250 if args.input == "raw" else 125
    return args


def pingpong(n_frames, n_unique):
    """frame t of the long sequence -> index into the rendered frames: 0,1,..,U-1,U-2,..,0,1,.."""
    if n_unique <= 1:
        return np.zeros(n_frames, np.int64)
    period = 2 * (n_unique - 1)
    t = np.arange(n_frames) % period
    return np.where(t < n_unique, t, period - t)


def build_sequence(args, seed, workers):
    from libviso_b200 import synth
    n_unique = min(args.unique, args.frames)
    frames, _ = synth.make_sequence(n_unique, seed=seed, n_features=args.features, workers=workers)
    return frames, pingpong(args.frames, n_unique)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=10)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(args):
    """DRAM bytes of one sad_match launch from the committed ncu capture of this exact config (else None)"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["sad_match_kernel"]
        if t["frames"] == args.frames and t["features"] == args.features:
            return int(t["dram_bytes_read"] + t["dram_bytes_write"])
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ CPU reference path

def cpu_backend():
    """The CPU arm that is TIMED is the oracle port: on the same inputs it returns exactly what the reference's own code
    returns (tests/test_ref_pin.py) and it is ~9x FASTER than oracle/_ref (the reference compiled against this
    repository's header stand-ins for OpenCV, whose cv::Mat arithmetic is scalar and allocation-heavy), i.e. the more
    demanding baseline.  VISO_CPU_BASELINE=reference times oracle/_ref instead ("kind": "reference")."""
    from oracle import oracle
    oracle.lib()
    if os.environ.get("VISO_CPU_BASELINE", "port") != "reference":
        return "port", oracle
    try:
        from oracle import ref
        if os.path.exists(ref.SO) or ref.available():
            ref.lib()
            return "reference", ref
    except Exception as e:  # noqa: BLE001 -- any failure to build / load falls back to the restated port
        print(f"[bench] oracle/_ref unavailable ({e}); timing the oracle port", file=sys.stderr)
    return "port", oracle


def cpu_pairs(kind, mod, frames, order, t0, n_pairs, H, seeds, front_end=0):
    """the CPU path over frames [t0, t0 + n_pairs] of the long sequence; returns (seconds, records)"""
    from oracle import oracle
    from libviso_b200 import synth
    P1, P2 = synth.kitti_calib()
    sub = [frames[order[t]] for t in range(t0, t0 + n_pairs + 1)]
    sd = np.ascontiguousarray(seeds[t0:t0 + n_pairs + 1])
    t = time.perf_counter()
    if front_end:
        # --input raw: the reference's front end per frame too (viso.cpp:1226-1231), with OpenCV doing what the
        # reference asks OpenCV to do (cv::cornerHarris, cv::Sobel)
        sub = [synth.make_features(f["imL"], f["imR"], front_end) for f in sub]
    if kind == "reference":
        out = mod.pipeline(sub, P1, P2, H, sd)
    else:
        out = oracle.sequence(sub, P1, P2, oracle.param_default(ransac_iter=H), sd)
    return time.perf_counter() - t, out["records"]


def make_seeds(n_frames, H, seed):
    return np.random.default_rng(424242 + seed).integers(0, 2 ** 32, size=(n_frames, H, 3), dtype=np.uint32)


_REF_STATE = None


def _ref_step(job):
    t0, n_pairs = job
    kind, mod, frames, order, seeds, H, front_end = _REF_STATE
    return cpu_pairs(kind, mod, frames, order, t0, n_pairs, H, seeds, front_end)[0]


def run_reference(args, rank, guard):
    """--impl reference: the reference's CPU path on the host cores"""
    if rank != 0:
        return
    import multiprocessing as mp
    kind, mod = cpu_backend()   # loaded in the parent: the forked workers inherit it
    cores = os.cpu_count() or 1
    frames, order = build_sequence(args, args.seed, min(cores, 16))
    seeds = make_seeds(args.frames, args.hyp, args.seed)
    pairs_per_proc = 6
    jobs_per_step = cores
    global _REF_STATE
    _REF_STATE = (kind, mod, frames, order, seeds, args.hyp, args.features if args.input == "raw" else 0)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(i):
            starts = [((i * jobs_per_step + j) * pairs_per_proc) % max(args.frames - pairs_per_proc - 1, 1) for j in range(jobs_per_step)]
            t = time.perf_counter()
            pool.map(_ref_step, [(s, pairs_per_proc) for s in starts])
            return time.perf_counter() - t
        for i in range(args.warmup):
            step(i)
        times = [step(args.warmup + i) for i in range(args.steps)]
    total = sum(times)
    n_pairs = args.steps * jobs_per_step * pairs_per_proc
    value = n_pairs / total
    what = ("the reference's own match_desc / match_circle / triangulate_rectified / ransac_minimize_reproj, compiled from "
            "its sources (oracle/_ref/libviso_ref.so, g++ -O2)") if kind == "reference" else "the oracle port (g++ -O2)"
    sample = (f"{jobs_per_step} processes x {pairs_per_proc} frame pairs per step, independent frame ranges of one "
              f"{args.frames}-frame sequence; {what}")
    if args.input == "raw":
        sample += "; front end per frame = cv2.cornerHarris + cv2.Sobel + numpy binning / patch gather (libviso_b200/synth.py)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "frame-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if (args.mode == "configs[4]" and args.gpus > 1) else "weak", "vs_baseline": None,
        "dtype": "u16 SAD / f64 pose", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    guard.emit(json.dumps(line))


def workload_config(args, world):
    per = [len([s for s in range(args.sequences) if s % world == r]) for r in range(world)]
    if args.mode == "configs[4]":
        wl = (f"{args.sequences} independent synthetic KITTI-shaped stereo sequences x {args.frames} frames "
              f"({args.sequences * (args.frames - 1)} frame pairs, BASELINE configs[4]) sharded s mod {world} over {world} GPU(s)")
    else:
        wl = f"synthetic {args.frames}-frame KITTI-shaped stereo sequence (BASELINE configs[1])"
    return {"workload": wl + f": 1241x376, ~{args.features} Harris features/image, 121-element Sobel descriptors, stereo + "
                             f"2x temporal SAD matching, circle closure, triangulation, RANSAC({args.hyp}) + Gauss-Newton",
            "sequences": args.sequences, "sequences_per_rank": per, "frames": args.frames, "features": args.features,
            "ransac_iter": args.hyp, "unique_rendered_frames": min(args.unique, args.frames),
            "l2": "inputs larger than L2 (every frame has its own HBM copy: ~3 GB per sequence vs 126 MB L2)",
            "input": {"images": "two 8-bit images + keypoints per frame, descriptors extracted on the device (viso.cpp:1004-1024)",
                      "raw": "two 8-bit images per frame; Harris detector (viso.cpp:925-976) and descriptors on the device",
                      "descriptors": "keypoints + n x 121 f32 descriptor matrices per frame (cv::Mat layout)"}[args.input],
            "e2e_pipeline": (f"chunks of {args.chunk} frames: H2D on a copy stream overlapped with the previous chunk's kernels; "
                             f"{args.e2e_buffers} sequence object(s) on separate contexts take alternate (step, sequence) items"),
            "parallelism": (f"sequence s on rank s mod {world}; one NCCL gather of the 64-byte records per step, inside the "
                            f"timed regions" if world > 1 else "1 GPU, no collective")}


class StdoutGuard:
    """stdout must carry exactly one JSON line: point fd 1 at stderr while libraries (NCCL's version banner, ...) may
    print, and give it back for the result"""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(text, flush=True)
        os.dup2(2, 1)


def gpu_uuid(dev):
    """'GPU-...' of CUDA device `dev` (what nvidia-smi -i and NVML take), or None with an older torch"""
    try:
        import torch
        u = getattr(torch.cuda.get_device_properties(dev), "uuid", None)
        return None if u is None else ("GPU-" + str(u) if not str(u).startswith("GPU-") else str(u))
    except Exception:
        return None


def nvml_handle(dev):
    import pynvml
    pynvml.nvmlInit()
    u = gpu_uuid(dev)
    if u:
        try:
            return pynvml.nvmlDeviceGetHandleByUUID(u.encode() if hasattr(u, "encode") else u)
        except Exception:
            pass
    idx = dev
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v for v in vis.split(",") if v.strip()]
        if dev < len(ids) and ids[dev].strip().isdigit():
            idx = int(ids[dev])
    return pynvml.nvmlDeviceGetHandleByIndex(idx)


def gpu_cpus(dev):
    """the CPUs NVML calls local to CUDA device `dev` (its NUMA node)"""
    import pynvml
    words = ((os.cpu_count() or 64) + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(nvml_handle(dev), words)
    return frozenset(64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1)


def pick_device(local_rank, world, n_visible):
    """CUDA device of this rank.  With as many visible GPUs as ranks: device = local rank.  With more (a 4-rank run on
    an 8-GPU box), the ranks are dealt round-robin over the GPUs' NUMA nodes instead of packed onto the first node: the
    end-to-end path streams ~1 GB per sequence per step from pinned host memory, and four copy streams behind one
    socket's memory and PCIe root measured 27.7 GB/s each against 51.7 GB/s for two (profiles/r02_g_bench_n4/n2).
    Every rank computes the same order, so the devices are distinct.  The device-resident numbers do not depend on it."""
    if n_visible <= world or os.environ.get("VISO_BENCH_PACK_GPUS"):
        return local_rank, "device = local rank"
    try:
        groups = {}
        for d in range(n_visible):
            groups.setdefault(gpu_cpus(d), []).append(d)
        if len(groups) < 2:
            # a VM shows one node whatever the board looks like (the 8-GPU boxes of this pool do: GPUs 0-3 then measure
            # 21 GB/s each and GPUs 4-7 31 GB/s with all eight copying): spread the ranks evenly over the device indices,
            # adjacent indices being the ones that share a PCIe switch / root on HGX boards
            stride = n_visible // world
            return local_rank * stride, "ranks spread over the %d visible GPUs with stride %d" % (n_visible, stride)
        lists = sorted(groups.values(), key=lambda g: g[0])
        order = [g[i] for i in range(max(len(g) for g in lists)) for g in lists if i < len(g)]
        return order[local_rank], "ranks dealt round-robin over %d NUMA nodes: %s" % (len(lists), order[:world])
    except Exception as e:  # no NVML: keep the plain mapping
        return local_rank, "device = local rank (%s)" % e


def bind_to_gpu_cpus(dev):
    """The pinned host buffers are first-touched by this process: keep it on the CPUs (NUMA node) next to its GPU.
    A no-op where NVML reports every CPU (single-socket boxes)."""
    try:
        cpus = set(gpu_cpus(dev)) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:  # no NVML, restricted cpuset, ...: keep the inherited affinity
        print(f"[bench] CPU binding skipped: {e}", file=sys.stderr)


class DeviceWords:
    """a raw device buffer of int32 words exposed through __cuda_array_interface__ (zero-copy torch view)"""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 2}


def tr_close(a, b):
    """north_star's pose tolerance: 1e-6 rad, 1e-6 relative translation"""
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return True
    scale = np.maximum(np.linalg.norm(b[..., 3:], axis=-1, keepdims=True), 1e-3)
    return bool(np.all(np.abs(a[..., :3] - b[..., :3]) <= 1e-6) and np.all(np.abs(a[..., 3:] - b[..., 3:]) <= 1e-6 * scale))


# ------------------------------------------------------------------------------------------------ configs[2], configs[3]

def extra_configs(ctx, api, synth, check):
    """BASELINE configs[2] (20 000 keypoints per image, SAD matching only) and configs[3] (4096 hypotheses x 10 000
    correspondences) through the same library; device times from CUDA events"""
    out = {}
    peak, _ = hbm_peak()
    n = 20000
    pair = synth.make_dense_pair(n, seed=2000)
    P1, P2 = synth.kitti_calib()
    prm = api.param_default(ransac_iter=1)

    # `copies` identical frames: one launch holds `copies` stereo jobs or 2 (copies - 1) temporal jobs, so that a
    # microsecond-scale roofline time is not hidden behind the launch latency
    copies = 9
    seq = ctx.sequence(copies, n, 121, 1)
    seq.set_calib(P1, P2)
    for t in range(copies):
        seq.upload_frame(t, pair["kpL"], pair["kpR"], pair["dL"], pair["dR"])
    seq.run(prm, np.zeros((copies, 1, 3), np.uint32))
    ctx.sync()
    dense0 = seq.get_dense(0, 0)
    b_call = (n + n) * (128 * 2 + 8) + 16 * n   # SURVEY 8d, u16 layout: both sets read once + the dense int4 output
    out["config3"] = {
        "workload": "BASELINE configs[2]: 1241x376 stereo pair, 20 000 keypoints per image, match_desc only; "
                    f"{copies} copies of the pair per launch",
        "algorithmic_bytes_per_call": b_call}
    for which, name, jobs in ((0, "stereo", copies), (1, "temporal", 2 * (copies - 1))):
        seq.time_match(which)
        runs = [seq.time_match(which) for _ in range(3)]
        ms = float(np.median([r[0] for r in runs])) / jobs
        pairs = runs[0][1] / jobs
        out["config3"][name] = {"ms": ms, "GB/s": b_call / (ms * 1e-3) / 1e9, "frac": b_call / (ms * 1e-3) / 1e9 / peak,
                                "sad_pairs_per_call": pairs, "pairs/s": pairs / (ms * 1e-3),
                                "queries_left_to_generic_kernel_per_call": runs[0][2] / jobs, "jobs_per_launch": jobs}
    seq.close()
    # configs[3]
    npts, H = 10000, 4096
    X, obs, tr_true = synth.make_ransac_problem(npts, seed=3000)
    table = api.randomsample_table(424242, H, npts)
    p = api.param_default(base=synth.BASE, f=synth.F_PX, cu=synth.CU, cv=synth.CV, ransac_iter=H)
    r = ctx.ransac_minimize_reproj(X, obs, p, table)
    dev = []
    for _ in range(3):
        ctx.sync()
        ctx.timer_begin()
        r = ctx.ransac_minimize_reproj(X, obs, p, table)
        dev.append(ctx.timer_end())
    ms = float(np.median(dev))
    out["config4"] = {
        "workload": "BASELINE configs[3]: 4096 hypotheses x 10 000 stereo correspondences (30 % outliers), one "
                    "ransac_minimize_reproj call through the C-ABI (host buffers in, inlier set out)",
        "ms": ms, "scoring_GFLOPs_f64": 38.0 * H * npts / (ms * 1e-3) / 1e9, "ok": bool(r["ok"]),
        "n_inliers": int(len(r["inliers"])), "best_hyp": int(r["best_hyp"])}
    if check:
        from oracle import oracle
        F_ = oracle.F_from_P(P1, P2)
        o = oracle.match_desc(pair["kpL"], pair["kpR"], pair["dL"], pair["dR"], oracle.match_params_stereo(F_))
        out["config3"]["stereo"]["equal_oracle"] = bool(all(np.array_equal(dense0[:, i], o[k]) for i, k in enumerate(("idx", "d1", "d2", "valid"))))
        po = oracle.param_default(base=synth.BASE, f=synth.F_PX, cu=synth.CU, cv=synth.CV, ransac_iter=H)
        t0 = time.perf_counter()
        o = oracle.ransac_minimize_reproj(X, obs, po, table)
        c4 = out["config4"]
        c4["cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t0)
        c4["hyp_ok_equal_oracle"] = bool(np.array_equal(o["hyp_ok"], r["hyp_ok"]))
        c4["hyp_count_equal_oracle"] = bool(np.array_equal(o["hyp_count"], r["hyp_count"]))
        c4["best_hyp_equal_oracle"] = bool(o["best_hyp"] == r["best_hyp"])
        c4["inliers_equal_oracle"] = bool(np.array_equal(o["inliers"], r["inliers"]))
        c4["tr_within_1e-6"] = tr_close(r["tr"], o["tr"])
    return out


# ------------------------------------------------------------------------------------------------ main

def main():
    args = parse_args()
    guard = StdoutGuard()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, guard)
        return

    import torch
    import torch.distributed as dist
    from libviso_b200 import api, build, synth
    from libviso_b200.distributed import gather_rows, rows_per_rank, shard_sequences, split_gathered
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libviso_b200 has no CPU path")
    dev, placement = pick_device(local_rank, world, torch.cuda.device_count())
    torch.cuda.set_device(dev)
    bind_to_gpu_cpus(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's own log lines: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cores = os.cpu_count() or 1
    F, H, S = args.frames, args.hyp, args.sequences
    my = shard_sequences(S, rank, world)                      # sequence s lives on rank s mod G
    per_rank_max = rows_per_rank(S, world)
    P1, P2 = synth.kitti_calib()
    use_img = args.input in ("images", "raw")
    use_raw = args.input == "raw"   # images only: keypoints detected on the device as well (viso.cpp:925-976)

    ctx = api.Context(dev)
    ctx.set_image_extent(synth.W, synth.H)
    param = api.param_default(ransac_iter=H)
    words = F * 16                                            # one record = 64 bytes = 16 int32 words

    # ---- the rank's sequences: rendered frames, host copies, one resident sequence object each ----
    data = []
    for s in my:
        frames, order = build_sequence(args, args.seed + s, max(1, min(16, cores // world)))
        seeds = make_seeds(F, H, args.seed + s)
        cap = max(max(len(f["kpL"]), len(f["kpR"])) for f in frames)
        data.append({"s": s, "frames": frames, "order": order, "seeds": seeds, "cap": cap})
    cap = max(d["cap"] for d in data)
    for d in data:
        seq = ctx.sequence(F, cap, 121, H)
        seq.set_calib(P1, P2)
        if use_img:
            seq.set_image_size(synth.W, synth.H)
        if use_raw:
            seq.set_detector(args.features)
        d["seq"] = seq
        pinned = []
        for f in d["frames"]:
            p = {}
            for k in ("kpL", "kpR") + (("imL", "imR") if use_img else ("dL", "dR")):
                dt = np.uint8 if k.startswith("im") else np.float32
                p[k] = torch.from_numpy(np.ascontiguousarray(f[k], dtype=dt)).pin_memory()
            p["nL"], p["nR"] = len(f["kpL"]), len(f["kpR"])
            pinned.append(p)
        d["pinned"] = pinned
        d["seeds_pin"] = torch.from_numpy(d["seeds"].view(np.int32)).pin_memory()
        for t in range(F):
            p = pinned[d["order"][t]]
            if use_raw:
                seq.upload_frame_raw_ptr(t, p["imL"].data_ptr(), p["imR"].data_ptr())
            elif use_img:
                seq.upload_frame_images_raw(t, p["imL"].data_ptr(), p["imR"].data_ptr(), p["kpL"].data_ptr(), p["nL"],
                                            p["kpR"].data_ptr(), p["nR"])
            else:
                seq.upload_frame_raw(t, p["kpL"].data_ptr(), p["nL"], p["kpR"].data_ptr(), p["nR"], p["dL"].data_ptr(), p["dR"].data_ptr())
        seq.set_seeds(d["seeds"], H)
        ctx.sync()
        d["rec_view"] = torch.as_tensor(DeviceWords(seq.records_device_ptr(), words), device="cuda")

    capq = data[0]["seq"].capacity()
    if use_raw:
        h2d_seq = F * 2 * synth.W * synth.H + H * F * 12
    elif use_img:  # keypoint rows are uploaded padded to the sequence capacity
        h2d_seq = F * (2 * synth.W * synth.H + 2 * capq * 8) + H * F * 12
    else:
        h2d_seq = None
    n_pairs_rank = len(my) * (F - 1)
    n_pairs_total = S * (F - 1)

    # the only collective: the records of every rank's sequences -> rank 0, from device memory, on the library's stream
    lib_stream = torch.cuda.ExternalStream(ctx.stream_ptr())
    stage = torch.zeros((2, per_rank_max, words), dtype=torch.int32, device="cuda") if world > 1 else None
    gathered = [torch.zeros((per_rank_max, words), dtype=torch.int32, device="cuda") for _ in range(world)] if (world > 1 and rank == 0) else None

    def gather_on(stream, buf):
        """dist.gather of buf (per_rank_max x words) on `stream`; blocks the stream, not the host"""
        if world == 1:
            return
        with torch.cuda.stream(stream):
            gather_rows(buf, rank, world, gathered)

    def resident_step(k):
        for d in data:
            d["seq"].run(param)
        if world > 1:
            with torch.cuda.stream(lib_stream):
                for j, d in enumerate(data):
                    stage[k & 1, j].copy_(d["rec_view"], non_blocking=True)
            gather_on(lib_stream, stage[k & 1])

    # ---- device-resident throughput ----
    for k in range(max(args.warmup, 3)):
        resident_step(k)
    ctx.sync()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    sampler = ClockSampler(gpu_uuid(dev) or dev)
    barrier()
    sampler.start()
    match_ms = []
    ctx.timer_begin()
    for k in range(args.steps):
        resident_step(k)
        match_ms.append(sum(d["seq"].match_ms() for d in data))
    dev_ms = ctx.timer_end()
    barrier()
    launches = ctx.launch_count() - l0
    recs = {d["s"]: d["seq"].download() for d in data}
    match_bytes = sad_pairs = sad_eval = n_pending = 0
    for d in data:
        mb, sp, se = d["seq"].stats()
        match_bytes += mb; sad_pairs += sp; sad_eval += se
        n_pending += d["seq"].last_pending()
    seq_bytes = data[0]["seq"].device_bytes()
    if world > 1 and rank == 0:   # what the gather delivered must be what the ranks computed
        g0 = gathered[0].cpu().numpy()
        for j, d in enumerate(data):
            got = g0[j].view(api.RECORD_DTYPE).reshape(-1).copy()
            got[0] = recs[d["s"]][0]   # frame 0 holds no pose; download() blanks it
            assert got.tobytes() == recs[d["s"]].tobytes(), "gathered records differ from the downloaded ones"

    # ---- end to end through the C-ABI with host buffers ----
    e2e_ms = None
    h2d = d2h = 0
    if not args.no_e2e:
        for d in data:
            if use_img:
                # the sequence as a camera / front end would leave it in pinned host memory: per frame the two images
                # back to back, keypoints padded to the sequence capacity -- so a chunk of frames is three large copies
                d["img_host"] = torch.empty((F, 2, synth.H, synth.W), dtype=torch.uint8).pin_memory()
                d["kpL_host"] = torch.zeros((F, capq, 2), dtype=torch.float32).pin_memory()
                d["kpR_host"] = torch.zeros((F, capq, 2), dtype=torch.float32).pin_memory()
                d["nL_host"] = torch.zeros(F, dtype=torch.int32); d["nR_host"] = torch.zeros(F, dtype=torch.int32)
                for t in range(F):
                    p = d["pinned"][d["order"][t]]
                    d["img_host"][t, 0] = p["imL"]; d["img_host"][t, 1] = p["imR"]
                    d["kpL_host"][t, :p["nL"]] = p["kpL"]; d["kpR_host"][t, :p["nR"]] = p["kpR"]
                    d["nL_host"][t] = p["nL"]; d["nR_host"][t] = p["nR"]
        img_b, kp_b = 2 * synth.H * synth.W, capq * 8
        if h2d_seq is None:
            h2d = sum(sum((d["pinned"][i]["nL"] + d["pinned"][i]["nR"]) * (8 + 121 * 4) for i in d["order"]) + H * F * 12 for d in data)
        else:
            h2d = len(my) * h2d_seq
        d2h = len(my) * F * 64

        # Two sequence objects on two contexts (each with its own compute and copy streams) take alternate work items
        # (step, sequence), so the tail of one item (the kernels of its last chunk, the record read-back) overlaps the
        # uploads of the next and the PCIe link never idles.  Every item uploads all of its inputs and reads its own
        # records back inside the timed region; with N > 1 the step's records are also gathered from device memory.
        lanes = [(ctx, data[0]["seq"])]
        if args.e2e_buffers > 1:
            ctx2 = api.Context(dev)
            ctx2.set_image_extent(synth.W, synth.H)
            if not args.e2e_separate_copy_streams:
                ctx2.share_copy_stream(ctx)   # uploads of the two lanes are served FIFO, not interleaved
            seq2 = ctx2.sequence(F, cap, 121, H)
            seq2.set_calib(P1, P2)
            if use_img:
                seq2.set_image_size(synth.W, synth.H)
            if use_raw:
                seq2.set_detector(args.features)
            lanes.append((ctx2, seq2))
        L = len(lanes)
        lane_stream = [torch.cuda.ExternalStream(c.stream_ptr()) for c, _ in lanes]
        lane_view = [torch.as_tensor(DeviceWords(sq.records_device_ptr(), words), device="cuda") for _, sq in lanes]
        rec_pins = [torch.zeros(words, dtype=torch.int32).pin_memory() for _ in lanes]
        lane_done = [torch.cuda.Event() for _ in lanes]
        gather_done = [None, None]   # per staging buffer: the event after the gather that last read it
        e2e_recs = {}

        def enqueue(item):
            k, j = divmod(item, len(data))
            d = data[j]
            li = item % L
            c, sq = lanes[li]
            c._ck(api.lib().viso_seq_set_seeds(sq.h, api._p(d["seeds_pin"].data_ptr()), H))
            for t0 in range(0, F, args.chunk):   # chunked pipeline: uploads of chunk k+1 overlap the kernels of chunk k
                t1 = min(F, t0 + args.chunk)
                if use_raw:
                    sq.upload_chunk_raw(t0, t1 - t0, d["img_host"].data_ptr() + t0 * img_b)
                elif use_img:
                    sq.upload_chunk_images_raw(t0, t1 - t0, d["img_host"].data_ptr() + t0 * img_b, d["kpL_host"].data_ptr() + t0 * kp_b,
                                               d["nL_host"].data_ptr() + 4 * t0, d["kpR_host"].data_ptr() + t0 * kp_b,
                                               d["nR_host"].data_ptr() + 4 * t0)
                else:
                    for t in range(t0, t1):
                        p = d["pinned"][d["order"][t]]
                        sq.upload_frame_raw(t, p["kpL"].data_ptr(), p["nL"], p["kpR"].data_ptr(), p["nR"], p["dL"].data_ptr(), p["dR"].data_ptr())
                sq.run_range(param, t0, t1)
            if world > 1:
                with torch.cuda.stream(lane_stream[li]):
                    if gather_done[k & 1] is not None:
                        lane_stream[li].wait_event(gather_done[k & 1])   # the staging buffer is free again
                    stage[k & 1, j].copy_(lane_view[li], non_blocking=True)
                    lane_done[li].record()
                if j == len(data) - 1:   # the step is complete on this rank: gather it (on lane 0's stream)
                    for o in range(1, L):
                        lane_stream[0].wait_event(lane_done[o])
                    gather_on(lane_stream[0], stage[k & 1])
                    with torch.cuda.stream(lane_stream[0]):
                        gather_done[k & 1] = torch.cuda.Event()
                        gather_done[k & 1].record()

        def collect(item):
            li = item % L
            lanes[li][1].download_raw(rec_pins[li].data_ptr())   # synchronises that lane only
            e2e_recs[data[item % len(data)]["s"]] = np.frombuffer(rec_pins[li].numpy().tobytes(), dtype=api.RECORD_DTYPE).copy()

        def e2e_run(n_items):
            for i in range(n_items):
                enqueue(i)
                if i >= L - 1:
                    collect(i - (L - 1))
            for i in range(max(0, n_items - (L - 1)), n_items):
                collect(i)

        e2e_run(2 * len(data))   # warm-up: two whole steps on every rank (the same number of gathers everywhere)
        for c, _ in lanes:
            c.sync()
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps * len(data))
        for c, _ in lanes:
            c.sync()
        torch.cuda.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0)   # host clock around fully synchronised work on every lane
        barrier()
        for s_id, r in e2e_recs.items():
            assert r.tobytes() == recs[s_id].tobytes(), "e2e records differ from the resident run"
        for c2, s2 in lanes[1:]:
            s2.close()
            c2.close()

    clocks = sampler.stop()  # sampled over both timed regions (device-resident and end-to-end)

    # ---- max over ranks; every rank's own numbers ----
    mm = float(np.mean(match_ms))
    on = "cuda" if world > 1 else "cpu"   # one GPU: nothing of torch's runs on the device
    tms = torch.tensor([dev_ms, e2e_ms or 0.0], dtype=torch.float64, device=on)
    mine = torch.tensor([dev_ms / args.steps, (e2e_ms or 0.0) / args.steps, mm, float(sad_pairs), float(clocks.get("sm_mhz") or 0.0),
                         float(n_pairs_rank), (h2d / ((e2e_ms or 1.0) / args.steps * 1e-3) / 1e9) if e2e_ms else 0.0,
                         float(match_bytes), float(n_pending), float(sad_eval)], dtype=torch.float64, device=on)
    if world > 1:
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(per_rank, mine)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        # rank 0 chains poses from the gathered records of the last resident step (viso.cpp:1313-1321)
        all_rec = None
        if rank == 0:
            all_rec = split_gathered(gathered, S, world)
    else:
        per_rank = [mine]
        all_rec = recs
    dev_ms_g, e2e_ms_g = [float(v) for v in tms.cpu()]

    if rank == 0:
        pr = np.stack([t.cpu().numpy() for t in per_rank])
        value = n_pairs_total * args.steps / (dev_ms_g * 1e-3)
        peak, peak_src = hbm_peak()
        # the dominant kernel on the slowest rank of the resident run
        slow = int(np.argmax(pr[:, 0]))
        mb, mmr = float(pr[slow, 7]), float(pr[slow, 2])
        achieved = mb / (mmr * 1e-3) / 1e9
        n_poses = {int(s_id): len(api.chain_poses(r)) for s_id, r in sorted(all_rec.items())}
        rl = list(all_rec.values())
        line = {
            "metric": METRIC, "value": value, "unit": "frame-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms_g / args.steps,
            "higher_is_better": True, "scaling": "strong" if (args.mode == "configs[4]" and world > 1) else "weak",
            "vs_baseline": None, "dtype": "u16 SAD / f64 pose",
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": int(launches), "placement": placement,
            "per_rank": {k: [round(float(v), 3) for v in pr[:, i]]
                         for i, k in enumerate(("ms_per_step", "e2e_ms_per_step", "sad_ms_per_step", "sad_pairs", "sm_mhz",
                                                "frame_pairs_per_step", "e2e_h2d_GBps"))},
            "roofline": {"bound": "hbm", "kernel": "sad_match_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args), "peak_source": peak_src,
                         "note": "HBM fraction as the contract defines it; the kernel is bound by the L1 data pipe and instruction "
                                 "issue, not by HBM (DESIGN.md section 4); figures of the slowest rank, all sad_match launches of a step",
                         "algorithmic_bytes_per_step": int(mb), "kernel_ms": mmr,
                         "sad_pairs_per_step": int(pr[slow, 3]), "queries_left_to_generic_kernel": int(pr[slow, 8]),
                         "sad_evaluated_per_step": int(pr[slow, 9]),
                         "kernel_share_of_step": mmr / float(pr[slow, 0]),
                         # the roof of the kernel's own access pattern: the same row gather without arithmetic,
                         # measured once on B200 by tools/ubench_rowgather.cu (profiles/r01_e_rowgather_roof.txt)
                         "gather_roof": {"rows_per_s": 80.1e9, "achieved_rows_per_s": float(pr[slow, 3]) / (mmr * 1e-3),
                                         "frac": float(pr[slow, 3]) / (mmr * 1e-3) / 80.1e9,
                                         "source": "tools/ubench_rowgather.cu, measured on B200 in round 1 (not re-measured by this run)"}},
            "poses": {"chained_per_sequence": n_poses, "ok_frame_pairs": int(sum(int(r["ok"].sum()) for r in rl)),
                      "circular_matches_mean": float(np.mean([r["n_circ"][1:].mean() for r in rl])),
                      "inliers_mean": float(np.mean([r["n_inliers"][1:].mean() for r in rl])),
                      # RANSAC scoring work of one sequence: hypotheses x circular matches, 38 flops each
                      "ransac_point_tests_per_sequence": int(args.hyp * int(rl[0]["n_circ"][1:].sum()))},
            "memory": {"sequence_object_bytes": int(seq_bytes)},
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": n_pairs_total * args.steps / (e2e_ms_g * 1e-3), "unit": "frame-pairs/s",
                           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                           "ms_per_step": e2e_ms_g / args.steps,
                           "bytes_are": "rank 0's; every rank moves the same per sequence it owns"}
        if world == 1 and not args.no_cpu:
            kind, mod = cpu_backend()
            d = data[0]
            # windows of 8 frame pairs spread over the whole sequence (not only its start)
            npairs = min(args.cpu_pairs, F - 1)
            win = 8
            nwin = max(1, npairs // win)
            starts = [int(round(i * (F - 1 - win) / max(nwin - 1, 1))) for i in range(nwin)] if F - 1 > win else [0]
            secs = 0.0
            same_int, same_tr, checked = True, True, 0
            for t0 in starts:
                w = min(win, F - 1 - t0)
                dt, rec_o = cpu_pairs(kind, mod, d["frames"], d["order"], t0, w, H, d["seeds"])
                secs += dt
                g = recs[d["s"]][t0 + 1:t0 + w + 1]
                o = rec_o[1:]
                keys = ("ok", "n_inliers", "n_circ") + (("best_hyp",) if kind == "port" else ())
                same_int &= all(np.array_equal(g[k], o[k]) for k in keys)
                same_tr &= tr_close(g["tr"], o["tr"])
                checked += w
            line["cpu_baseline"] = {"value": checked / secs, "unit": "frame-pairs/s", "cores": 1, "kind": kind,
                                    "sample": f"{len(starts)} windows of {win} frame pairs spread over the {F}-frame sequence "
                                              f"({checked} pairs), " + ("the reference's own functions (oracle/_ref, g++ -O2)" if kind == "reference"
                                                                       else "CPU oracle port (g++ -O2)") + f", {secs:.1f} s",
                                    "records_match_gpu": bool(same_int), "tr_within_1e-6": bool(same_tr)}
        if world == 1 and not args.no_extra:
            try:
                line["extra"] = extra_configs(ctx, api, synth, check=not args.no_cpu)
            except Exception as e:  # noqa: BLE001 -- the headline line must survive a failure of the side measurements
                line["extra"] = {"error": repr(e)}
        guard.emit(json.dumps(line))
    for d in data:
        d["seq"].close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
