#!/usr/bin/env python
"""bench.py -- frame-pairs/s of the libviso hot path (match + circle + triangulate + RANSAC/GN pose) on B200.

A "step" is one pass of the whole per-frame loop of sequence_odometry (reference src/viso.cpp:1205-1327, minus
detection / description / debug output) over one synthetic KITTI-shaped 1241x376 stereo sequence of --frames frames
(BASELINE.json configs[1]: 1000 frames, ~2k features per image, 50 RANSAC hypotheses per frame pair).

  value  device throughput: features already resident in HBM (the reference's f32 cv::Mat descriptor layout + keypoints)
         when the timed region starts; K steps timed with CUDA events on the library's stream
  e2e    the same through the C-ABI with HOST buffers: every step copies all frames' keypoints + descriptors + sample
         seeds from pinned host memory and reads the 64-byte records back
  roofline   the sad_match kernel: algorithmic bytes (per frame pair, u16 layout) / its CUDA-event time
  cpu_baseline   the CPU oracle (restated reference path, 1 thread) on a bounded sample of the same sequence

Multi-GPU: one process per GPU (torchrun), each rank owns one independent sequence (seed 1000+rank) -- frame pairs
and sequences are independent, so there is no data-path collective; the per-frame-pair records are gathered to
rank 0 over NCCL and chained into poses there.  Weak scaling.

`--impl reference` times the reference's CPU path (the oracle port; the reference itself cannot be built in this
image) on the host cores with one process per core over independent frame ranges.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--features", type=int, default=2040)
    ap.add_argument("--hyp", type=int, default=50)
    ap.add_argument("--unique", type=int, default=64,
                    help="distinct rendered frames; the sequence drives back and forth over them (every frame has "
                         "its own HBM copy, so the working set is the full --frames)")
    ap.add_argument("--seed", type=int, default=1000, help="rank r renders the sequence with seed --seed + r")
    ap.add_argument("--cpu-pairs", type=int, default=96, help="frame pairs in the cpu_baseline sample")
    ap.add_argument("--chunk", type=int, default=0,
                    help="frames per upload/compute chunk of the e2e pipeline (0: 125, or 250 with --input raw -- measured)")
    ap.add_argument("--e2e-buffers", type=int, default=2, choices=[1, 2],
                    help="sequence objects (each on its own context / streams) taking alternate e2e steps")
    ap.add_argument("--input", default="images", choices=["images", "descriptors", "raw"],
                    help="what crosses the boundary per frame: the two 8-bit images + keypoints (descriptors extracted "
                         "on the device, viso.cpp:1004-1024) or the reference's n x 121 f32 descriptor matrices")
    ap.add_argument("--e2e-separate-copy-streams", action="store_true",
                    help="A/B switch: every e2e lane uploads on its own copy stream (pieces of the two lanes interleave)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
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
    order = pingpong(args.frames, n_unique)
    return frames, order


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


def cpu_oracle_pairs(frames, order, t0, n_pairs, H, seeds, front_end=0):
    """run the CPU oracle over frames [t0, t0+n_pairs] of the long sequence; returns (seconds, records)"""
    from oracle import oracle
    from libviso_b200 import synth
    P1, P2 = synth.kitti_calib()
    sub = [frames[order[t]] for t in range(t0, t0 + n_pairs + 1)]
    param = oracle.param_default(ransac_iter=H)
    sd = np.ascontiguousarray(seeds[t0:t0 + n_pairs + 1])
    t = time.perf_counter()
    if front_end:
        # --input raw: the reference's front end per frame too (viso.cpp:1226-1231), with OpenCV doing what the
        # reference asks OpenCV to do (cv::cornerHarris, cv::Sobel) -- the oracle's own scalar detector would be an
        # unfairly slow stand-in
        sub = [synth.make_features(f["imL"], f["imR"], front_end) for f in sub]
    out = oracle.sequence(sub, P1, P2, param, sd)
    return time.perf_counter() - t, out["records"]


def make_seeds(n_frames, H, seed):
    return np.random.default_rng(424242 + seed).integers(0, 2 ** 32, size=(n_frames, H, 3), dtype=np.uint32)


def run_reference(args, rank, world, guard):
    """--impl reference: the reference's CPU path (oracle port) on the host cores"""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    frames, order = build_sequence(args, 1000, min(cores, 16))
    seeds = make_seeds(args.frames, args.hyp, 1000)
    pairs_per_proc = 6
    jobs_per_step = cores
    global _REF_STATE
    _REF_STATE = (frames, order, seeds, args.hyp, args.features if args.input == "raw" else 0)
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
    sample = (f"{jobs_per_step} processes x {pairs_per_proc} frame pairs per step, independent frame ranges of the same "
              f"{args.frames}-frame sequence")
    if args.input == "raw":
        sample += "; front end per frame = cv2.cornerHarris + cv2.Sobel + numpy binning / patch gather (libviso_b200/synth.py)"
    line = {
        "impl": "reference", "metric": "frame-pairs/s (match+RANSAC pose) at 1241x376", "value": value,
        "unit": "frame-pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16 SAD / f64 pose", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    guard.emit(json.dumps(line))


_REF_STATE = None


def _ref_step(job):
    t0, n_pairs = job
    frames, order, seeds, H, front_end = _REF_STATE
    return cpu_oracle_pairs(frames, order, t0, n_pairs, H, seeds, front_end)[0]


def workload_config(args):
    return {"workload": f"synthetic {args.frames}-frame KITTI-shaped stereo sequence per GPU (BASELINE configs[1]): "
                        f"1241x376, ~{args.features} Harris features/image, 121-element Sobel descriptors, stereo + "
                        f"2x temporal SAD matching, circle closure, triangulation, RANSAC({args.hyp}) + Gauss-Newton",
            "frames": args.frames, "features": args.features, "ransac_iter": args.hyp,
            "unique_rendered_frames": min(args.unique, args.frames),
            "l2": "inputs larger than L2 (every frame has its own HBM copy: ~3 GB per sequence vs 126 MB L2)",
            "input": {"images": "two 8-bit images + keypoints per frame, descriptors extracted on the device (viso.cpp:1004-1024)",
                      "raw": "two 8-bit images per frame; Harris detector (viso.cpp:925-976) and descriptors on the device",
                      "descriptors": "keypoints + n x 121 f32 descriptor matrices per frame (cv::Mat layout)"}[args.input],
            "e2e_pipeline": (f"chunks of {args.chunk} frames: H2D on a copy stream overlapped with the previous chunk's kernels; "
                             f"{args.e2e_buffers} sequence object(s) on separate contexts take alternate steps"),
            "parallelism": f"{args.gpus} independent sequence(s), one per GPU, NCCL gather of 64-byte records"}


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


def bind_to_gpu_cpus(local_rank):
    """The pinned host buffers are first-touched by this process: keep it on the CPUs (NUMA node) next to its GPU.
    A no-op where NVML reports every CPU (single-socket boxes)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [v for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].strip().isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:  # no NVML, restricted cpuset, ...: keep the inherited affinity
        print(f"[bench] CPU binding skipped: {e}", file=sys.stderr)


def main():
    args = parse_args()
    guard = StdoutGuard()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, guard)
        return

    import torch
    import torch.distributed as dist
    from libviso_b200 import api, build, synth
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libviso_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_cpus(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's own log lines: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cores = os.cpu_count() or 1
    frames, order = build_sequence(args, args.seed + rank, max(1, min(16, cores // world)))
    F, H = args.frames, args.hyp
    seeds = make_seeds(F, H, args.seed + rank)
    P1, P2 = synth.kitti_calib()
    cap = max(max(len(f["kpL"]), len(f["kpR"])) for f in frames)

    ctx = api.Context(local_rank)
    ctx.set_image_extent(synth.W, synth.H)
    seq = ctx.sequence(F, cap, 121, H)
    seq.set_calib(P1, P2)
    param = api.param_default(ransac_iter=H)

    # pinned host copies of the unique frames (the e2e path uploads from these every step)
    use_img = args.input in ("images", "raw")
    use_raw = args.input == "raw"   # images only: keypoints detected on the device as well (viso.cpp:925-976)
    if use_img:
        seq.set_image_size(synth.W, synth.H)
    if use_raw:
        seq.set_detector(args.features)
    pinned = []
    for f in frames:
        p = {}
        for k in ("kpL", "kpR") + (("imL", "imR") if use_img else ("dL", "dR")):
            dt = np.uint8 if k.startswith("im") else np.float32
            p[k] = torch.from_numpy(np.ascontiguousarray(f[k], dtype=dt)).pin_memory()
        p["nL"], p["nR"] = len(f["kpL"]), len(f["kpR"])
        pinned.append(p)
    seeds_pin = torch.from_numpy(seeds.view(np.int32)).pin_memory()
    rec_pin = torch.zeros(F * 16, dtype=torch.int32).pin_memory()

    def upload_range(t0, t1):
        for t in range(t0, t1):
            p = pinned[order[t]]
            if use_raw:
                seq.upload_frame_raw_ptr(t, p["imL"].data_ptr(), p["imR"].data_ptr())
            elif use_img:
                seq.upload_frame_images_raw(t, p["imL"].data_ptr(), p["imR"].data_ptr(), p["kpL"].data_ptr(), p["nL"],
                                            p["kpR"].data_ptr(), p["nR"])
            else:
                seq.upload_frame_raw(t, p["kpL"].data_ptr(), p["nL"], p["kpR"].data_ptr(), p["nR"],
                                     p["dL"].data_ptr(), p["dR"].data_ptr())

    def upload_all():
        upload_range(0, F)

    if use_raw:
        h2d = F * 2 * synth.W * synth.H + seeds.nbytes
    elif use_img:  # keypoint rows are uploaded padded to the sequence capacity
        h2d = F * (2 * synth.W * synth.H + 2 * seq.capacity() * 8) + seeds.nbytes
    else:
        h2d = sum((pinned[i]["nL"] + pinned[i]["nR"]) * (8 + 121 * 4) for i in order) + seeds.nbytes
    d2h = F * 64
    n_pairs = F - 1

    # ---- device-resident throughput ----
    upload_all()
    seq.set_seeds(seeds, H)
    ctx.sync()
    for _ in range(max(args.warmup, 3)):
        seq.run(param)
    ctx.sync()
    l0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    match_ms = []
    ctx.timer_begin()
    for _ in range(args.steps):
        seq.run(param)
        match_ms.append(seq.match_ms())
    dev_ms = ctx.timer_end()
    barrier()
    launches = ctx.launch_count() - l0
    rec = seq.download()
    match_bytes, sad_pairs, sad_eval = seq.stats()
    n_pending = seq.last_pending()

    # ---- end to end through the C-ABI with host buffers ----
    e2e_ms = None
    if not args.no_e2e:
        if use_img:
            # the sequence as a camera / front-end would leave it in pinned host memory: per frame the two images back
            # to back, keypoints padded to the sequence capacity -- so a chunk of frames is three large copies
            capq = seq.capacity()
            img_host = torch.empty((F, 2, synth.H, synth.W), dtype=torch.uint8).pin_memory()
            kpL_host = torch.zeros((F, capq, 2), dtype=torch.float32).pin_memory()
            kpR_host = torch.zeros((F, capq, 2), dtype=torch.float32).pin_memory()
            nL_host = torch.zeros(F, dtype=torch.int32); nR_host = torch.zeros(F, dtype=torch.int32)
            for t in range(F):
                p = pinned[order[t]]
                img_host[t, 0] = p["imL"]; img_host[t, 1] = p["imR"]
                kpL_host[t, :p["nL"]] = p["kpL"]; kpR_host[t, :p["nR"]] = p["kpR"]
                nL_host[t] = p["nL"]; nR_host[t] = p["nR"]
            img_b, kp_b = 2 * synth.H * synth.W, capq * 8

        # Two sequence objects on two contexts (each with its own compute and copy streams) take alternate steps, so the
        # tail of step k (the kernels of its last chunk, the record read-back) overlaps the uploads of step k+1 and
        # the PCIe link never idles.  Every step still uploads all of its inputs and reads its own records back inside
        # the timed region (the read-back of step k is issued after step k+1 has been enqueued).
        lanes = [(ctx, seq)]
        if args.e2e_buffers > 1:
            ctx2 = api.Context(local_rank)
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
        rec_pins = [rec_pin] + [torch.zeros(F * 16, dtype=torch.int32).pin_memory() for _ in lanes[1:]]

        def enqueue(i):
            c, sq = lanes[i % len(lanes)]
            c._ck(api.lib().viso_seq_set_seeds(sq.h, api._p(seeds_pin.data_ptr()), H))
            for t0 in range(0, F, args.chunk):   # chunked pipeline: uploads of chunk k+1 overlap the kernels of chunk k
                t1 = min(F, t0 + args.chunk)
                if use_raw:
                    sq.upload_chunk_raw(t0, t1 - t0, img_host.data_ptr() + t0 * img_b)
                elif use_img:
                    sq.upload_chunk_images_raw(t0, t1 - t0, img_host.data_ptr() + t0 * img_b, kpL_host.data_ptr() + t0 * kp_b,
                                               nL_host.data_ptr() + 4 * t0, kpR_host.data_ptr() + t0 * kp_b,
                                               nR_host.data_ptr() + 4 * t0)
                else:
                    p0 = sq
                    for t in range(t0, t1):
                        p = pinned[order[t]]
                        p0.upload_frame_raw(t, p["kpL"].data_ptr(), p["nL"], p["kpR"].data_ptr(), p["nR"],
                                            p["dL"].data_ptr(), p["dR"].data_ptr())
                sq.run_range(param, t0, t1)

        def collect(i):
            lanes[i % len(lanes)][1].download_raw(rec_pins[i % len(lanes)].data_ptr())   # synchronises that lane only

        def e2e_run(n):
            for i in range(n):
                enqueue(i)
                if i >= len(lanes) - 1:
                    collect(i - (len(lanes) - 1))
            for i in range(max(0, n - (len(lanes) - 1)), n):
                collect(i)

        e2e_run(2 * len(lanes))
        for c, _ in lanes:
            c.sync()
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        for c, _ in lanes:
            c.sync()
        e2e_ms = 1e3 * (time.perf_counter() - t0)   # host clock around fully synchronised work on both lanes
        barrier()
        for rp in rec_pins:
            rec_e2e = np.frombuffer(rp.numpy().tobytes(), dtype=api.RECORD_DTYPE)
            assert rec_e2e.tobytes() == rec.tobytes(), "e2e records differ from the resident run"
        for c2, s2 in lanes[1:]:
            s2.close()
            c2.close()

    clocks = sampler.stop()  # sampled over both timed regions (device-resident and end-to-end)

    # ---- max over ranks, gather records (the only collective: 64 B per frame pair) ----
    tms = torch.tensor([dev_ms, e2e_ms or 0.0, float(np.mean(match_ms))], dtype=torch.float64, device="cuda")
    # every rank's own times and clocks (the sequences differ per rank, so does their work)
    mine = torch.tensor([dev_ms / args.steps, (e2e_ms or 0.0) / args.steps, float(np.mean(match_ms)), float(sad_pairs),
                         float(clocks.get("sm_mhz") or 0.0)], dtype=torch.float64, device="cuda")
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, mine)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    else:
        per_rank = [mine]
    from libviso_b200.distributed import gather_records
    got = gather_records({rank: rec}, world, F, rank, world, device="cuda")   # sequence r lives on rank r
    all_rec = [got[s] for s in range(world)] if rank == 0 else None
    dev_ms, e2e_ms_g, mm = [float(v) for v in tms.cpu()]

    if rank == 0:
        n_poses = [len(api.chain_poses(r)) for r in all_rec]
        value = world * n_pairs * args.steps / (dev_ms * 1e-3)
        peak, peak_src = hbm_peak()
        achieved = match_bytes / (mm * 1e-3) / 1e9
        line = {
            "metric": "frame-pairs/s (match+RANSAC pose) at 1241x376", "value": value, "unit": "frame-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16 SAD / f64 pose",
            "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "gpu_launches": int(launches),
            "per_rank": {k: [round(float(t[i]), 3) for t in per_rank]
                         for i, k in enumerate(("ms_per_step", "e2e_ms_per_step", "sad_ms", "sad_pairs", "sm_mhz"))},
            "roofline": {"bound": "hbm", "kernel": "sad_match_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args), "peak_source": peak_src,
                         "note": "HBM fraction as the contract defines it; the kernel is bound by the L1 data pipe (66% of peak "
                                 "wavefronts in the committed ncu capture) and the ALU pipe (51%), not by HBM: DESIGN.md section 4",
                         "algorithmic_bytes_per_launch": int(match_bytes), "kernel_ms": mm,
                         "sad_pairs_per_launch": int(sad_pairs), "queries_left_to_generic_kernel": int(n_pending), "sad_evaluated_per_launch": int(sad_eval),
                         "kernel_share_of_step": mm / (dev_ms / args.steps),
                         # the roof of the kernel's own access pattern: the same row gather without arithmetic,
                         # measured once on B200 by tools/ubench_rowgather.cu (profiles/r01_e_rowgather_roof.txt)
                         "gather_roof": {"rows_per_s": 80.1e9, "achieved_rows_per_s": sad_pairs / (mm * 1e-3),
                                         "frac": sad_pairs / (mm * 1e-3) / 80.1e9,
                                         "source": "tools/ubench_rowgather.cu, measured on B200 in round 1 (not re-measured by this run)"}},
            "poses": {"chained_per_sequence": n_poses, "ok_frame_pairs": int(sum(int(r["ok"].sum()) for r in all_rec)),
                      "circular_matches_mean": float(np.mean([r["n_circ"][1:].mean() for r in all_rec])),
                      "inliers_mean": float(np.mean([r["n_inliers"][1:].mean() for r in all_rec])),
                      # RANSAC scoring work of one launch on rank 0: hypotheses x circular matches, 38 flops each
                      # (SURVEY 8d); FP64 rates to hold it against: tools/ubench_fp64.cu
                      "ransac_point_tests_per_launch": int(args.hyp * int(all_rec[0]["n_circ"][1:].sum()))},
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": world * n_pairs * args.steps / (e2e_ms_g * 1e-3), "unit": "frame-pairs/s",
                           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                           "ms_per_step": e2e_ms_g / args.steps}
        if world == 1 and not args.no_cpu:
            from oracle import oracle
            oracle.build()
            npairs = min(args.cpu_pairs, n_pairs)
            secs, rec_o = cpu_oracle_pairs(frames, order, 0, npairs, H, seeds)
            same = all(np.array_equal(rec[k][:npairs + 1], rec_o[k]) for k in ("ok", "n_inliers", "n_circ", "best_hyp"))
            line["cpu_baseline"] = {"value": npairs / secs, "unit": "frame-pairs/s", "cores": 1, "kind": "port",
                                    "sample": f"first {npairs} frame pairs of the same sequence, CPU oracle "
                                              f"(g++ -O2), {secs:.1f} s", "records_match_gpu": bool(same)}
        guard.emit(json.dumps(line))
    seq.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
