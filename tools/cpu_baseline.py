#!/usr/bin/env python
"""CPU baseline table of BASELINE.md section 2: the oracle (restated reference path) on the host cores.

  1 core -O2 (headline), 1 core -O0 (the reference's CMake sets no optimisation level), and min(8, nproc) independent
  frame ranges concurrently at -O2.  1 warm-up + 5 timed repetitions, median and min.  Prints one JSON object.
"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKER = r'''
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
from oracle import oracle
from libviso_b200 import synth
frames, _ = synth.make_sequence(int(sys.argv[1]), seed=1000, n_features=2040, workers=1)
P1, P2 = synth.kitti_calib()
H = 50
seeds = np.random.default_rng(1).integers(0, 2 ** 32, size=(len(frames), H, 3), dtype=np.uint32)
prm = oracle.param_default(ransac_iter=H)
ts = []
for rep in range(int(sys.argv[2])):
    t = time.perf_counter(); oracle.sequence(frames, P1, P2, prm, seeds); ts.append(time.perf_counter() - t)
print(json.dumps(ts))
''' % ROOT


def run(n_frames, reps, so=None, procs=1):
    env = dict(os.environ)
    if so:
        env["VISO_ORACLE_SO"] = so
    ps = [subprocess.Popen([sys.executable, "-c", WORKER, str(n_frames), str(reps)], stdout=subprocess.PIPE, env=env)
          for _ in range(procs)]
    outs = [json.loads(p.communicate()[0]) for p in ps]
    return outs


def main():
    n_frames, reps = 9, 6
    pairs = n_frames - 1
    o2 = run(n_frames, reps)[0][1:]
    o0 = run(n_frames, reps, so=os.path.join(ROOT, "oracle", "libviso_oracle_O0.so"))[0][1:]
    procs = min(8, os.cpu_count() or 1)
    multi = run(n_frames, reps, procs=procs)
    agg = [procs * pairs / max(m[i] for m in multi) for i in range(1, reps)]
    med = lambda v: sorted(v)[len(v) // 2]
    print(json.dumps({
        "workload": f"{pairs} frame pairs of the synthetic KITTI-shaped sequence (~2040 features, 50 hypotheses), CPU oracle",
        "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else "?",
        "nproc": os.cpu_count(),
        "O2_1core_frame_pairs_per_s": {"median": pairs / med(o2), "best": pairs / min(o2)},
        "O0_1core_frame_pairs_per_s": {"median": pairs / med(o0), "best": pairs / min(o0)},
        f"O2_{procs}procs_frame_pairs_per_s": {"median": med(agg), "best": max(agg)},
    }, indent=1))


if __name__ == "__main__":
    main()
