"""The C++ host layer (libviso_b200/host: the reference's own API over the C-ABI) and the drop-in proof.

* tests/host/test_host.cpp -- written against libviso_b200/host/viso.h: the reference's test_nl_rigid_motion1
  (test/test.cpp:152-168) and the per-frame loop of sequence_odometry (viso.cpp:1240-1321), every result compared
  with the oracle in C++.
* build/dropin/ref_kitti, build/dropin/ref_tester -- the REFERENCE'S OWN src/kitti.cpp and test/test.cpp, compiled
  unchanged from the read-only reference tree (tools/build_dropin.py, run by __graft_entry__.build() where the tree
  exists) against this library.  Here (CPU) they are built and must refuse to run without a GPU; on the GPU box the
  prebuilt binaries run end to end and are checked against the oracle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import build_dropin  # noqa: E402

COMPAT = os.path.join(ROOT, "compat")


def build_host_test(tmp_path, api, oracle):
    exe = str(tmp_path / "test_host")
    lib_dir = os.path.join(ROOT, "libviso_b200")
    ora_dir = os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + COMPAT, os.path.join(ROOT, "tests", "host", "test_host.cpp"),
           os.path.join(lib_dir, "host", "viso.cpp"), "-L" + lib_dir, "-lviso_b200", "-L" + ora_dir, "-lviso_oracle",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + ora_dir, "-lz", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_host_layer_compiles_and_refuses_without_gpu(tmp_path, api, oracle):
    import torch
    exe = build_host_test(tmp_path, api, oracle)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_host_layer_matches_oracle")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_host_layer_matches_oracle(tmp_path, api, oracle):
    exe = build_host_test(tmp_path, api, oracle)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "host test OK" in r.stdout


def test_kitti_io_formats(tmp_path):
    """loadCalib / savePoses (reference src/kitti.cpp:23-64): KITTI calib.txt in, 12-number pose lines out"""
    exe = str(tmp_path / "test_kitti_io")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I" + COMPAT, os.path.join(ROOT, "tests", "host", "test_kitti_io.cpp"),
                           "-o", exe])
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "kitti_io OK" in r.stdout, r.stdout + r.stderr


# ------------------------------------------------------------------------------------------------ the reference's callers

def dropin(api):
    """the reference's kitti.cpp / test.cpp built unchanged against this library; prebuilt files where the tree is absent"""
    built = build_dropin.build()
    t = build_dropin.targets()
    if built is None and not all(os.path.exists(p) for p in t):
        pytest.skip("the reference tree is not here and build/dropin holds no prebuilt binaries")
    return t


def write_kitti_tree(home, frames, P1, P2):
    """$KITTI_HOME/sequences/00/{calib.txt, image_0/%06d.png, image_1/%06d.png}: the names the reference's driver opens
    (kitti.cpp:108-110).  The files hold binary PGM data: the stand-in cv::imread decodes by content (libpng is not
    installed here; with a real OpenCV write real PNGs)."""
    seq = home / "sequences" / "00"
    for d in ("image_0", "image_1"):
        (seq / d).mkdir(parents=True)
    with open(seq / "calib.txt", "w") as f:
        f.write("P0: " + " ".join(repr(float(v)) for v in P1.reshape(-1)) + "\n")
        f.write("P1: " + " ".join(repr(float(v)) for v in P2.reshape(-1)) + "\n")
    for t, fr in enumerate(frames):
        for d, im in (("image_0", fr["imL"]), ("image_1", fr["imR"])):
            with open(seq / d / ("%06d.png" % t), "wb") as f:
                f.write(b"P5\n# synthetic\n%d %d\n255\n" % (im.shape[1], im.shape[0]))
                f.write(im.tobytes())


def test_reference_callers_build_unchanged_and_refuse_without_gpu(tmp_path, api, small_sequence):
    """src/kitti.cpp and test/test.cpp of the reference compile and link against libviso_b200/host as they are"""
    import torch
    from libviso_b200 import synth
    kitti, tester, _ = dropin(api)
    assert os.access(kitti, os.X_OK) and os.access(tester, os.X_OK)
    r = subprocess.run([tester], capture_output=True, text=True)   # no data file: the reference's test exits 0 (test.cpp:119-122)
    assert r.returncode == 0 and "Running 1 test case" in r.stdout
    if torch.cuda.is_available():
        return
    frames, _ = small_sequence
    write_kitti_tree(tmp_path, frames[:2], *synth.kitti_calib())
    r = subprocess.run([kitti, "sha", "00"], capture_output=True, text=True, env=dict(os.environ, KITTI_HOME=str(tmp_path)))
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_reference_kitti_driver_matches_oracle(tmp_path, api, oracle, small_sequence):
    """the reference's driver (kitti.cpp:79-118, compiled unchanged) end to end on this library: calib.txt + image files
    in, pose file out, against the oracle's front end + pipeline + pose chaining on the same images and sample seeds"""
    from libviso_b200 import synth
    kitti, _, _ = dropin(api)
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    write_kitti_tree(tmp_path, frames, P1, P2)
    r = subprocess.run([kitti, "sha", "00", "1", "4"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, KITTI_HOME=str(tmp_path), VISO_LOG_LEVEL="warning"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = np.loadtxt(tmp_path / "results" / "00" / "sha" / "data" / "00.txt").reshape(-1, 12)
    sel = frames[1:5]                                      # begin 1, end 4 inclusive (viso.h:88)
    oframes = oracle.frames_from_images([(f["imL"], f["imR"]) for f in sel], 1200)   # MAX_FEATURE_NUM, viso.cpp:1172
    H = 50
    # the host layer's seed stream: one std::mt19937(424242), 32 bits per draw, frame-major
    mt = np.random.MT19937(); mt._legacy_seeding(424242)
    seeds = mt.random_raw(len(sel) * H * 3).astype(np.uint32).reshape(len(sel), H, 3)
    o = oracle.sequence(oframes, P1, P2, oracle.param_default(ransac_iter=H), seeds)
    want = o["poses"][:, :3, :].reshape(-1, 12)
    assert got.shape == want.shape and len(got) == len(sel)
    assert np.abs(got - want).max() < 2e-6                 # "%lf": six decimals


@pytest.mark.gpu
def test_reference_test_suite_runs_on_real_data(tmp_path, api, oracle):
    """the reference's Boost.Test case test_nl_rigid_motion1 (test/test.cpp:152-168, compiled unchanged) with a data
    file in its CSV format (test.cpp:124-139): ransac_minimize_reproj must return true, and the tr it prints must be
    the oracle's for the same calibration (test.cpp:158-161) and the same sample stream"""
    from libviso_b200 import synth
    _, tester, redirect = dropin(api)
    base, f, cu, cv = .5707, 645.24, 635.96, 194.13        # test.cpp:158-161
    rng = np.random.default_rng(12)
    n = 400
    X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-2, 3, n), rng.uniform(4, 60, n)])
    tr = np.array([0.01, -0.02, 0.005, 0.05, -0.02, -1.0])
    T = oracle.tr2mat(tr)
    Xc = T[:3, :3] @ X + T[:3, 3:4]
    obs = np.stack([f * Xc[0] / Xc[2] + cu, f * Xc[1] / Xc[2] + cv, f * (Xc[0] - base) / Xc[2] + cu, f * Xc[1] / Xc[2] + cv])
    obs += rng.standard_normal(obs.shape) * 0.3
    bad = rng.random(n) < 0.3
    obs[:, bad] += rng.uniform(-50, 50, (4, int(bad.sum())))
    csv = tmp_path / "data.csv"
    with open(csv, "w") as fh:
        fh.write("%d\n" % n)
        for i in range(n):
            fh.write(" ".join(repr(float(v)) for v in [0, 0, 0, 0, *obs[:, i], *X[:, i]]) + "\n")
    r = subprocess.run([tester], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, LD_PRELOAD=redirect, VISO_TEST_DATA_CSV=str(csv)))
    assert r.returncode == 0 and "No errors detected" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("tr: ")][0]
    got = np.array([float(v) for v in line.split()[1:]])
    # what was written is what the test reads back (17 significant digits survive repr / %lf)
    p = oracle.param_default(base=base, f=f, cu=cu, cv=cv, ransac_iter=50)
    o = oracle.ransac_minimize_reproj(X, obs, p, oracle.randomsample_table(424242, 50, n))
    assert o["ok"]
    assert np.abs(got - o["tr"]).max() <= 1e-5 * max(1.0, np.abs(o["tr"]).max())   # "%g": six significant digits
