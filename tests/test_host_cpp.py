"""The C++ host layer (libviso_b200/host: the reference's own signatures over the C-ABI).

CPU: it compiles against the compat headers and links against the C-ABI library.
GPU: tests/host/test_host.cpp -- the reference's test_nl_rigid_motion1 (test/test.cpp:152-168) and the per-frame loop
of sequence_odometry (viso.cpp:1240-1321), every result compared with the oracle in C++."""
import os
import subprocess

import pytest

from conftest import ROOT


def build_host_test(tmp_path, api, oracle):
    exe = str(tmp_path / "test_host")
    lib_dir = os.path.join(ROOT, "libviso_b200")
    ora_dir = os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host", "test_host.cpp"),
           os.path.join(lib_dir, "host", "viso.cpp"), "-L" + lib_dir, "-lviso_b200", "-L" + ora_dir, "-lviso_oracle",
           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + ora_dir, "-o", exe]
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
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "host", "test_kitti_io.cpp"),
                           "-o", exe])
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "kitti_io OK" in r.stdout, r.stdout + r.stderr


def build_kitti_driver(tmp_path, api):
    exe = str(tmp_path / "kitti")
    lib_dir = os.path.join(ROOT, "libviso_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(lib_dir, "host", "kitti.cpp"),
                           os.path.join(lib_dir, "host", "viso.cpp"), "-L" + lib_dir, "-lviso_b200",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


def write_kitti_tree(home, frames, P1, P2):
    """$KITTI_HOME/sequences/00/{calib.txt, image_0/%06d.pgm, image_1/%06d.pgm}"""
    seq = home / "sequences" / "00"
    for d in ("image_0", "image_1"):
        (seq / d).mkdir(parents=True)
    with open(seq / "calib.txt", "w") as f:
        f.write("P0: " + " ".join(repr(float(v)) for v in P1.reshape(-1)) + "\n")
        f.write("P1: " + " ".join(repr(float(v)) for v in P2.reshape(-1)) + "\n")
    for t, fr in enumerate(frames):
        for d, im in (("image_0", fr["imL"]), ("image_1", fr["imR"])):
            with open(seq / d / ("%06d.pgm" % t), "wb") as f:
                f.write(b"P5\n# synthetic\n%d %d\n255\n" % (im.shape[1], im.shape[0]))
                f.write(im.tobytes())


def test_kitti_driver_refuses_without_gpu(tmp_path, api, small_sequence):
    import torch
    from libviso_b200 import synth
    exe = build_kitti_driver(tmp_path, api)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_kitti_driver_matches_oracle")
    frames, _ = small_sequence
    write_kitti_tree(tmp_path, frames[:2], *synth.kitti_calib())
    r = subprocess.run([exe, "sha", "00"], capture_output=True, text=True, env=dict(os.environ, KITTI_HOME=str(tmp_path)))
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_kitti_driver_matches_oracle(tmp_path, api, oracle, small_sequence):
    """the reference's driver (kitti.cpp:80-111) end to end: calib.txt + image files in, pose file out, against the
    oracle's front end + pipeline + pose chaining on the same images and the same sample seeds"""
    import numpy as np
    from libviso_b200 import synth
    exe = build_kitti_driver(tmp_path, api)
    frames, _ = small_sequence
    P1, P2 = synth.kitti_calib()
    write_kitti_tree(tmp_path, frames, P1, P2)
    r = subprocess.run([exe, "sha", "00", "1", "4"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, KITTI_HOME=str(tmp_path), VISO_MAX_FEATURES="1200"))
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.loadtxt(tmp_path / "results" / "00" / "sha" / "data" / "00.txt").reshape(-1, 12)
    sel = frames[1:5]                                      # begin 1, end 4 inclusive (viso.h:88)
    oframes = oracle.frames_from_images([(f["imL"], f["imR"]) for f in sel], 1200)
    H = 50
    # the host layer's seed stream: one std::mt19937(424242), 32 bits per draw, frame-major
    mt = np.random.MT19937(); mt._legacy_seeding(424242)
    seeds = mt.random_raw(len(sel) * H * 3).astype(np.uint32).reshape(len(sel), H, 3)
    o = oracle.sequence(oframes, P1, P2, oracle.param_default(ransac_iter=H), seeds)
    want = o["poses"][:, :3, :].reshape(-1, 12)
    assert got.shape == want.shape and len(got) == len(sel)
    assert np.abs(got - want).max() < 2e-6                 # "%lf": six decimals
