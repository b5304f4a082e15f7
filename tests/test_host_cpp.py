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
