"""libviso_b200/csrc/glibc_sincos.h against the libm the oracle links (glibc's sin / cos, which the reference calls at
viso.cpp:1410-1411): compiled as host code, the header must reproduce libm bit for bit over every argument range the
estimation can visit.  The device build of the same header is checked in tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def build_replica(directory):
    so = os.path.join(str(directory), "libsincos_replica.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-builtin-sin", "-fno-builtin-cos", "-mfma", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "host", "sincos_replica.cpp"), "-o", so])
    return C.CDLL(so)


def libm_sincos(lib, x):
    """glibc's sin() and cos() (not numpy's vectorised kernels, which may differ in the last place)"""
    x = np.ascontiguousarray(x, np.float64)
    s = np.empty_like(x); c = np.empty_like(x)
    lib.libm_sincos(x.ctypes.data_as(C.c_void_p), len(x), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p))
    return s, c


@pytest.fixture(scope="module")
def replica(tmp_path_factory):
    return build_replica(tmp_path_factory.mktemp("sc"))


def sincos_arguments(seed=0, n=400000):
    rng = np.random.default_rng(seed)
    parts = [rng.uniform(-0.126, 0.126, n), rng.uniform(-0.855469, 0.855469, n), rng.uniform(-2.426265, 2.426265, n),
             rng.uniform(-40, 40, n), rng.uniform(-1e4, 1e4, n), rng.uniform(-1.05414e8, 1.05414e8, n),
             rng.standard_normal(n) * 1e-3, rng.standard_normal(n) * 1e-9, 10.0 ** rng.uniform(-300, 8, n) * rng.choice([-1, 1], n),
             np.array([0.0, -0.0, 2.0 ** -27, 2.0 ** -26, 0.126, 0.855469, 2.426265, np.pi / 2, np.pi, 105414349.9, 1e-310, -1e-310]),
             np.arange(-2000, 2000) * (np.pi / 4), np.arange(0, 110 * 8) / 1024.0]
    # the boundaries between the five ranges, a few ulps either side
    for b in (2.0 ** -26, 2.0 ** -27, 0.126, np.float64.fromhex("0x1.b6p-1"), np.float64.fromhex("0x1.368fdp+1"), 105414350.0):
        parts.append(np.nextafter(b, np.inf) + np.arange(-8, 9) * np.spacing(b))
    return np.ascontiguousarray(np.concatenate(parts))


def test_gen_table_matches_local_libm():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_sincostab.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_replica_is_bit_identical_to_libm(replica):
    x = sincos_arguments()
    s0 = np.empty_like(x); c0 = np.empty_like(x); s1 = np.empty_like(x); c1 = np.empty_like(x)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    replica.libm_sincos(p(x), len(x), p(s0), p(c0))
    replica.replica_sincos(p(x), len(x), p(s1), p(c1))
    bad_s = np.nonzero(s0.view(np.int64) != s1.view(np.int64))[0]
    bad_c = np.nonzero(c0.view(np.int64) != c1.view(np.int64))[0]
    assert len(bad_s) == 0, (len(bad_s), x[bad_s[:5]], s0[bad_s[:5]], s1[bad_s[:5]])
    assert len(bad_c) == 0, (len(bad_c), x[bad_c[:5]], c0[bad_c[:5]], c1[bad_c[:5]])
    assert len(x) > 3_000_000
