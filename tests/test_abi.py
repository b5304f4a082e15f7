"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/viso_b200.h declares,
refuses to work without a GPU (no CPU fallback), and its host bookkeeping agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "viso_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(viso_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(api):
    lib = api.lib()
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/viso_b200.h but not exported"
    assert lib.viso_abi_version() == 1


def test_struct_layouts(api):
    assert C.sizeof(api.MatchParams) == 16 + 3 * 8 + 72
    assert C.sizeof(api.Param) == 56
    assert api.RECORD_DTYPE.itemsize == 64


def test_no_cpu_fallback(api):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.VisoError):
        api.Context(0)


def test_product_does_not_touch_oracle():
    """the product package must not import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "libviso_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "viso_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_host_bookkeeping_matches_oracle(api, oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        tr = rng.standard_normal(6) * np.array([0.1, 0.1, 0.1, 1, 1, 1])
        assert np.array_equal(api.tr2mat(tr), oracle.tr2mat(tr))
        pose = oracle.tr2mat(rng.standard_normal(6) * 0.3)
        ok1, p1 = api.pose_update(pose, tr)
        ok2, p2 = oracle.pose_update(pose, tr)
        assert ok1 and ok2 and np.array_equal(p1, p2)
    from libviso_b200 import synth
    P1, P2 = synth.kitti_calib()
    assert np.array_equal(api.F_from_P(P1, P2), oracle.F_from_P(P1, P2))
    assert np.array_equal(api.F_from_P(P1, P2, False), oracle.F_from_P(P1, P2, False))
    assert np.array_equal(api.randomsample_table(424242, 64, 1000), oracle.randomsample_table(424242, 64, 1000))
    seeds = rng.integers(0, 2 ** 32, size=(500, 3), dtype=np.uint32)
    for N in (3, 4, 7, 300, 10000):
        t = api.samples_from_seeds(seeds, N)
        assert np.array_equal(t, oracle.samples_from_seeds(seeds, N))
        assert (t[:, 0] < t[:, 1]).all() and (t[:, 1] < t[:, 2]).all() and t.min() >= 0 and t.max() < N


def test_chain_poses_matches_oracle(api, oracle):
    rng = np.random.default_rng(4)
    rec = np.zeros(12, api.RECORD_DTYPE)
    rec["tr"] = rng.standard_normal((12, 6)) * 0.05
    rec["ok"] = rng.random(12) < 0.8
    rec["n_circ"] = np.where(rng.random(12) < 0.9, 100, 2)
    rec["ok"][0] = 0
    poses = api.chain_poses(rec)
    pose = np.eye(4); want = [pose]
    for t in range(1, 12):
        if rec["ok"][t] and rec["n_circ"][t] >= 3:
            _, pose = oracle.pose_update(pose, rec["tr"][t])
            want.append(pose)
    assert np.array_equal(poses, np.stack(want))
