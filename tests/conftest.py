import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """the CPU oracle (test infrastructure): builds oracle/libviso_oracle.so on first use"""
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def api():
    from libviso_b200 import build
    build.build()
    from libviso_b200 import api as a
    a.lib()
    return a


@pytest.fixture(scope="session")
def ctx(api):
    """a viso_ctx on cuda:0; the product has no CPU path, so this fails (not skips) without a GPU"""
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def small_sequence():
    """6 synthetic KITTI-shaped frames, ~600 features per image (fast enough for the CPU oracle)"""
    from libviso_b200 import synth
    frames, gt = synth.make_sequence(6, seed=1000, n_features=600)
    return frames, gt


def make_seeds(n_frames, H, seed=424242):
    return np.random.default_rng(seed).integers(0, 2 ** 32, size=(n_frames, H, 3), dtype=np.uint32)
