import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

CACHE_DIR = os.environ.get("VXRT_CACHE", "/tmp/vxrt_cache")
LEVEL_FNV_NODEPTH = 0x2f8d49bd81549f5a          # SURVEY.md 8c (from the reference's own host code)
LEVEL_FNV_DEPTH = 0x4c58cc4001a22afa


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "ref: needs the reference build under oracle/_ref")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.build_oracle()
    return oracle_lib.Oracle()


def load_default_level(oracle):
    """reference default level WITH depth field (fingerprint 4c58cc4001a22afa); cached on disk because the
    depth-field sweep costs ~15 s of CPU."""
    os.makedirs(CACHE_DIR, exist_ok=True)
    path = os.path.join(CACHE_DIR, "default_level_depth.npy")
    if os.path.exists(path):
        v = np.load(path)
        if v.size == 512 * 96 * 512 and oracle.fnv(v) == LEVEL_FNV_DEPTH:
            return v
    v = oracle.default_level(depth_field=True)
    assert oracle.fnv(v) == LEVEL_FNV_DEPTH
    tmp = path + ".%d.tmp.npy" % os.getpid()
    np.save(tmp, v)
    os.replace(tmp, path)
    return v


@pytest.fixture(scope="session")
def default_level(oracle):
    v = load_default_level(oracle)
    v.setflags(write=False)
    return v


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def vx():
    """the product package; on a GPU box the CUDA library must load (no fallback)"""
    import voxel_rt_b200
    voxel_rt_b200.load_library()
    return voxel_rt_b200
