"""CPU: the kernels in the built libvxrt.so are, instruction for instruction, the ones whose B200 measurements are committed
under profiles/ (tests/golden/sass_fingerprints.json, made by tests/golden/make_sass_fingerprints.py).  Experiments live
behind template parameters / macros that leave the production kernels' SASS untouched; changing a production kernel means
regenerating the fingerprints on purpose -- and measuring again."""
import json
import os
import re
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_production_kernels_are_the_measured_ones(vx):
    if not (shutil.which("cuobjdump") and shutil.which("c++filt") and shutil.which("nvcc")):
        pytest.skip("cuobjdump / c++filt / nvcc unavailable")
    import make_sass_fingerprints as msf
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "sass_fingerprints.json")))
    if msf.nvcc_version() != want["nvcc"]:
        pytest.skip("fingerprints were made with nvcc %s" % want["nvcc"])
    got = msf.fingerprints(vx.build.LIB)
    assert set(got) == set(want["kernels"])
    changed = sorted(k for k in got if got[k] != want["kernels"][k])
    assert not changed, changed
    assert want.get("measured"), "the fingerprint file names the profiles/ entry its kernels were measured in"
