"""Per-kernel fingerprints of the SASS in voxel-rt_b200/libvxrt.so (instruction text without addresses and encodings).
tests/test_sass_fingerprint.py compares the built library with tests/golden/sass_fingerprints.json, so that a change to
the kernels that were measured on the B200 (profiles/) is a deliberate act: re-run this script when one is intended
and re-measure.  Usage: python tests/golden/make_sass_fingerprints.py [--write]"""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "voxel-rt_b200", "libvxrt.so")
OUT = os.path.join(ROOT, "tests", "golden", "sass_fingerprints.json")


def fingerprints(lib=LIB):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?)\s*;", line)
        if cur and m:
            out[cur].append(m.group(1))
    return {k: {"instructions": len(v), "sha1": hashlib.sha1("\n".join(v).encode()).hexdigest()} for k, v in out.items()}


def nvcc_version():
    out = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout
    m = re.search(r"release [\d.]+, V([\d.]+)", out)
    return m.group(1) if m else "unknown"


if __name__ == "__main__":
    measured = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--measured=")]
    fp = {"nvcc": nvcc_version(), "measured": measured[0] if measured else "not yet measured", "kernels": fingerprints()}
    if "--write" in sys.argv:
        with open(OUT, "w") as f:
            json.dump(fp, f, indent=1, sort_keys=True)
        print("wrote", OUT, len(fp["kernels"]), "kernels")
    else:
        print(json.dumps(fp, indent=1, sort_keys=True))
