"""CPU: bench.py's reference arm (`--impl reference`: the reference's own shader on the host cores, no GPU involved) prints ONE JSON
line with the contract's keys; the B200 arm refuses to run without a GPU (no CPU fallback, no oracle behind the product path)."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT, env=dict(os.environ, **(env or {})))


@pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_arm_line_has_the_contract_keys():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "C1_720p")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/sec (primary+shadow)" and d["unit"] == "Mrays/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"] == "C1_720p" and d["config"]["width"] == 1280
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "fshader.glsl" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_arm_under_a_multi_rank_launch_only_rank_0_works():
    """the driver launches both arms the same way; under torchrun every rank but 0 exits 0 without work"""
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload", "C1_720p",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout[-500:], r.stderr[-500:])


def test_b200_arm_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except Exception:
        pass
    r = run_bench("--steps", "1", "--warmup", "1", "--no-extra")
    assert r.returncode != 0
    assert "oracle" not in r.stdout.lower()
    assert not any(l.startswith("{") for l in r.stdout.splitlines())          # no JSON line: nothing was measured
