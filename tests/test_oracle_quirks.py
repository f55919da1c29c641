"""CPU: quirks of the reference's castRay that shape what the CUDA path may and may not skip, pinned on the oracle (and on
the compiled reference shader where oracle/_ref exists), plus the analysis profile scripts/where_iterations_go.py uses."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# A local-light ray of config C2 (1920x1080): from the ground (y = 37) toward the light at (183, 40, 151).
TIE_START = (260.999146, 37.0000458, 194.75412)
TIE_DIR = (-0.874806762, 0.0486004092, -0.482028097)
TIE_DIST = 62


def test_tie_lock_sends_a_ray_straight_down_the_z_axis(oracle, default_level):
    """fshader.glsl:87-104: `if (x < y && x < z) ... else if (y < x && y < z) ... else z`.  When intersect.x == intersect.y
    and both are smaller than intersect.z, neither strict test holds, the else branch steps z -- which leaves x and y
    tied, so EVERY following iteration steps z too, until a depth-field jump re-bases the ray.  The cells such a ray tests
    leave the geometric line (here by 60 cells), which is why no occupancy structure can prove a castRay miss from the
    geometry of the ray alone ("what a finer hierarchy could still save")."""
    r, hit_pos, normal, steps = oracle.cast_ray(default_level, gc.DIMS, TIE_START, TIE_DIR, TIE_DIST)
    assert (r, steps) == (7391987, 61.0)
    x, y, z = r % 512, (r // 512) % 96, r // (512 * 96)
    assert (x, y, z) == (243, 37, 150) and default_level[r] >= 0          # a tree trunk 60 cells off the line ...
    assert abs(hit_pos[0] - 181.592) < 1e-3 and abs(hit_pos[2] - 150.9999) < 1e-3     # ... while hitPos stays on it
    assert list(normal) == [0.0, 0.0, 1.0]
    # the geometric ray itself is unobstructed: nudged by one ulp it reaches the light (budget exhausted, no hit)
    nudged = (TIE_START[0], float(np.nextafter(np.float32(TIE_START[1]), np.float32(100))), TIE_START[2])
    r2, _, _, steps2 = oracle.cast_ray(default_level, gc.DIMS, nudged, TIE_DIR, TIE_DIST)
    assert r2 == -1 and steps2 > 40


@pytest.mark.ref
def test_tie_lock_is_the_reference_shaders_behaviour(default_level):
    if not os.path.exists(os.path.join(ol.REF_DIR, "libref_shader.so")):
        pytest.skip("oracle/_ref not built")
    rs = ol.RefShader()
    rs.upload(default_level)
    ret, out7 = rs.cast_rays(np.array([TIE_START], np.float32), np.array([TIE_DIR], np.float32), np.array([TIE_DIST], np.int32))
    assert ret[0] == 7391987 and out7[0][6] == 61.0 and abs(out7[0][0] - 181.592) < 1e-3


def test_iteration_profile_accounts_for_every_ray(oracle, default_level):
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import where_iterations_go as wig
    W, H = 320, 180
    fr = gc.frame_cases(W, H)["C2"]
    pr = wig.profile(oracle, np.ascontiguousarray(default_level), gc.DIMS, fr, W, H)
    ref = oracle.render(default_level, gc.DIMS, fr, W, H)
    rays = {k: sum(c["rays"] for n, c in pr["cells"].items() if n.startswith(k + "/")) for k in wig.KINDS}
    assert [rays[k] for k in wig.KINDS] == [int(v) for v in ref["counters"][:3]]
    assert sum(c["iterations"] for c in pr["cells"].values()) == int(ref["counters"][3])
    assert pr["cells"]["primary/hit"]["rays"] == int(ref["counters"][4])
    assert (pr["ymin"], pr["ymax"]) == (0, 51)
    s = wig.summarise(pr)
    assert s["removed_total"] + s["still_executed"] == s["iterations"] and 0 < s["still_executed_by_misses"] < s["still_executed"]
