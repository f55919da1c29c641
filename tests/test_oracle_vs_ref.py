"""CPU, build container only: the oracle bit-for-bit against the reference itself (oracle/_ref/*): the
reference's host objects behind a no-op GL shim, and its fshader.glsl compiled as C++ through its own GLM.
Skipped where the _ref build is absent."""
import os

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol

pytestmark = [pytest.mark.ref,
              pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref not built (needs /root/reference)")]


@pytest.fixture(scope="module")
def ref_shader(default_level):
    rs = ol.RefShader()
    rs.upload(np.ascontiguousarray(default_level))
    return rs


def test_random_rays_bit_exact(oracle, ref_shader, default_level):
    starts, dirs, dists = gc.kat_rays(3000, seed=99)
    ret, out7 = ref_shader.cast_rays(starts, dirs, dists)
    for i in range(len(dists)):
        r, hp, hn, st = oracle.cast_ray(default_level, gc.DIMS, starts[i], dirs[i], dists[i])
        assert r == ret[i], i
        mine = np.concatenate([hp, hn, [np.float32(st)]]).astype(np.float32)
        assert np.array_equal(mine.view(np.uint32), out7[i].view(np.uint32)), i


@pytest.mark.parametrize("name", ["C2", "C3ii_pitched", "sparse_lights"])
@pytest.mark.parametrize("size", [(128, 72), (100, 37)])
def test_frames_bit_exact(oracle, ref_shader, default_level, name, size):
    W, H = size
    fr = gc.frame_cases(W, H)[name]
    ref_shader.set_frame(fr)
    rgba, steps = ref_shader.render(W, H)
    out = oracle.render(default_level, gc.DIMS, fr, W, H, want_f32=True)
    assert np.array_equal(rgba.view(np.uint32), out["rgba_f32"].view(np.uint32))
    assert float(steps.sum(dtype=np.float64)) == float(out["counters"][3])


def test_overbright_clamp_light_sets_bit_exact(oracle, ref_shader, default_level):
    """the light sets of tests/test_wide_oracle.py (clamp ending the loop in the first half / the second half / never, gaps between
    the slots, the last active slot not slot 15, odd counts, negative / infinite / NaN weights) through the REFERENCE's shader:
    what the early-out and the inactive slots do is pinned on the reference itself, not only on the restatement"""
    import test_wide_oracle as two
    W, H = 96, 54
    aspect = np.float32(W) / np.float32(H)
    for name, ls in two.light_sets().items():
        fr = ol.make_frame(gc.CAM, aspect=aspect, lights=ls)
        ref_shader.set_frame(fr)
        rgba, _ = ref_shader.render(W, H)
        out = oracle.render(default_level, gc.DIMS, fr, W, H, want_f32=True)["rgba_f32"]
        same = rgba.view(np.uint32) == out.view(np.uint32)
        both_nan = np.isnan(rgba) & np.isnan(out)
        assert bool((same | both_nan).all()), (name, int((~(same | both_nan)).any(axis=2).sum()))


def test_random_poses_bit_exact(oracle, ref_shader, default_level):
    rs = np.random.RandomState(5)
    rh = ol.RefHost()
    W, H = 64, 36
    for k in range(6):
        rot, _ = rh.mouse_look(float(rs.uniform(-1.5, 1.5)), float(rs.uniform(-3.1, 3.1)))
        cam = (float(rs.uniform(20, 490)), float(rs.uniform(38, 90)), float(rs.uniform(20, 490)))
        lights = [(cam[0] + float(rs.uniform(-40, 40)), float(rs.uniform(37, 60)), cam[2] + float(rs.uniform(-40, 40)), float(rs.uniform(0.1, 1.0)))
                  for _ in range(int(rs.randint(0, 17)))]
        fr = ol.make_frame(cam, rotate=rot, aspect=np.float32(W) / np.float32(H), lights=lights, view=int(k == 5))
        ref_shader.set_frame(fr)
        rgba, _ = ref_shader.render(W, H)
        out = oracle.render(default_level, gc.DIMS, fr, W, H, want_f32=True)
        assert np.array_equal(rgba.view(np.uint32), out["rgba_f32"].view(np.uint32)), k


def test_pitched_matrix_is_what_the_reference_builds():
    rot, d = ol.RefHost().mouse_look(0.5, 0.6)
    assert [float(x) for x in rot] == gc.PITCHED_ROTATE
    import voxel_rt_b200
    assert voxel_rt_b200.scenes.PITCHED_ROTATE == gc.PITCHED_ROTATE


@pytest.mark.skipif(not os.path.isdir(ol.REFERENCE_SRC), reason="reference tree absent")
def test_host_half_fix_depth_and_remove_sphere(oracle):
    """small edits through the reference's own removeSphere / fixDepthField vs the restatement, incl. cells next
    to the grid faces (the out-of-grid = solid rule)"""
    rh = ol.RefHost()
    v = rh.level_nodepth()
    mine = v.copy()
    for (x, y, z) in [(100, 40, 100), (0, 40, 0), (511, 95, 511), (3, 37, 300), (200, 50, 200), (256, 94, 10)]:
        rh.L.ref_host_fix_depth_field(x, y, z)
        oracle.fix_depth_field(mine, gc.DIMS, x, y, z)
    assert np.array_equal(rh.voxels(), mine)
    for (x, y, z, r) in [(150, 36, 150, 7), (2, 35, 200, 5), (509, 30, 509, 7), (300, 93, 40, 4)]:
        rh.L.ref_host_remove_sphere(x, y, z, r)
        oracle.remove_sphere(mine, gc.DIMS, x, y, z, r)
        assert np.array_equal(rh.voxels(), mine), (x, y, z, r)
    for s, e in [((180.0, 25.0, 140.0), (210.0, 55.0, 170.0)), ((210.5, 55.0, 170.0), (180.0, 25.5, 140.0)),
                 ((-10.0, 20.0, 100.0), (20.0, 50.0, 130.0)), ((490.0, 80.0, 490.0), (520.0, 110.0, 520.0))]:
        off, size, n = rh.partial(s, e)
        first, count, n2 = oracle.partial_ranges(gc.DIMS, s, e)
        assert n == n2
        assert np.array_equal(off, first * 4) and np.array_equal(size, count.astype(np.int64) * 4)
