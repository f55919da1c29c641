"""CPU: size-independent properties of the restated algorithm (oracle) -- the same properties the GPU tests rely on
at full size: idempotence of the depth field, monotonicity of destruction, consistency of the per-pixel outputs."""
import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol


def small_grid(seed, dims=(40, 24, 36), fill=0.03):
    rs = np.random.RandomState(seed)
    n = dims[0] * dims[1] * dims[2]
    v = np.full(n, -1, np.int32)
    solid = rs.rand(n) < fill
    v[solid] = rs.randint(0, 1 << 24, int(solid.sum()))
    v.reshape(dims[2], dims[1], dims[0])[:, :5, :] = 0x445566
    return np.ascontiguousarray(v), dims


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_depth_field_values_and_idempotence(oracle, seed):
    v, dims = small_grid(seed)
    solid_before = v >= 0
    oracle.compute_depth_field(v, dims)
    assert np.array_equal(v >= 0, solid_before)                       # never creates or removes solids
    jumps = -v[v < -1].view(np.float32)
    allowed = {np.float32(np.sqrt(np.float64(k))) for k in range(4, 37)}
    assert set(np.unique(jumps)).issubset(allowed)                    # render.cpp:77-81,240-247: -sqrt(k), 2 <= jump <= 6
    again = v.copy()
    oracle.compute_depth_field(again, dims)
    assert np.array_equal(again, v)                                   # only the sign of neighbours is read
    # a jump never reaches a solid: every solid is at pre-shrunk distance >= jump (spot check along +x)
    g = v.reshape(dims[2], dims[1], dims[0])
    zs, ys, xs = np.nonzero(g < -1)
    for z, y, x in list(zip(zs, ys, xs))[:400]:
        j = -np.array([g[z, y, x]], np.int32).view(np.float32)[0]
        reach = int(np.floor(j))                                      # cells x+1 .. x+reach must be non-solid
        seg = g[z, y, x + 1: x + 1 + reach]
        assert (seg < 0).all()


def test_remove_sphere_is_monotone_and_local(oracle):
    v, dims = small_grid(7)
    oracle.compute_depth_field(v, dims)
    before = v.copy()
    c, r = (20, 8, 18), 6
    oracle.remove_sphere(v, dims, c[0], c[1], c[2], r)
    assert not ((v >= 0) & (before < 0)).any()                        # destruction never adds a solid
    g0, g1 = before.reshape(dims[2], dims[1], dims[0]), v.reshape(dims[2], dims[1], dims[0])
    changed = np.argwhere(g0 != g1)
    R = r + 3
    assert (np.abs(changed - np.array([c[2], c[1], c[0]])) <= R).all()   # level.cpp:43-55: only the r+3 box is touched
    again = v.copy()
    oracle.remove_sphere(again, dims, c[0], c[1], c[2], r)
    assert np.array_equal(again, v)                                   # idempotent


def test_partial_ranges_shapes(oracle):
    dims = gc.DIMS
    first, count, n = oracle.partial_ranges(dims, (180.0, 25.0, 140.0), (210.0, 55.0, 170.0))
    assert n == 900 and (count == 31).all() and len(np.unique(first)) == 900       # 30 x 30 rows of 31 voxels
    f2, c2, n2 = oracle.partial_ranges(dims, (210.0, 55.0, 170.0), (180.0, 25.0, 140.0))   # reversed corners are swapped
    assert n2 == n and np.array_equal(f2, first) and np.array_equal(c2, count)
    assert oracle.partial_ranges(dims, (-10.0, 25.0, 140.0), (20.0, 55.0, 170.0))[2] == 0     # start.x < 0: nothing (a13)


def test_pixel_outputs_are_consistent(oracle, default_level):
    W, H = 128, 72
    for name in ("C2", "C3ii_pitched", "sparse_lights"):
        fr = gc.frame_cases(W, H)[name]
        out = oracle.render(default_level, gc.DIMS, fr, W, H)
        hit = out["hit_index"] >= 0
        c = out["counters"]
        assert c[0] == W * H and c[1] == hit.sum() == c[4]
        assert ((out["cast_mask"] & 1) == hit).all()                  # a global-light ray per hit pixel, none otherwise
        assert (out["occl_mask"] & ~out["cast_mask"] == 0).all()      # only cast rays can be occluded
        assert c[2] == sum(int(bin(int(m) >> 1).count("1")) for m in out["cast_mask"].ravel())
        assert (default_level[out["hit_index"][hit]] >= 0).all()      # the hit voxel is solid
        assert (out["steps"] >= 1).all() and out["steps"].max() <= 384
        sky = np.array([153, 179, 204, 255], np.uint8)                # (0.6, 0.7, 0.8, 1) through the UNORM8 rule
        assert (out["rgba8"][~hit] == sky).all()


def test_step_count_view_ignores_lights(oracle, default_level):
    W, H = 96, 54
    a = oracle.render(default_level, gc.DIMS, gc.frame_cases(W, H)["C3i"], W, H)
    fr = gc.frame_cases(W, H)["C1"]
    fr.view_depth_field = 1
    b = oracle.render(default_level, gc.DIMS, fr, W, H)
    assert np.array_equal(a["rgba8"], b["rgba8"]) and a["counters"][2] == 0 and a["counters"][1] == 0
    grey = np.minimum(255, np.floor(np.minimum(1.0, a["steps"].astype(np.float32) / np.float32(100)) * np.float32(255) + np.float32(0.5))).astype(np.uint8)
    assert np.array_equal(a["rgba8"][..., 0], grey)
