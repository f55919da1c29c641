"""CPU: the oracle (oracle/libvxo.so) against the committed golden vectors (tests/golden/golden.json), which
were produced by the reference itself (oracle/_ref builds, see tests/golden/make_golden.py)."""
import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol


def h64(o, a):
    return "%016x" % o.fnv(np.ascontiguousarray(a))


def test_depth_offset_table(oracle, golden):
    dist, xyz = oracle.depth_offsets()
    g = golden["ref_host"]["depth_offsets"]
    assert len(dist) == g["count"] == 1419
    assert h64(oracle, dist) == g["dist_fnv"]
    assert h64(oracle, xyz) == g["xyz_fnv"]
    # render.cpp:77-81: pre-shrunk distance, all within radius 7
    assert dist.max() == 0.0 and dist.min() >= -7.0


def test_default_level_without_depth_field(oracle, golden):
    v = oracle.default_level(depth_field=False)
    g = golden["ref_host"]["nodepth"]
    assert h64(oracle, v) == g["fnv"] == "2f8d49bd81549f5a"
    assert int((v >= 0).sum()) == g["solid"] and int((v == -1).sum()) == g["empty"]


def test_default_level_with_depth_field(oracle, golden, default_level):
    v = default_level
    g = golden["ref_host"]
    assert h64(oracle, v) == g["depth"]["fnv"] == "4c58cc4001a22afa"
    assert int((v >= 0).sum()) == g["depth"]["solid"]
    assert int((v == -1).sum()) == g["depth"]["empty"]
    assert int((v < -1).sum()) == g["depth"]["field"]
    vals, cnt = np.unique(v[v < -1], return_counts=True)
    assert {"%08x" % (int(a) & 0xffffffff): int(c) for a, c in zip(vals, cnt)} == g["depth_values"]
    for key, val in g["spot"].items():
        x, y, z = (int(t) for t in key.split(","))
        assert int(v[x + 512 * y + 512 * 96 * z]) == val


def test_destroy_sequence_and_upload_pattern(oracle, golden, default_level):
    """controls.cpp:100-110 -> level.cpp:30-56 -> render.cpp:204-223, four right-clicks incl. two next to the low
    faces where the reference uploads nothing."""
    v = default_level.copy()
    for d in golden["ref_host"]["destroys"]:
        centre = oracle.do_destroy(v, gc.DIMS, d["cam"], d["dir"])
        assert h64(oracle, v) == d["fnv"]
        assert int((v >= 0).sum()) == d["solid"] and int((v == -1).sum()) == d["empty"]
        first, count, n = oracle.partial_ranges(gc.DIMS, centre - np.float32(15), centre + np.float32(15))
        assert n == d["calls"]
        assert int(count.astype(np.int64).sum()) * 4 == d["bytes"]
        if n:
            assert int(first[0]) * 4 == d["first_offset"]
            assert h64(oracle, first * 4) == d["offsets_fnv"]
            assert h64(oracle, count.astype(np.int64) * 4) == d["sizes_fnv"]


def test_cast_ray_known_answers(oracle, golden, default_level):
    g = golden["ref_shader"]["kat"]
    starts, dirs, dists = gc.kat_rays(g["n"], g["seed"])
    ret = np.zeros(g["n"], np.int32)
    out7 = np.zeros((g["n"], 7), np.float32)
    for i in range(g["n"]):
        r, hp, hn, st = oracle.cast_ray(default_level, gc.DIMS, starts[i], dirs[i], dists[i])
        ret[i] = r
        out7[i, :3] = hp; out7[i, 3:6] = hn; out7[i, 6] = st
    assert [int(x) for x in ret[:16]] == g["first16_ret"]
    assert [float(x) for x in out7[:16, 6]] == g["first16_steps"]
    assert int((ret >= 0).sum()) == g["hits"]
    assert h64(oracle, ret) == g["ret_fnv"]
    assert h64(oracle, out7) == g["out7_fnv"]


@pytest.mark.parametrize("name", ["C1", "C2", "C3i", "C3ii_pitched", "sparse_lights", "low_sun"])
def test_frames(oracle, golden, default_level, name):
    W, H = golden["width"], golden["height"]
    fr = gc.frame_cases(W, H)[name]
    out = oracle.render(default_level, gc.DIMS, fr, W, H, want_f32=True)
    g = golden["ref_shader"]["frames"][name]
    assert h64(oracle, out["rgba_f32"]) == g["fcolor_fnv"]
    assert h64(oracle, out["rgba8"]) == g["rgba8_fnv"]
    assert float(out["counters"][3]) == g["total_steps"]
    go = golden["oracle"]["frames"][name]
    assert h64(oracle, out["hit_index"]) == go["hit_fnv"]
    assert h64(oracle, out["occl_mask"]) == go["occl_fnv"]
    assert h64(oracle, out["cast_mask"]) == go["cast_fnv"]
    assert [int(x) for x in out["counters"]] == go["counters"]


def test_row_band_rendering_matches_full_frame(oracle, default_level):
    W, H = 96, 54
    fr = gc.frame_cases(W, H)["C2"]
    full = oracle.render(default_level, gc.DIMS, fr, W, H)
    a = oracle.render(default_level, gc.DIMS, fr, W, H, y0=0, y1=20)
    b = oracle.render(default_level, gc.DIMS, fr, W, H, y0=20, y1=H)
    assert np.array_equal(full["rgba8"][:20], a["rgba8"][:20]) and np.array_equal(full["rgba8"][20:], b["rgba8"][20:])
    assert list(full["counters"]) == list(a["counters"] + b["counters"])


def test_shader_index_matches_host_index_in_bounds_and_wraps(oracle):
    L = oracle.L
    d = ol.Dims(*gc.DIMS)
    rs = np.random.RandomState(3)
    for _ in range(2000):
        x, y, z = int(rs.randint(-3, 515)), int(rs.randint(-3, 99)), int(rs.randint(-3, 515))
        assert L.vxo_shader_index(d, x, y, z) == L.vxo_host_index(d, x, y, z)
    # fshader.glsl:33-52 multiplies before it range-checks: y = INT_MIN wraps to 0 (y*512 mod 2^32 == 0)
    assert L.vxo_shader_index(d, 5, -2 ** 31, 7) == 5 + 512 * 96 * 7
    assert L.vxo_shader_index(d, 5, 2 ** 31 - 1, 7) == -1
