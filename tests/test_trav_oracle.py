"""CPU: the traversal grid of the CUDA path, restated on the host (oracle/vxo_trav.c), against the oracle proper (oracle/vxo.c).
The argument behind the kernels' runs -- while the band word of a -1 cell promises K further -1 cells in the travel quadrant of
its y layer (and U - 1 in the layer above), a step needs no index arithmetic, range test or load, because a step moves one cell
along one axis whatever the float state says -- is checked here on whole frames, edits, cluttered small grids next to the faces,
the known-answer rays (degenerate directions, NaNs) and the reference's tie-lock ray: castRay on the traversal grid must
return the same index, hitPos, hitNormal and iteration count, and every frame the same pixels, masks and counters."""
import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol
import test_oracle_quirks as q


def same_frame(a, b):
    for k in ("rgba8", "hit_index", "steps", "occl_mask", "cast_mask", "counters"):
        assert np.array_equal(a[k], b[k]), k


def test_band_words_follow_their_definition(oracle, default_level):
    trav, bad = oracle.trav_build(default_level, gc.DIMS)
    assert bad == 0
    assert np.array_equal(oracle.trav_canonical(trav), default_level)          # nothing but the -1 cells changes
    g = default_level.reshape(gc.DIMS[2], gc.DIMS[1], gc.DIMS[0])
    t = trav.reshape(g.shape)
    rs = np.random.RandomState(1)
    free = g == -1
    zs, ys, xs = np.nonzero(free)
    checked = 0
    for i in rs.choice(len(zs), 3000, replace=False):
        x, y, z = int(xs[i]), int(ys[i]), int(zs[i])
        w = int(t[z, y, x]) & 0xFFFFFFFF
        assert w >> 30 == 2                                                      # negative, bit 30 clear
        for qd in range(4):
            sx, sz = (1 if qd & 1 else -1), (1 if qd & 2 else -1)
            K, U = (w >> (7 * qd)) & 15, (w >> (7 * qd + 4)) & 7

            def is_free(a, b, yy):
                xx, zz = x + sx * a, z + sz * b
                return 0 <= xx < gc.DIMS[0] and 0 <= zz < gc.DIMS[2] and 0 <= yy < gc.DIMS[1] and free[zz, yy, xx]
            for n in range(1, K + 1):                                            # the promise holds ...
                assert all(is_free(a, n - a, y) for a in range(n + 1))
            if K < 15:                                                           # ... and is maximal below the cap
                assert not all(is_free(a, K + 1 - a, y) for a in range(K + 2))
            assert U <= K and U <= 7
            for n in range(U):
                assert all(is_free(a, n - a, y + 1) for a in range(n + 1))
            if U < min(K, 7):
                assert not all(is_free(a, U - a, y + 1) for a in range(U + 1))
            checked += 1
    assert checked == 12000
    # open ground: full runs in every quadrant; the layer above the band (y = 39) holds jumps, so from y = 38 nothing is promised upward
    w37, w38 = int(t[287, 37, 315]) & 0xFFFFFFFF, int(t[287, 38, 315]) & 0xFFFFFFFF          # between the trees
    assert g[287, 39, 315] == np.float32(-2.0).view(np.int32)
    assert all((w37 >> (7 * qd)) & 127 == (15 | (7 << 4)) for qd in range(4))
    assert all((w38 >> (7 * qd)) & 127 == 15 for qd in range(4))


@pytest.mark.parametrize("name", ["C1", "C2", "C3i", "C3ii_pitched", "sparse_lights", "low_sun"])
def test_frames_on_the_traversal_grid_equal_the_oracle(oracle, default_level, name):
    W, H = 480, 270
    trav, _ = oracle.trav_build(default_level, gc.DIMS)
    fr = gc.frame_cases(W, H)[name]
    a = oracle.trav_render(trav, gc.DIMS, fr, W, H)
    same_frame(a, oracle.render(default_level, gc.DIMS, fr, W, H))
    st = [int(v) for v in a["stats"]]
    assert sum(st[:6]) == int(a["counters"][3])                                  # every iteration is a run step or a checked step
    if name == "C2":
        assert st[2] > 0.55 * (st[2] + st[5])                                    # most local-light iterations need no load


def test_known_answer_rays_and_the_tie_lock(oracle, default_level, golden):
    trav, _ = oracle.trav_build(default_level, gc.DIMS)
    g = golden["ref_shader"]["kat"]
    starts, dirs, dists = gc.kat_rays(g["n"], g["seed"])
    for s, d, n in list(zip(starts, dirs, dists)) + [(q.TIE_START, q.TIE_DIR, q.TIE_DIST)]:
        a = oracle.cast_ray(default_level, gc.DIMS, s, d, n)
        b = oracle.trav_cast_ray(trav, gc.DIMS, s, d, n)
        assert a[0] == b[0] and a[3] == b[3]
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


def test_edits_random_poses_and_cluttered_small_grids(oracle, default_level):
    W, H = 256, 144
    level = default_level.copy()
    rs = np.random.RandomState(9)
    for c in [(200, 40, 180), (195, 37, 190), (3, 38, 3), (508, 36, 300), (210, 50, 200)]:   # craters, incl. next to the faces
        oracle.remove_sphere(level, gc.DIMS, c[0], c[1], c[2], 7)
    trav, bad = oracle.trav_build(level, gc.DIMS)
    assert bad == 0 and np.array_equal(oracle.trav_canonical(trav), level)
    for k in range(10):
        a, b = float(rs.uniform(-1.5, 1.5)), float(rs.uniform(-3.1, 3.1))
        ca, sa, cb, sb = np.cos(a), np.sin(a), np.cos(b), np.sin(b)
        rotx = np.array([[1, 0, 0, 0], [0, ca, sa, 0], [0, -sa, ca, 0], [0, 0, 0, 1]], np.float32)
        roty = np.array([[cb, 0, -sb, 0], [0, 1, 0, 0], [sb, 0, cb, 0], [0, 0, 0, 1]], np.float32)
        rot = (roty.T @ rotx.T).T.astype(np.float32)
        cam = (float(rs.uniform(150, 260)), float(rs.uniform(37.5, 70)), float(rs.uniform(140, 240)))
        lights = [(cam[0] + float(rs.uniform(-50, 50)), float(rs.uniform(36, 60)), cam[2] + float(rs.uniform(-50, 50)), float(rs.uniform(0.1, 1.2)))
                  for _ in range(int(rs.randint(1, 17)))]
        fr = ol.make_frame(cam, rotate=rot.ravel(), aspect=np.float32(W) / np.float32(H), lights=lights, view=int(k % 5 == 4),
                           light_pos=(float(rs.uniform(-500, 1000)), float(rs.uniform(-100, 1600)), float(rs.uniform(-500, 1000))))
        same_frame(oracle.trav_render(trav, gc.DIMS, fr, W, H), oracle.render(level, gc.DIMS, fr, W, H))
    for dims in ((40, 24, 40), (15, 15, 15), (9, 30, 70)):
        lvl = np.full(dims[0] * dims[1] * dims[2], -1, np.int32)
        gv = lvl.reshape(dims[2], dims[1], dims[0])
        gv[:, :5, :] = 0x406040
        for _ in range(12):
            x, y, z = rs.randint(1, dims[0] - 1), rs.randint(5, dims[1] - 4), rs.randint(1, dims[2] - 1)
            gv[max(0, z - 1):z + 2, 5:y, max(0, x - 1):x + 2] = int(rs.randint(0, 1 << 24))
        for with_depth in (False, True):                                         # without a depth field every empty cell is a band cell
            if with_depth:
                oracle.compute_depth_field(lvl, dims)
            trav, bad = oracle.trav_build(lvl, dims)
            assert bad == 0
            for cam in ((dims[0] / 2.0, dims[1] * 0.6, 1.5), (1.2, 6.5, dims[2] - 1.5), (dims[0] - 0.5, dims[1] - 0.5, dims[2] / 2.0)):
                fr = ol.make_frame(cam, rotate=gc.PITCHED_ROTATE, aspect=np.float32(16) / np.float32(9),
                                   light_pos=(dims[0] / 2.0, dims[1] * 5.0, dims[2] / 2.0),
                                   lights=[(2.0 + (dims[0] - 4) * i / 5.0, 6.5 + i % 3, 3.0 + (dims[2] - 6) * i / 5.0, 0.5) for i in range(6)])
                same_frame(oracle.trav_render(trav, dims, fr, 192, 108), oracle.render(lvl, dims, fr, 192, 108))


def test_values_that_cannot_be_encoded_are_reported(oracle):
    dims = (8, 8, 8)
    lvl = np.full(512, -1, np.int32)
    lvl[:64] = 0x102030
    lvl[100] = np.float32(-1.5).view(np.int32)                                   # a "jump" below 2: negative with bit 30 clear
    lvl[101] = np.int32(-2 ** 31)
    lvl[102] = np.float32(-3.0).view(np.int32)                                   # a proper jump
    trav, bad = oracle.trav_build(lvl, dims)
    assert bad == 2
