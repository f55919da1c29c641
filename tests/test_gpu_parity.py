"""GPU (-m gpu): the CUDA path, called through the C ABI (include/vxrt.h via voxel_rt_b200.Renderer), against the
oracle on the same inputs and against the committed golden vectors.  Bit-exact everywhere: hit index, step
count, shadow masks, ray/fetch counters AND the RGBA8 frame (the float path is IEEE-exact on both sides, so
the RGB tolerance of +-1 LSB on 99.9 % of pixels that BASELINE.json allows is not even needed)."""
import os

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol

pytestmark = pytest.mark.gpu

RGB_TOL_LSB = 0          # tolerance used below (BASELINE.json would allow 1 LSB on 0.1 % of pixels)


def h64(o, a):
    return "%016x" % o.fnv(np.ascontiguousarray(a))


def to_vx_frame(vx, fr):
    """oracle_lib.Frame and voxel_rt_b200.Frame have the same layout"""
    import ctypes as C
    out = vx.Frame()
    C.memmove(C.byref(out), C.byref(fr), C.sizeof(out))
    return out


@pytest.fixture(scope="module")
def ren(vx, default_level):
    r = vx.Renderer(grid=gc.DIMS, width=160, height=90, debug=True)
    r.updateGeometry(default_level)
    yield r
    r.close()


def check_frame(vx, oracle, ren, level, dims, fr, W, H):
    if (ren.width, ren.height) != (W, H):
        ren.reshape(W, H)
    ren.updateUniforms(to_vx_frame(vx, fr))
    ren.draw()
    rgba = ren.readPixels()
    dbg = ren.readDebug()
    st = ren.stats()
    ref = oracle.render(level, dims, fr, W, H)
    assert np.array_equal(dbg["hit_index"], ref["hit_index"])
    assert np.array_equal(dbg["steps"], ref["steps"])
    assert np.array_equal(dbg["cast_mask"], ref["cast_mask"])
    assert np.array_equal(dbg["occl_mask"], ref["occl_mask"])
    diff = np.abs(rgba.astype(np.int16) - ref["rgba8"].astype(np.int16))
    assert diff.max() <= RGB_TOL_LSB, "max RGBA8 difference %d LSB on %d pixels" % (diff.max(), int((diff.max(axis=2) > 0).sum()))
    c = ref["counters"]
    assert (st["rays_primary"], st["rays_global"], st["rays_local"], st["fetches"], st["hit_pixels"]) == tuple(int(x) for x in c)
    assert 0 <= st["rays_dark"] <= st["rays_global"] + st["rays_local"]
    return rgba, dbg, st


@pytest.mark.parametrize("name", ["C1", "C2", "C3i", "C3ii_pitched", "sparse_lights", "low_sun"])
def test_golden_frames(vx, oracle, golden, default_level, ren, name):
    W, H = golden["width"], golden["height"]
    fr = gc.frame_cases(W, H)[name]
    rgba, dbg, st = check_frame(vx, oracle, ren, default_level, gc.DIMS, fr, W, H)
    assert h64(oracle, rgba) == golden["ref_shader"]["frames"][name]["rgba8_fnv"]       # the reference shader's own frame
    assert float(st["fetches"]) == golden["ref_shader"]["frames"][name]["total_steps"]
    assert h64(oracle, dbg["hit_index"]) == golden["oracle"]["frames"][name]["hit_fnv"]
    assert h64(oracle, dbg["occl_mask"]) == golden["oracle"]["frames"][name]["occl_fnv"]


def test_cast_ray_known_answers(vx, oracle, golden, ren):
    g = golden["ref_shader"]["kat"]
    starts, dirs, dists = gc.kat_rays(g["n"], g["seed"])
    ret, out7 = ren.castRays(starts, dirs, dists)
    assert [int(x) for x in ret[:16]] == g["first16_ret"]
    assert h64(oracle, ret) == g["ret_fnv"]
    assert h64(oracle, out7) == g["out7_fnv"]


def test_tie_lock_ray(vx, oracle, default_level, ren):
    """fshader.glsl:87-104 with intersect.x == intersect.y < intersect.z: the else branch steps z for the rest of the segment
    (tests/test_oracle_quirks.py); both loop forms of ray.cuh (even index: compiler-scheduled, odd: PTX empty-cell runs)"""
    import test_oracle_quirks as q
    starts = np.array([q.TIE_START, q.TIE_START], np.float32)
    dirs = np.array([q.TIE_DIR, q.TIE_DIR], np.float32)
    ret, out7 = ren.castRays(starts, dirs, np.array([q.TIE_DIST, q.TIE_DIST], np.int32))
    r, hp, hn, st = oracle.cast_ray(default_level, gc.DIMS, q.TIE_START, q.TIE_DIR, q.TIE_DIST)
    assert r == 7391987 and [int(v) for v in ret] == [r, r]
    for k in range(2):
        assert np.array_equal(out7[k][:3].view(np.uint32), hp.view(np.uint32)) and list(out7[k][3:6]) == list(hn) and out7[k][6] == st


def test_step_reciprocals_match_ieee(vx, ren):
    """ray.cuh: 1 / |dir + 0.000001| through refined_rcp == rcp.rn on every float of [2^-40, 4)"""
    assert ren.selftestReciprocal() == 0


def test_fast_division_matches_ieee(vx, ren):
    """ray.cuh div_by (hoisted reciprocal + 3 FFMA) == __fdiv_rn on 2^31 random operand pairs of its domain"""
    assert ren.selftestDivision(1 << 31, seed=12345) == 0


@pytest.mark.parametrize("cfg", [("C1", 1280, 720), ("C2", 1920, 1080), ("C3i", 1920, 1080), ("C3ii_pitched", 1280, 720)])
def test_baseline_configs_vs_oracle(vx, oracle, default_level, ren, cfg):
    name, W, H = cfg
    check_frame(vx, oracle, ren, default_level, gc.DIMS, gc.frame_cases(W, H)[name], W, H)


@pytest.mark.parametrize("size", [(100, 37), (33, 9), (31, 7), (1, 1), (800, 600)])
def test_ragged_frame_sizes(vx, oracle, default_level, ren, size):
    W, H = size
    check_frame(vx, oracle, ren, default_level, gc.DIMS, gc.frame_cases(W, H)["C3ii_pitched"], W, H)


def test_random_poses(vx, oracle, default_level, ren):
    rs = np.random.RandomState(11)
    import oracle_lib as ol
    W, H = 192, 108
    for k in range(12):
        a, b = float(rs.uniform(-1.5, 1.5)), float(rs.uniform(-3.1, 3.1))
        ca, sa, cb, sb = np.cos(a), np.sin(a), np.cos(b), np.sin(b)
        rotx = np.array([[1, 0, 0, 0], [0, ca, sa, 0], [0, -sa, ca, 0], [0, 0, 0, 1]], np.float32)     # columns
        roty = np.array([[cb, 0, -sb, 0], [0, 1, 0, 0], [sb, 0, cb, 0], [0, 0, 0, 1]], np.float32)
        rot = (roty.T @ rotx.T).T.astype(np.float32)                                                   # column-major rotY*rotX
        cam = (float(rs.uniform(5, 505)), float(rs.uniform(38, 94)), float(rs.uniform(5, 505)))
        lights = [(cam[0] + float(rs.uniform(-50, 50)), float(rs.uniform(37, 70)), cam[2] + float(rs.uniform(-50, 50)), float(rs.uniform(0.1, 1.2)))
                  for _ in range(int(rs.randint(0, 17)))]
        fr = ol.make_frame(cam, rotate=rot.ravel(), aspect=np.float32(W) / np.float32(H), lights=lights, view=int(k % 6 == 5),
                           light_pos=(float(rs.uniform(-500, 1000)), float(rs.uniform(100, 1600)), float(rs.uniform(-500, 1000))))
        check_frame(vx, oracle, ren, default_level, gc.DIMS, fr, W, H)


def test_camera_inside_solid_and_outside_grid(vx, oracle, default_level, ren):
    import oracle_lib as ol
    W, H = 96, 54
    for cam in [(100.0, 10.0, 100.0), (-20.0, 60.0, 100.0), (256.0, 300.0, 256.0), (100.0, 37.0, 100.0), (0.5, 40.5, 0.5)]:
        fr = ol.make_frame(cam, rotate=gc.PITCHED_ROTATE, aspect=np.float32(W) / np.float32(H), lights=gc.lights_4x4(cam))
        check_frame(vx, oracle, ren, default_level, gc.DIMS, fr, W, H)


def test_production_variant_without_counters_renders_the_same_pixels(vx, default_level):
    """vxrt_set_stats(0) selects the kernel variants without the per-iteration counter; frames must not change"""
    W, H = 320, 180
    for dims_level in ("ref", "other"):
        if dims_level == "ref":
            dims, level = gc.DIMS, default_level
        else:
            dims = (64, 48, 80)
            level = np.full(dims[0] * dims[1] * dims[2], -1, np.int32)
            level.reshape(dims[2], dims[1], dims[0])[:, :20, :] = 0x804020
        with vx.Renderer(grid=dims, width=W, height=H) as r:
            r.updateGeometry(level)
            if dims_level == "other":
                r.buildDepthField()
            for name in ("C2", "C3i", "C3ii_pitched"):
                fr = to_vx_frame(vx, gc.frame_cases(W, H)[name])
                if dims_level == "other":
                    fr.cam_pos[:] = [30.0, 30.0, 10.0]
                r.setStats(True)
                a = r.renderFrameHost(fr)
                sa = r.stats()
                r.setStats(False)
                b = r.renderFrameHost(fr)
                sb = r.stats()
                assert np.array_equal(a, b), (dims_level, name)
                assert sa["hit_pixels"] == sb["hit_pixels"] and sa["rays_global"] == sb["rays_global"]
                assert sa["fetches"] > 0 and (sb["fetches"] == 0 or name == "C3i")


@pytest.mark.parametrize("size", [(416, 240), (100, 37)])
def test_banded_readback_renders_the_same_frame(vx, oracle, default_level, size):
    """vxrt_render_frame_host renders in bands of tile rows and copies each band out while the next renders:
    same frame for every band count, page-locked or pageable destination, whole-frame or tile-partition context"""
    W, H = size
    fr = gc.frame_cases(W, H)["C3ii_pitched"]
    want = oracle.render(default_level, gc.DIMS, fr, W, H)["rgba8"]
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.updateGeometry(default_level)
        r.setStats(True)                                             # counted variants: the fetch counter is compared below
        pinned = r.hostFrameBuffer()
        for nb in (1, 2, 3, 4, 7, 16):
            r.setReadbackBands(nb)
            pinned[:] = 0
            assert np.array_equal(r.renderFrameHost(to_vx_frame(vx, fr), pinned), want), nb
            assert np.array_equal(r.renderFrameHost(to_vx_frame(vx, fr)), want), nb       # pageable numpy destination
            st = r.stats()
            assert st["fetches"] == int(oracle.render(default_level, gc.DIMS, fr, W, H)["counters"][3])
        with pytest.raises(vx.VxrtError):
            r.setReadbackBands(0)
    parts = []
    for rank in range(3):
        with vx.Renderer(grid=gc.DIMS, width=W, height=H, rank=rank, world=3) as r:
            r.updateGeometry(default_level)
            r.setReadbackBands(5)
            parts.append(r.renderFrameHost(to_vx_frame(vx, fr)))
    assert np.array_equal(vx.tiles.assemble(np.stack(parts), W, H), want)


def test_pipelined_frame_submission(vx, oracle, default_level):
    """vxrt_submit_frame_host: a stream of different frames, read-back of frame k overlapping the kernels of k+1,
    two device buffers / two host buffers; every frame must arrive intact"""
    W, H = 416, 240
    names = ["C2", "C3ii_pitched", "C3i", "sparse_lights", "C1", "low_sun", "C2"]
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.updateGeometry(default_level)
        bufs = [r.hostFrameBuffer() for _ in range(len(names))]
        for k, name in enumerate(names):
            r.submitFrameHost(to_vx_frame(vx, gc.frame_cases(W, H)[name]), bufs[k])
        r.waitFrames()
        for k, name in enumerate(names):
            assert np.array_equal(bufs[k], oracle.render(default_level, gc.DIMS, gc.frame_cases(W, H)[name], W, H)["rgba8"]), (k, name)
        # 7 submits alternate between two device buffers: read_rgba8 must return the LAST frame, whichever buffer holds it
        assert np.array_equal(r.readPixels(), bufs[6])
        r.submitFrameHost(to_vx_frame(vx, gc.frame_cases(W, H)["low_sun"]), bufs[0])
        r.waitFrames()
        assert np.array_equal(r.readPixels(), bufs[5]) and np.array_equal(bufs[0], bufs[5])
        with pytest.raises(vx.VxrtError, match="page-locked"):
            r.submitFrameHost(to_vx_frame(vx, gc.frame_cases(W, H)["C1"]), np.empty((H, W, 4), np.uint8))
        # synchronous calls still work afterwards
        assert np.array_equal(r.renderFrameHost(to_vx_frame(vx, gc.frame_cases(W, H)["C1"])), bufs[4])


def test_back_to_back_frames_across_launch_order_refreshes(vx, oracle, default_level):
    """100 frames queued back to back (no synchronisation in between), alternating two poses so that a tile a bad launch order
    drops would keep the OTHER pose's pixels: the launch orders are refreshed from the block times on a side stream while the
    next frames already overwrite those times (the sorts work on a snapshot), and adopted by a later frame; every sampled
    frame must equal the oracle, and the counted rays of a frame rendered with adopted orders must equal a fresh context's"""
    W, H = 416, 240
    cases = gc.frame_cases(W, H)
    names = ["C2", "C3ii_pitched"]
    want = [oracle.render(default_level, gc.DIMS, cases[n], W, H)["rgba8"] for n in names]
    for fusion in (0, 1):
        with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
            r.updateGeometry(default_level)
            r.setFusion(fusion)
            bufs = [r.hostFrameBuffer() for _ in range(4)]
            for k in range(100):
                r.submitFrameHost(to_vx_frame(vx, cases[names[k & 1]]), bufs[k & 3])
            r.waitFrames()
            for j in range(4):                                       # frames 96..99
                assert np.array_equal(bufs[j], want[j & 1]), (fusion, j)
            r.setStats(True)
            r.updateUniforms(to_vx_frame(vx, cases["C2"])); r.draw(); st = r.stats()
            assert np.array_equal(r.readPixels(), want[0])
        with vx.Renderer(grid=gc.DIMS, width=W, height=H) as fresh:
            fresh.updateGeometry(default_level)
            fresh.setStats(True)
            fresh.updateUniforms(to_vx_frame(vx, cases["C2"])); fresh.draw(); st0 = fresh.stats()
        for key in ("rays_primary", "rays_global", "rays_local", "fetches", "hit_pixels"):
            assert st[key] == st0[key], (fusion, key, st[key], st0[key])


def test_miss_culling_never_changes_a_frame(vx, oracle, default_level):
    """production kernels end a ray as a miss once its cell is beyond every row that holds a solid voxel (ray.cuh CULL);
    frames must equal the counted (uncullled) variants and the oracle for cameras above / inside / below the solid rows,
    looking up and down, and the summary must follow uploads and voxel placement"""
    import oracle_lib as ol
    W, H = 192, 108
    rs = np.random.RandomState(21)
    level = default_level.copy()
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.updateGeometry(level)
        poses = []
        for k in range(10):
            a, b = float(rs.uniform(-1.5, 1.5)), float(rs.uniform(-3.1, 3.1))
            ca, sa, cb, sb = np.cos(a), np.sin(a), np.cos(b), np.sin(b)
            rotx = np.array([[1, 0, 0, 0], [0, ca, sa, 0], [0, -sa, ca, 0], [0, 0, 0, 1]], np.float32)
            roty = np.array([[cb, 0, -sb, 0], [0, 1, 0, 0], [sb, 0, cb, 0], [0, 0, 0, 1]], np.float32)
            rot = (roty.T @ rotx.T).T.astype(np.float32)
            cam = (float(rs.uniform(20, 490)), float(rs.choice([38.5, 45.0, 51.5, 52.5, 60.0, 94.0])), float(rs.uniform(20, 490)))
            lights = [(cam[0] + float(rs.uniform(-40, 40)), float(rs.uniform(37, 70)), cam[2] + float(rs.uniform(-40, 40)), 0.6) for _ in range(6)]
            poses.append(ol.make_frame(cam, rotate=rot.ravel(), aspect=np.float32(W) / np.float32(H), lights=lights,
                                       light_pos=(float(rs.uniform(-500, 1000)), float(rs.uniform(-200, 1600)), float(rs.uniform(-500, 1000)))))

        def check_all(lvl):
            for fr in poses:
                r.setStats(False)
                got = r.renderFrameHost(to_vx_frame(vx, fr))
                r.setStats(True)
                counted = r.renderFrameHost(to_vx_frame(vx, fr))
                assert np.array_equal(got, counted)
                assert np.array_equal(got, oracle.render(lvl, gc.DIMS, fr, W, H)["rgba8"])
        check_all(level)
        # light weights that are negative / infinite / NaN: the "term is zero" shortcut must not apply to non-finite ones
        weird = ol.make_frame(gc.CAM, aspect=np.float32(W) / np.float32(H),
                              lights=[(190.0, 40.0, 170.0, -0.5), (200.0, 45.0, 180.0, float("inf")), (185.0, 39.0, 165.0, float("nan")),
                                      (205.0, 38.0, 160.0, 1e38), (195.0, 60.0, 175.0, 0.0)])
        poses.append(weird)
        check_all(level)
        # a solid voxel far above everything else must become visible (placeVoxel extends the summary) ...
        cam = poses[0].cam_pos
        for dy in (70, 90, 95):
            x, z = int(cam[0]) + 3, int(cam[2]) + 3
            r.placeVoxel(x, dy, z, 0xff00ff)
            level[x + 512 * dy + 512 * 96 * z] = 0xff00ff
        check_all(level)
        # ... and so must solids that arrive through a partial upload
        level[300 + 512 * 80 + 512 * 96 * 300: 300 + 512 * 80 + 512 * 96 * 300 + 40] = 0x00ffff
        r.uploadRange(300 + 512 * 80 + 512 * 96 * 300, level[300 + 512 * 80 + 512 * 96 * 300: 300 + 512 * 80 + 512 * 96 * 300 + 40])
        check_all(level)
    # an all-empty grid and a grid whose only solids sit at the top
    for fill_row in (None, 15):
        dims = (32, 16, 32)
        lvl = np.full(dims[0] * dims[1] * dims[2], -1, np.int32)
        if fill_row is not None:
            lvl.reshape(dims[2], dims[1], dims[0])[:, fill_row, :] = 0x808080
        fr = ol.make_frame((16.0, 8.0, 16.0), rotate=gc.PITCHED_ROTATE, aspect=np.float32(W) / np.float32(H), light_pos=(16.0, 96.0, 16.0))
        with vx.Renderer(grid=dims, width=W, height=H) as r2:
            r2.updateGeometry(lvl)
            r2.setStats(False)
            assert np.array_equal(r2.renderFrameHost(to_vx_frame(vx, fr)), oracle.render(lvl, dims, fr, W, H)["rgba8"])


def test_render_is_idempotent_and_view_toggle(vx, ren):
    W, H = 160, 90
    ren.reshape(W, H)
    ren.setL2Prefetch(0)                                             # (its sweep kernel would add one launch per frame)
    ren.setFusion(0)                                                 # two passes: the launch counts below are theirs
    fr = to_vx_frame(vx, gc.frame_cases(W, H)["C2"])
    a = ren.renderFrameHost(fr)
    b = ren.renderFrameHost(fr)
    assert np.array_equal(a, b)
    fr.view_depth_field = 1
    c = ren.renderFrameHost(fr)
    assert np.array_equal(c[..., 0], c[..., 1]) and np.array_equal(c[..., 1], c[..., 2]) and (c[..., 3] == 255).all()
    assert ren.stats()["rays_local"] == 0 and ren.stats()["kernel_launches"] == ren.stats()["kernel_launches"] >= 1
    ren.setTileOrdering(False)
    ren.draw()
    assert ren.stats()["kernel_launches"] == 1                      # step-count view: the primary kernel only
    ren.setTileOrdering(True)
    ren.draw(); ren.draw()
    assert ren.stats()["kernel_launches"] in (1, 2)                 # + the tile-order kernel (frames of >= 64 tiles)
    assert np.array_equal(ren.readPixels(), c)                      # launch order does not change pixels
    ren.reshape(640, 360)                                            # 900 tiles: ordering active from the second frame on
    fr2 = to_vx_frame(vx, gc.frame_cases(640, 360)["C3ii_pitched"])
    ren.setTileOrdering(False)
    want = ren.renderFrameHost(fr2)
    ren.setTileOrdering(True)
    ren.updateUniforms(fr2)
    launches = []
    for _ in range(10):
        ren.draw()
        launches.append(ren.stats()["kernel_launches"])
        assert np.array_equal(ren.readPixels(), want)
    assert launches[0] == 4 and launches[1] == 4 and launches[2] == 2 and launches[8] == 4     # orders refreshed on frames 0, 1, 8, ...
    ren.setL2Prefetch(1)                                             # the cold-L2 sweep: one more launch, same pixels
    ren.draw()
    assert ren.stats()["kernel_launches"] == 3 and np.array_equal(ren.readPixels(), want)
    ren.setL2Prefetch(0)
    ren.setFusion(1)                                                 # one fused kernel per frame (+ the order refresh on frames 0, 1, 8, ...)
    launches = []
    for _ in range(10):
        ren.draw()
        launches.append(ren.stats()["kernel_launches"])
        assert np.array_equal(ren.readPixels(), want)
    assert launches[0] == 2 and launches[1] == 2 and launches[2] == 1 and launches[8] == 2
    ren.setFusion(2)
    ren.setL2Prefetch(2)


# ---- other grid shapes ---------------------------------------------------------------------------
def random_grid(oracle, dims, seed, fill=0.08):
    rs = np.random.RandomState(seed)
    w, h, d = dims
    v = np.full(w * h * d, -1, np.int32)
    solid = rs.rand(w * h * d) < fill
    v[solid] = rs.randint(0, 1 << 24, int(solid.sum()))
    g = v.reshape(d, h, w)
    g[:, : h // 4, :] = 0x336699                                   # a floor
    v = np.ascontiguousarray(g.ravel())
    oracle.compute_depth_field(v, dims)
    return v


@pytest.mark.parametrize("dims", [(64, 48, 80), (33, 17, 29), (128, 128, 128)])
def test_other_grid_shapes(vx, oracle, dims):
    import oracle_lib as ol
    level = random_grid(oracle, dims, seed=sum(dims))
    W, H = 128, 72
    with vx.Renderer(grid=dims, width=W, height=H, debug=True) as r:
        r.updateGeometry(level)
        cam = (dims[0] * 0.4, dims[1] * 0.8, dims[2] * 0.3)
        lights = [(cam[0] + 5 * i, dims[1] * 0.5, cam[2] + 7 * i, 0.6) for i in range(4)]
        for view in (0, 1):
            fr = ol.make_frame(cam, rotate=gc.PITCHED_ROTATE, aspect=np.float32(W) / np.float32(H), lights=lights, view=view,
                               light_pos=(dims[0] / 2, dims[0] * 3.0, dims[0] / 2))
            check_frame(vx, oracle, r, level, dims, fr, W, H)


# ---- grid plumbing: uploads, edits, depth field ------------------------------------------------------
def test_device_level_generator_matches_reference_level(vx, oracle, golden):
    """initVoxels() (level.cpp:82-138, incl. the origin-carving quirk) generated on the device, then the device
    depth-field sweep: both reference fingerprints"""
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.initVoxels()
        assert h64(oracle, r.downloadGrid()) == golden["ref_host"]["nodepth"]["fnv"] == "2f8d49bd81549f5a"
        r.buildDepthField()
        assert h64(oracle, r.downloadGrid()) == golden["ref_host"]["depth"]["fnv"]


def test_device_terrain_generator_matches_host_statement(vx):
    import terrain_port
    dims = (160, 128, 96)
    with vx.Renderer(grid=dims, width=32, height=8) as r:
        r.generateTerrain(seed=0x5EED)
        got = r.downloadGrid()
        assert np.array_equal(got, terrain_port.generate(dims, 0x5EED))
        assert r.terrainHeight(17, 33) == terrain_port.height(0x5EED, 17, 33, dims[1])


def test_upload_download_round_trip_and_ranges(vx, oracle, default_level):
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.updateGeometry(default_level)
        assert oracle.fnv(r.downloadGrid()) == 0x4c58cc4001a22afa
        host = default_level.copy()
        host[1000:1100] = np.arange(100, dtype=np.int32)
        r.uploadRange(1000, host[1000:1100])                       # one glBufferSubData
        assert np.array_equal(r.downloadGrid(), host)
        with pytest.raises(vx.VxrtError):
            r.uploadRange(host.size - 10, host[:100])              # GL_INVALID_VALUE in the reference
        with pytest.raises(vx.VxrtError):
            r.updateGeometry(host[:-1])


def test_upload_rows_is_a_batch_of_sub_data_calls(vx, oracle, default_level):
    """vxrt_upload_rows == the 900 equally long glBufferSubData calls of one updatePartialGeometry (render.cpp:214-221),
    incl. rows that run past x = 511 into the next row; a row outside the buffer fails the whole batch"""
    host = default_level.copy()
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.updateGeometry(host)
        edited = host.copy()
        oracle.remove_sphere(edited, gc.DIMS, 195, 40, 155, 7)
        oracle.remove_sphere(edited, gc.DIMS, 505, 40, 155, 7)
        dev = host.copy()
        first, count, n = oracle.partial_ranges(gc.DIMS, (180.0, 25.0, 140.0), (210.0, 55.0, 170.0))
        assert n == 900 and set(count.tolist()) == {31}
        wrapping = np.array([500 + 512 * (y + 96 * z) for z in range(148, 163) for y in range(33, 48)], np.int64)
        for firsts, length in ((first, 31), (wrapping, 31)):       # the second batch runs past x = 511 into the next row
            rows = np.stack([edited[f:f + length] for f in firsts])
            r.uploadRows(firsts, rows)
            for f in firsts:
                dev[f:f + length] = edited[f:f + length]
            assert np.array_equal(r.downloadGrid(), dev)
        assert not np.array_equal(dev, host)
        with pytest.raises(vx.VxrtError):
            r.uploadRows([0, host.size - 3], np.zeros((2, 8), np.int32))
        assert np.array_equal(r.downloadGrid(), dev)
        r.uploadRows(np.zeros(0, np.int64), np.zeros((0, 8), np.int32))      # empty batch: nothing happens


def test_device_depth_field_builder_matches_reference_fingerprint(vx, oracle, golden):
    nodepth = oracle.default_level(depth_field=False)
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.updateGeometry(nodepth)
        r.buildDepthField()
        out = r.downloadGrid()
    assert h64(oracle, out) == golden["ref_host"]["depth"]["fnv"] == "4c58cc4001a22afa"


def test_device_depth_field_small_grids(vx, oracle):
    for dims in [(40, 24, 36), (16, 16, 16), (70, 9, 11)]:
        rs = np.random.RandomState(sum(dims))
        n = dims[0] * dims[1] * dims[2]
        v = np.full(n, -1, np.int32)
        solid = rs.rand(n) < 0.02
        v[solid] = rs.randint(0, 1 << 24, int(solid.sum()))
        want = v.copy()
        oracle.compute_depth_field(want, dims)
        with vx.Renderer(grid=dims, width=32, height=8) as r:
            r.updateGeometry(v)
            r.buildDepthField()
            assert np.array_equal(r.downloadGrid(), want), dims


def test_device_destroy_sequence_matches_reference(vx, oracle, golden, default_level):
    """right-click destruction as a device edit (north_star item 2): same four doDestroy calls as the golden
    sequence, grid fingerprint after each; host mirror synced from the touched box only."""
    host = default_level.copy()
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.updateGeometry(host)
        for d in golden["ref_host"]["destroys"]:
            r.doDestroy(d["cam"], d["dir"], host_voxels=host)
            assert h64(oracle, r.downloadGrid()) == d["fnv"]
            assert h64(oracle, host) == d["fnv"]
        # idempotence: repeating an edit changes nothing
        d = golden["ref_host"]["destroys"][0]
        r.doDestroy(d["cam"], d["dir"])
        assert h64(oracle, r.downloadGrid()) == golden["ref_host"]["destroys"][-1]["fnv"]


def test_random_edits_and_single_voxel_ops(vx, oracle, default_level):
    rs = np.random.RandomState(2)
    host = default_level.copy()
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.updateGeometry(host)
        for k in range(12):
            c = (int(rs.randint(-5, 517)), int(rs.randint(20, 60)), int(rs.randint(-5, 517)))
            rad = int(rs.randint(0, 10))
            r.removeSphere(c, rad)
            oracle.remove_sphere(host, gc.DIMS, c[0], c[1], c[2], rad)
        r.placeVoxel(10, 50, 10, 0x123456); oracle.L.vxo_place_voxel
        host[10 + 512 * 50 + 512 * 96 * 10] = 0x123456
        r.placeVoxel(-1, 50, 10, 0x123456)                          # out of bounds: ignored (render.cpp:256-262)
        r.destroyVoxel(10, 30, 10)
        host[10 + 512 * 30 + 512 * 96 * 10] = -1
        r.destroyVoxel(512, 30, 10)
        assert np.array_equal(r.downloadGrid(), host)


def test_update_partial_geometry_semantics(vx, oracle, default_level):
    """render.cpp:204-223 incl. its quirks: rows skipped when the box starts out of bounds, start/end swap."""
    host = default_level.copy()
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.updateGeometry(host)
        dev = host.copy()                                           # what the device should hold
        edited = host.copy()
        oracle.remove_sphere(edited, gc.DIMS, 195, 40, 155, 7)
        oracle.remove_sphere(edited, gc.DIMS, 5, 40, 155, 7)
        for s, e in [((180.0, 25.0, 140.0), (210.0, 55.0, 170.0)), ((-10.0, 25.0, 140.0), (20.0, 55.0, 170.0)),
                     ((210.0, 55.0, 170.0), (180.0, 25.0, 140.0))]:
            first, count, n = oracle.partial_ranges(gc.DIMS, s, e)
            rows = r.updatePartialGeometry(s, e, edited)
            assert rows == n
            for f, c in zip(first, count):
                dev[f:f + c] = edited[f:f + c]
            assert np.array_equal(r.downloadGrid(), dev)
        assert not np.array_equal(dev, edited)                      # the edit near x=0 never reached the device (a13)


# ---- tile partition on one GPU -------------------------------------------------------------------
@pytest.mark.parametrize("rows", [False, True], ids=["tiles", "tile_rows"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_partition_reassembles_the_frame(vx, default_level, world, rows):
    """both partitions (vxrt_set_partition): tiles dealt in groups of `world` (rotated per tile row), and whole tile rows -> rank row % world (a rank's local
    buffer is then its 8-row strips); 416x236: the last tile row is cut by the frame's edge"""
    import torch
    W, H = 416, 236
    fr = to_vx_frame(vx, gc.frame_cases(W, H)["C3ii_pitched"])
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as full:
        full.updateGeometry(default_level)
        full.setStats(True)
        want = full.renderFrameHost(fr)
        st_full = full.stats()
    parts, tot = [], dict(rays_primary=0, rays_global=0, rays_local=0, fetches=0, hit_pixels=0)
    for rank in range(world):
        with vx.Renderer(grid=gc.DIMS, width=W, height=H, rank=rank, world=world) as r:
            r.updateGeometry(default_level)
            r.setPartition(1 if rows else 0)
            r.setStats(True)
            parts.append(r.renderFrameHost(fr))
            st = r.stats()
            for k in tot:
                tot[k] += st[k]
            if rank == 0:                                            # device-side un-tiling (assemble_kernel)
                keep = r
                gathered_host = None
    gathered = np.stack(parts)
    assert np.array_equal(vx.tiles.assemble(gathered, W, H, rows=rows), want)
    assert all(tot[k] == st_full[k] for k in tot)
    with vx.Renderer(grid=gc.DIMS, width=W, height=H, rank=0, world=world) as r:
        r.setPartition(1 if rows else 0)
        g = torch.from_numpy(gathered).cuda()
        dst = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        r.assembleTiles(g.data_ptr(), dst.data_ptr())
        r.sync()
        assert np.array_equal(dst.cpu().numpy(), want)


def test_peer_memory_frame_target_protocol_on_one_gpu(vx, oracle, default_level):
    """the gather-free multi-GPU path with all 'ranks' on one device: every context stores its tiles straight into
    the owner's double-buffered raster frame, completion flags are released / acquired on the streams"""
    import torch
    W, H, world = 416, 240, 3
    ctxs = [vx.Renderer(grid=gc.DIMS, width=W, height=H, rank=r, world=world) for r in range(world)]
    try:
        for r in ctxs:
            r.updateGeometry(default_level)
        ctxs[0].p2pExport()
        # an importer whose frame extents or world size differ from the owner's allocation would store out of bounds: refused
        for bad in (dict(width=W + 32, height=H, rank=1, world=world), dict(width=W, height=H, rank=1, world=world + 1)):
            with vx.Renderer(grid=(16, 16, 16), **bad) as other:
                with pytest.raises(vx.VxrtError, match="owner's frame"):
                    other.p2pAttach(ctxs[0])
        for r in ctxs[1:]:
            r.p2pAttach(ctxs[0])
        names = ["C3ii_pitched", "C2", "C3i", "sparse_lights", "C1"]          # 5 frames > 2 buffers: exercises the back-pressure
        out = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
        stream = torch.cuda.ExternalStream(ctxs[0].stream_ptr())
        for k, name in enumerate(names):
            fr = gc.frame_cases(W, H)[name]
            for r in reversed(ctxs):                                           # launch order must not matter
                r.updateUniforms(to_vx_frame(vx, fr))
                r.draw()
            ptr = ctxs[0].p2pWaitFrame()

            class _Buf:
                __cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "|u1", "data": (ptr, False), "version": 3}
            with torch.cuda.stream(stream):
                out.copy_(torch.as_tensor(_Buf(), device="cuda"))
            ctxs[0].p2pReleaseFrame()
            ctxs[0].sync()
            assert np.array_equal(out.cpu().numpy(), oracle.render(default_level, gc.DIMS, fr, W, H)["rgba8"]), name
        # pipelined owner read-back: acquire -> D2H -> release on the copy stream, next frame already rendering
        bufs = [ctxs[0].hostFrameBuffer(full_frame=True) for _ in range(2)]
        seq = ["C2", "C3ii_pitched", "C2", "C3ii_pitched", "C2", "C3ii_pitched"]
        for k, name in enumerate(seq):
            for r in ctxs:
                r.updateUniforms(to_vx_frame(vx, gc.frame_cases(W, H)[name]))
                r.draw()
            ctxs[0].p2pReadback(bufs[k & 1])
        for r in ctxs:
            r.waitFrames()
        assert np.array_equal(bufs[0], oracle.render(default_level, gc.DIMS, gc.frame_cases(W, H)["C2"], W, H)["rgba8"])
        assert np.array_equal(bufs[1], oracle.render(default_level, gc.DIMS, gc.frame_cases(W, H)["C3ii_pitched"], W, H)["rgba8"])
        assert all(r.p2pError() == 0 for r in ctxs)
        with pytest.raises(vx.VxrtError):
            ctxs[1].p2pWaitFrame()                                             # only the owner may wait
    finally:
        del out
        torch.cuda.synchronize()
        for r in ctxs:
            r.close()


@pytest.mark.parametrize("rows", [False, True], ids=["kernel_stores", "strip_dma"])
def test_host_frames_on_one_gpu(vx, oracle, default_level, rows):
    """frames straight to host memory: three 'ranks' on one device store their tiles' pixels into one raster in shared
    page-locked memory (one handle created, one opened -- the way a second process maps it); completion flags, the
    release back-pressure and two alternating host frames.  strip_dma: the tile-row partition, where a rank renders its 8-row
    strips into a local buffer and one strided DMA per frame moves them into the host frame (236 rows: the last strip is cut)"""
    W, H, world = 416, (236 if rows else 240), 3
    tag = "/vxrt_test_%d" % os.getpid()
    ctxs = [vx.Renderer(grid=gc.DIMS, width=W, height=H, rank=r, world=world) for r in range(world)]
    created = [vx.HostFrame(tag + "_a", W, H, create=True), vx.HostFrame(tag + "_b", W, H, create=True)]
    opened = [vx.HostFrame(tag + "_a", W, H, create=False), vx.HostFrame(tag + "_b", W, H, create=False)]
    try:
        with pytest.raises(vx.VxrtError):
            vx.HostFrame(tag + "_a", W, H, create=True)                       # the name exists
        with pytest.raises(vx.VxrtError, match="different size"):
            vx.HostFrame(tag + "_a", W + 32, H, create=False)
        for r in ctxs:
            r.updateGeometry(default_level)
            r.setPartition(1 if rows else 0)
            r.setStats(False)
        names = ["C3ii_pitched", "C2", "C3i", "sparse_lights", "C1", "C2"]
        cases = gc.frame_cases(W, H)
        want = {n: oracle.render(default_level, gc.DIMS, cases[n], W, H)["rgba8"] for n in set(names)}
        for k, name in enumerate(names):                                      # frame k goes to host frame k & 1, its seq = k // 2 + 1
            seq = k // 2 + 1
            for r in reversed(ctxs):                                          # launch order must not matter
                hf = (created if r.rank == 0 else opened)[k & 1]               # rank 0 through its own mapping, the others through the opened one
                r.renderToHostFrame(to_vx_frame(vx, cases[name]), hf, seq)
            if k >= 1:                                                        # consume frame k - 1 while frame k renders
                j = k - 1
                created[j & 1].wait(world, j // 2 + 1)
                assert np.array_equal(created[j & 1].pixels(), want[names[j]]), names[j]
                created[j & 1].release(j // 2 + 1)
        j = len(names) - 1
        created[j & 1].wait(world, j // 2 + 1)
        assert np.array_equal(opened[j & 1].pixels(), want[names[j]])         # both mappings show the same memory
        # back-pressure: frame 4 of host frame b is not released yet -> rendering frame 5 into it must refuse (bounded wait)
        # (not exercised here: the wait is 4 s long); a context with other extents is refused at once
        with vx.Renderer(grid=(16, 16, 16), width=64, height=32) as other:
            with pytest.raises(vx.VxrtError, match="other extents"):
                other.renderToHostFrame(to_vx_frame(vx, cases["C1"]), created[0], 9)
        with pytest.raises(vx.VxrtError, match="did not deliver"):
            created[0].wait(world, 99, timeout_ms=50)
    finally:
        for r in ctxs:
            r.sync(); r.close()
        for h in opened + created:
            h.close()
    assert not os.path.exists("/dev/shm" + tag + "_a") and not os.path.exists("/dev/shm" + tag + "_b")


# ---- grid files ------------------------------------------------------------------------------------
def test_grid_file_round_trip_through_the_device(vx, oracle, default_level, tmp_path):
    """vxrt_save_grid / vxrt_load_grid against the host-side reader / writer of the same format (gridfile.py)"""
    W, H = 192, 108
    p1, p2 = str(tmp_path / "dev.vxg"), str(tmp_path / "host.vxg")
    fr = ol.make_frame(gc.CAM, aspect=np.float32(W) / np.float32(H), lights=gc.lights_4x4(gc.CAM))
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.initVoxels(); r.buildDepthField()                 # device-generated level
        r.removeSphere((195, 40, 165), 7)
        r.saveGrid(p1)
        edited = r.downloadGrid()
    hd = vx.gridfile.read_header(p1)
    assert hd["dims"] == gc.DIMS and hd["fnv"] == oracle.fnv(edited)
    got, _ = vx.gridfile.read_grid(p1)
    assert np.array_equal(got, edited)
    want_level = default_level.copy()
    oracle.remove_sphere(want_level, gc.DIMS, 195, 40, 165, 7)
    assert np.array_equal(got, want_level)
    vx.gridfile.write_grid(p2, default_level, gc.DIMS)     # host-written file -> device
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.loadGrid(p2)
        assert np.array_equal(r.downloadGrid(), default_level)
        r.updateUniforms(to_vx_frame(vx, fr)); r.draw()
        assert np.array_equal(r.readPixels(), oracle.render(default_level, gc.DIMS, fr, W, H)["rgba8"])
        r.loadGrid(p1)                                       # and the device-written one (culling summary must follow)
        r.draw()
        assert np.array_equal(r.readPixels(), oracle.render(want_level, gc.DIMS, fr, W, H)["rgba8"])


def test_grid_file_errors(vx, tmp_path):
    rng = np.random.default_rng(4)
    dims = (16, 8, 12)
    v = rng.integers(-1, 100, size=16 * 8 * 12, dtype=np.int32)
    good = str(tmp_path / "good.vxg")
    vx.gridfile.write_grid(good, v, dims)
    raw = bytearray(open(good, "rb").read())
    with vx.Renderer(grid=dims, width=32, height=8) as r:
        with pytest.raises(vx.VxrtError, match="before any grid upload"):
            r.saveGrid(str(tmp_path / "none.vxg"))
        with pytest.raises(vx.VxrtError, match="cannot open"):
            r.loadGrid(str(tmp_path / "missing.vxg"))
        r.loadGrid(good)
        assert np.array_equal(r.downloadGrid(), v)
        flipped = bytearray(raw); flipped[64 + 5] ^= 1
        open(str(tmp_path / "flip.vxg"), "wb").write(bytes(flipped))
        with pytest.raises(vx.VxrtError, match="fingerprint"):
            r.loadGrid(str(tmp_path / "flip.vxg"))
        with pytest.raises(vx.VxrtError, match="before any grid upload"):      # a failed load leaves no grid in use
            r.draw()
        open(str(tmp_path / "short.vxg"), "wb").write(bytes(raw[:-4]))
        with pytest.raises(vx.VxrtError, match="shorter"):
            r.loadGrid(str(tmp_path / "short.vxg"))
        open(str(tmp_path / "magic.vxg"), "wb").write(b"NOTAGRID" + bytes(raw[8:]))
        with pytest.raises(vx.VxrtError, match="VXRTGRD1"):
            r.loadGrid(str(tmp_path / "magic.vxg"))
        r.loadGrid(good); r.draw()
    with vx.Renderer(grid=(16, 8, 13), width=32, height=8) as r:
        with pytest.raises(vx.VxrtError, match="extents"):
            r.loadGrid(good)


# ---- API state / error behaviour -----------------------------------------------------------------
def test_errors_are_loud(vx):
    with vx.Renderer(grid=(16, 16, 16), width=32, height=8) as r:
        with pytest.raises(vx.VxrtError, match="before any grid upload"):
            r.draw()
        with pytest.raises(vx.VxrtError):
            r.readPixels()
    with pytest.raises(vx.VxrtError):
        vx.Renderer(grid=(2048, 2048, 2048))
    with pytest.raises(vx.VxrtError):
        vx.Renderer(grid=(16, 16, 16), rank=2, world=2)


def test_local_light_slots(vx):
    with vx.Renderer(grid=(16, 16, 16), width=32, height=8) as r:
        f = r.getFrame()
        assert [list(l) for l in f.lights] == [[-1.0, -1.0, -1.0, 0.0]] * 16      # render.cpp:304-311
        for i in range(16):
            assert r.placeLocalLight(1.0 + i, 2.0, 3.0, 0.5) == i                  # first free slot, render.cpp:375-385
        assert r.placeLocalLight(9.0, 9.0, 9.0, 0.5) == 16                         # 17th is silently dropped
        f = r.getFrame()
        assert list(f.lights[15]) == [16.0, 2.0, 3.0, 0.5]
        r.reshape(64, 16)
        assert r.getFrame().aspect == 4.0                                          # reshape, render.cpp:410


# ---- the traversal grid (csrc/trav.cuh): the device copy the rays read ------------------------------------------------------
def check_traversal_grid(oracle, r, dims, what):
    """the device's traversal words == a host rebuild (oracle/vxo_trav.c) from the device's own reference-layout grid"""
    level = r.downloadGrid()
    want, bad = oracle.trav_build(level, dims)
    got = r.downloadTraversal()
    nbad = int((got != want).sum())
    assert nbad == 0, "%s: %d traversal words differ from the host rebuild" % (what, nbad)
    assert r.traversalActive() == (bad == 0), what
    return level


def test_traversal_grid_is_kept_coherent_by_every_entry_point(vx, oracle, golden, default_level, tmp_path):
    dims = gc.DIMS
    with vx.Renderer(grid=dims, width=32, height=8) as r:
        r.updateGeometry(default_level)                                          # full upload
        check_traversal_grid(oracle, r, dims, "upload")
        host = default_level.copy()
        for dd in golden["ref_host"]["destroys"]:                                # the reference's destroy sequence (device edits)
            r.doDestroy(dd["cam"], dd["dir"])
        lvl = check_traversal_grid(oracle, r, dims, "destroy sequence")
        assert h64(oracle, lvl) == golden["ref_host"]["destroys"][-1]["fnv"]
        rs = np.random.RandomState(17)
        for _ in range(6):                                                       # craters incl. next to / across the faces
            c = (int(rs.randint(-3, 515)), int(rs.randint(28, 50)), int(rs.randint(-3, 515)))
            r.removeSphere(c, int(rs.randint(1, 9)))
        check_traversal_grid(oracle, r, dims, "random craters")
        r.placeVoxel(200, 37, 200, 0x123456); r.placeVoxel(201, 38, 200, 0x123456); r.destroyVoxel(100, 36, 100)
        r.placeVoxel(0, 37, 0, 0x10); r.placeVoxel(511, 38, 511, 0x10)
        check_traversal_grid(oracle, r, dims, "single voxels")
        pts = np.stack([rs.randint(150, 260, 400), rs.randint(36, 42, 400), rs.randint(150, 260, 400)], 1)
        vals = rs.randint(0, 1 << 24, 400).astype(np.int32)
        vals[::7] = -1
        pts[5] = pts[3]; pts[9] = (-4, 40, 10)                                   # a repeated cell (last wins), a cell outside the grid
        r.placeVoxels(pts, vals)
        lvl2 = check_traversal_grid(oracle, r, dims, "voxel batch")
        want = lvl.copy()

        def idx(x, y, z):
            return x + 512 * y + 512 * 96 * z
        for (x, y, z), v in zip(pts, vals):                                      # the batch replayed in order on the host (its cells only)
            if 0 <= x < 512 and 0 <= y < 96 and 0 <= z < 512:
                want[idx(x, y, z)] = v
        cells = [idx(x, y, z) for x, y, z in pts if 0 <= x < 512 and 0 <= y < 96 and 0 <= z < 512]
        assert np.array_equal(lvl2[cells], want[cells])
        # partial uploads of host data: updatePartialGeometry, a range, a batch of rows
        host = lvl2.copy()
        hv = host.reshape(512, 96, 512)
        hv[300:310, 37:45, 300:312] = 0x777777
        assert r.updatePartialGeometry((299.0, 36.0, 299.0), (313.0, 46.0, 311.0), host) > 0
        check_traversal_grid(oracle, r, dims, "updatePartialGeometry")
        first = 40 + 512 * 38 + 512 * 96 * 60
        r.uploadRange(first, np.full(3000, 0x00ff00, np.int32))                  # crosses rows of one z slab
        r.uploadRange(512 * 96 * 70 - 100, np.full(300, -1, np.int32))           # crosses a slab boundary
        check_traversal_grid(oracle, r, dims, "uploadRange")
        firsts = np.array([10 + 512 * 37 + 512 * 96 * z for z in range(400, 420)], np.int64)
        r.uploadRows(firsts, np.full((20, 50), 0x0000ff, np.int32))
        check_traversal_grid(oracle, r, dims, "uploadRows")
        path = str(tmp_path / "g.vxg")
        r.saveGrid(path)
        r.initVoxels()                                                           # device generator, no depth field: every empty cell is a band cell
        check_traversal_grid(oracle, r, dims, "initVoxels")
        r.buildDepthField()
        assert np.array_equal(check_traversal_grid(oracle, r, dims, "buildDepthField"), default_level)
        r.loadGrid(path)
        check_traversal_grid(oracle, r, dims, "loadGrid")


def test_traversal_on_and_off_render_the_same_frames(vx, oracle, default_level):
    """vxrt_set_traversal: rays on the traversal grid vs the plain kernels on the reference-layout grid -- frames, debug planes
    and counters are those of the oracle both ways; so are castRay's known answers incl. the tie-lock ray"""
    import test_oracle_quirks as q
    W, H = 640, 360
    level = default_level.copy()
    with vx.Renderer(grid=gc.DIMS, width=W, height=H, debug=True) as r:
        r.updateGeometry(level)
        for c in [(200, 40, 180), (195, 37, 190), (210, 50, 200)]:
            r.removeSphere(c, 7)
            oracle.remove_sphere(level, gc.DIMS, c[0], c[1], c[2], 7)
        for on in (True, False):
            r.setTraversal(on)
            assert r.traversalActive() == on
            for name in ("C2", "C3ii_pitched", "low_sun", "C3i"):
                check_frame(vx, oracle, r, level, gc.DIMS, gc.frame_cases(W, H)[name], W, H)
            ret, out7 = r.castRays(np.array([q.TIE_START] * 2, np.float32), np.array([q.TIE_DIR] * 2, np.float32), np.array([q.TIE_DIST] * 2, np.int32))
            rr, hp, hn, st = oracle.cast_ray(level, gc.DIMS, q.TIE_START, q.TIE_DIR, q.TIE_DIST)
            assert [int(v) for v in ret] == [rr, rr] and out7[0][6] == st and out7[1][6] == st
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:                    # production kernels (no counters, culling)
        r.updateGeometry(level)
        fr = gc.frame_cases(W, H)["C2"]
        want = oracle.render(level, gc.DIMS, fr, W, H)["rgba8"]
        for on in (True, False, True):
            r.setTraversal(on)
            assert np.array_equal(r.renderFrameHost(to_vx_frame(vx, fr)), want), on


def test_values_the_traversal_grid_cannot_encode_fall_back_to_the_plain_kernels(vx, oracle):
    """a negative value with bit 30 clear (a "jump" shorter than 2: the reference never produces one) cannot be told from a band
    word: the context then renders from the reference-layout grid -- same pixels as the oracle -- and says so"""
    dims = (48, 32, 48)
    lvl = np.full(dims[0] * dims[1] * dims[2], -1, np.int32)
    gv = lvl.reshape(dims[2], dims[1], dims[0])
    gv[:, :6, :] = 0x406040
    oracle.compute_depth_field(lvl, dims)
    fr = ol.make_frame((24.0, 12.0, 3.0), aspect=np.float32(16) / np.float32(9), light_pos=(24.0, 150.0, 24.0),
                       lights=[(8.0 + 6 * i, 8.0, 10.0 + 5 * i, 0.5) for i in range(5)])
    W, H = 192, 108
    with vx.Renderer(grid=dims, width=W, height=H, debug=True) as r:
        r.updateGeometry(lvl)
        assert r.traversalActive()
        check_frame(vx, oracle, r, lvl, dims, fr, W, H)
        odd = lvl.copy()
        odd.reshape(gv.shape)[10:20, 7, 10:30] = np.float32(-1.25).view(np.int32)     # jumps of 1.25
        r.updateGeometry(odd)
        assert not r.traversalActive()
        check_frame(vx, oracle, r, odd, dims, fr, W, H)
        r.updateGeometry(lvl)                                                    # a full upload starts the count over
        assert r.traversalActive()
        r.placeVoxel(5, 8, 5, int(np.float32(-1.5).view(np.int32)))
        assert not r.traversalActive()


# ---- the published configurations at the sizes bench.py publishes (BASELINE.json configs[2..4]) -----------------------
def to_ol_frame(fr):
    import ctypes as C
    out = ol.Frame()
    C.memmove(C.byref(out), C.byref(fr), C.sizeof(out))
    return out


@pytest.mark.parametrize("name", ["C2", "C3i", "C3ii_pitched"])
def test_published_4k_frames_on_the_production_kernels(vx, oracle, default_level, name):
    """BASELINE configs[2] at 3840x2160 ("C2" lights at 4K == workload C3ii_4k, the headline): a non-debug context with the
    counters off runs exactly the kernels bench.py times (culling, unlit rays skipped, launch ordering from the second frame
    on); every pixel of the 4K frame against the oracle, for the first frame and for a frame rendered with launch orders"""
    W, H = 3840, 2160
    fr = gc.frame_cases(W, H)[name]
    want = oracle.render(default_level, gc.DIMS, fr, W, H)["rgba8"]
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.updateGeometry(default_level)
        out = r.hostFrameBuffer()
        for k in range(3):                                           # frame 0: no launch order yet; 1, 2: slowest-first orders
            out[:] = 0x5A                                            # poison: a dropped tile cannot hide behind the last frame
            r.renderFrameHost(to_vx_frame(vx, fr), out)
            bad = int((out != want).any(axis=2).sum())
            assert bad == 0, "%s frame %d: %d of %d pixels differ from the oracle" % (name, k, bad, W * H)
        r.updateUniforms(to_vx_frame(vx, fr))
        for k in range(2):                                           # vxrt_render (whole-frame launch, ordered) + read-back
            r.draw()
        assert np.array_equal(r.readPixels(), want)


def test_c5_thousand_edits_match_the_reference_replay(vx, oracle, default_level):
    """BASELINE configs[4]: 1000 right-click edits (controls.cpp:100-110 -> level.cpp:30-56) on the device grid, a 4K
    production frame every 100 edits; grid fingerprint and frames against the oracle's replay of the same edits"""
    W, H = 3840, 2160
    edits = vx.scenes.edit_centres(1000)
    fr = gc.frame_cases(W, H)["C3ii_pitched"]
    level = default_level.copy()
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.updateGeometry(level)
        out = r.hostFrameBuffer()
        for k, c in enumerate(edits):
            r.removeSphere(c, 7)
            oracle.remove_sphere(level, gc.DIMS, int(c[0]), int(c[1]), int(c[2]), 7)
            if k % 100 == 99:
                out[:] = 0x5A
                r.renderFrameHost(to_vx_frame(vx, fr), out)
                want = oracle.render(level, gc.DIMS, fr, W, H)["rgba8"]
                bad = int((out != want).any(axis=2).sum())
                assert bad == 0, "after %d edits: %d pixels differ" % (k + 1, bad)
        got = r.downloadGrid()
    assert h64(oracle, got) == h64(oracle, level)


def test_c4_terrain_1024_depth_field_and_frame_rows(vx, oracle):
    """BASELINE configs[3]: the 1024^3 synthetic terrain generated and depth-fielded on the device (4 GiB).  The depth
    field against the oracle's fixDepthField (render.cpp:226-253) on >= 10^5 sampled cells incl. every face of the grid
    (only the SIGN of the neighbours enters, so the downloaded grid serves as input), the height field against the host
    statement of the generator, then 96 rows of the 4K production frame against the oracle on the downloaded grid"""
    import terrain_port
    dims = vx.scenes.TERRAIN_GRID
    W, H = 3840, 2160
    w, h, d = dims
    with vx.Renderer(grid=dims, width=W, height=H) as r:
        r.generateTerrain(vx.scenes.TERRAIN_SEED)
        r.buildDepthField()
        level = r.downloadGrid()
        surf = r.terrainHeight(w // 2, d // 2)
        vfr = vx.scenes.terrain_frame(W, H, surf)
        out = r.hostFrameBuffer()
        out[:] = 0x5A
        r.renderFrameHost(vfr, out)
        got = np.array(out, copy=True)
    g = level.reshape(d, h, w)
    # generator: columns against the host statement (surface height, material bands)
    rs = np.random.RandomState(4)
    for _ in range(200):
        x, z = int(rs.randint(0, w)), int(rs.randint(0, d))
        s = terrain_port.height(vx.scenes.TERRAIN_SEED, x, z, h)
        assert (g[z, :s + 1, x] >= 0).all() and g[z, h - 1, x] < 0       # solid up to the surface (trees only add solids above it)
    # depth field: sampled cells, all six faces included
    n = 120000
    xs, ys, zs = rs.randint(0, w, n), rs.randint(0, h, n), rs.randint(0, d, n)
    k = n // 12
    xs[:k] = 0; xs[k:2 * k] = w - 1; ys[2 * k:3 * k] = 0; ys[3 * k:4 * k] = h - 1; zs[4 * k:5 * k] = 0; zs[5 * k:6 * k] = d - 1
    near = slice(6 * k, 9 * k)                                        # cells close above the surface (where the values vary)
    for i in range(near.start, near.stop):
        ys[i] = min(h - 1, terrain_port.height(vx.scenes.TERRAIN_SEED, int(xs[i]), int(zs[i]), h) + 1 + int(rs.randint(0, 9)))
    work = np.where(level < 0, np.int32(-1), level)                   # the level before the depth field (render.cpp:349-352: -1)
    for x, y, z in zip(xs, ys, zs):
        oracle.fix_depth_field(work, dims, int(x), int(y), int(z))
    idx = xs.astype(np.int64) + w * ys.astype(np.int64) + w * h * zs.astype(np.int64)
    bad = int((work[idx] != level[idx]).sum())
    assert bad == 0, "%d of %d sampled depth-field cells differ from the oracle" % (bad, n)
    del work
    # frame: 12 blocks of 8 rows spread over the frame (sky, horizon, ground)
    fr = to_ol_frame(vfr)
    for b in range(12):
        y0 = (b * 22 + 3) * 8
        ref = oracle.render(level, dims, fr, W, H, y0=y0, y1=y0 + 8)["rgba8"]
        assert np.array_equal(got[y0:y0 + 8], ref[y0:y0 + 8]), "rows %d..%d" % (y0, y0 + 8)


def test_stats_modes_count_the_reference_rule_and_the_executed_work(vx, oracle, default_level):
    """vxrt_set_stats: 1 = every ray the reference casts, marched to its end (== the oracle's counters); 2 = what the production
    kernels execute (unlit rays not traced, certain misses cut short); 0 = no counters.  Same pixels in every mode."""
    W, H = 640, 360
    for name in ("C2", "C3ii_pitched", "low_sun"):
        fr = gc.frame_cases(W, H)[name]
        ref = oracle.render(default_level, gc.DIMS, fr, W, H)
        with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
            r.updateGeometry(default_level)
            frames, st = {}, {}
            for mode in (0, 1, 2):
                r.setStats(mode)
                frames[mode] = r.renderFrameHost(to_vx_frame(vx, fr))
                st[mode] = r.stats()
                assert np.array_equal(frames[mode], ref["rgba8"]), (name, mode)
            c = [int(x) for x in ref["counters"]]
            assert (st[1]["rays_primary"], st[1]["rays_global"], st[1]["rays_local"], st[1]["fetches"], st[1]["hit_pixels"]) == tuple(c)
            assert st[0]["fetches"] == 0 and st[0]["rays_local"] == 0 and st[0]["hit_pixels"] == c[4]
            # mode 2: traced + skipped == the reference's rays; never more iterations than the reference
            assert st[2]["rays_global"] + st[2]["rays_local"] + st[2]["rays_dark"] == st[1]["rays_global"] + st[1]["rays_local"]
            assert st[2]["rays_dark"] == st[1]["rays_dark"] and 0 < st[2]["fetches"] < st[1]["fetches"]
            assert st[2]["hit_pixels"] == c[4] and st[2]["rays_primary"] == c[0]
            with pytest.raises(vx.VxrtError):
                r.setStats(3)


def test_fused_and_overlapped_frames_render_the_same_pixels(vx, oracle, default_level):
    """vxrt_set_fusion: one kernel per frame, every block traces its tile's primary rays and then shades its own hits;
    vxrt_set_overlap: two kernels, the shade kernel launched with programmatic stream serialization, its blocks wait for their
    tile's ready flag.  Same pixels as the oracle for production and counted kernels, across
    frames that differ (a stale flag or hit slot would show), frame sizes, launch orders and a tile-partition context"""
    level = default_level
    for W, H in ((640, 360), (1920, 1080), (100, 37)):
        names = ["C2", "C3ii_pitched", "low_sun", "C3i", "sparse_lights", "C2"]
        want = {n: oracle.render(level, gc.DIMS, gc.frame_cases(W, H)[n], W, H)["rgba8"] for n in set(names)}
        with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
            r.updateGeometry(level)
            out = r.hostFrameBuffer()
            for fusion, mode in ((0, 1), (0, 0), (0, 2), (1, 0), (2, 0)):       # two passes (overlapped / not / auto), fused, auto
                r.setFusion(fusion)
                r.setOverlap(mode)
                for stats in (0, 1):
                    r.setStats(stats)
                    for k, n in enumerate(names * 2):
                        r.updateUniforms(to_vx_frame(vx, gc.frame_cases(W, H)[n]))
                        r.draw()                                             # whole-frame launch: the overlapped path
                        got = r.readPixels()
                        bad = int((got != want[n]).any(axis=2).sum())
                        assert bad == 0, (W, H, fusion, mode, stats, k, n, bad)
            r.setStats(0)
            r.setFusion(0)
            r.setOverlap(1)
            out[:] = 0x5A
            r.renderFrameHost(to_vx_frame(vx, gc.frame_cases(W, H)["C2"]), out)   # banded path: not overlapped, same pixels
            assert np.array_equal(out, want["C2"])
    W, H, world = 1280, 720, 4
    fr = gc.frame_cases(W, H)["C3ii_pitched"]
    want = oracle.render(level, gc.DIMS, fr, W, H)["rgba8"]
    parts = []
    for rank in range(world):
        with vx.Renderer(grid=gc.DIMS, width=W, height=H, rank=rank, world=world) as r:
            r.updateGeometry(level)
            r.setFusion(rank % 2)                                                # ranks may even differ in how they schedule their tiles
            r.setOverlap(1)
            r.updateUniforms(to_vx_frame(vx, fr))
            for _ in range(3):
                r.draw()
            parts.append(r.readPixels())
    assert np.array_equal(vx.tiles.assemble(np.stack(parts), W, H), want)


def test_wide_tiles_render_the_same_pixels(vx, oracle, default_level):
    """vxrt_set_wide_tiles: the heaviest tiles of a fused frame get two blocks and two threads per hit pixel -- one walks the
    global shadow ray and the first half of the active lights, the other the second half without the early-out, combined in slot
    order afterwards.  Same pixels as the oracle with 0 / 8 / 64 wide tiles: light sets whose overbright clamp ends the loop in
    the first half, in the second half or never; gaps between the slots; negative / infinite / NaN weights; an odd number of
    lights; culling off (unlit rays traced)"""
    import oracle_lib as ol
    W, H = 640, 360
    aspect = np.float32(W) / np.float32(H)
    cam = gc.CAM
    near = [(cam[0] + dx, 40.0 + (k % 3), cam[2] + 20.0 + dz, w) for k, (dx, dz, w) in enumerate(
        [(-12, 0, 0.5), (-8, 6, 0.5), (-4, 12, 0.5), (0, 18, 0.5), (4, 24, 0.5), (8, 30, 0.5), (12, 36, 0.5), (-10, 40, 0.5),
         (-6, 34, 0.5), (-2, 28, 0.5), (2, 22, 0.5), (6, 16, 0.5), (10, 10, 0.5), (14, 4, 0.5), (-14, 8, 0.5), (0, 44, 0.5)])]
    gaps = [(-1, -1, -1, 0)] * 16
    for slot, l in zip((1, 2, 6, 7, 8, 13, 15), near):
        gaps[slot] = l
    frames = {
        "strong_16": ol.make_frame(cam, aspect=aspect, lights=near),                                    # clamp early (first half)
        "weak_16": ol.make_frame(cam, aspect=aspect, lights=[(x, y, z, 0.07) for x, y, z, _ in near]),   # clamp late or never
        "mixed_16": ol.make_frame(cam, aspect=aspect, lights=[(x, y, z, 0.02 if k < 8 else 0.6) for k, (x, y, z, _) in enumerate(near)]),
        "gaps_7": ol.make_frame(cam, aspect=aspect, lights=gaps),
        "odd_5": ol.make_frame(cam, aspect=aspect, lights=near[:5]),
        "weird": ol.make_frame(cam, aspect=aspect, lights=[(190.0, 40.0, 170.0, -0.5), (200.0, 45.0, 180.0, float("inf")), (185.0, 39.0, 165.0, float("nan")),
                                                            (205.0, 38.0, 160.0, 1e38), (195.0, 60.0, 175.0, 0.0), (192.0, 41.0, 172.0, 0.3)]),
        "pitched": gc.frame_cases(W, H)["C3ii_pitched"],
        "one_light": ol.make_frame(cam, aspect=aspect, lights=near[:1]),
        "no_lights": gc.frame_cases(W, H)["C1"],
    }
    want = {n: oracle.render(default_level, gc.DIMS, fr, W, H)["rgba8"] for n, fr in frames.items()}
    with vx.Renderer(grid=gc.DIMS, width=W, height=H) as r:
        r.updateGeometry(default_level)
        r.setFusion(1)
        r.setStats(False)
        for culling in (True, False):
            r.setCulling(culling)
            for wide in (64, 8, 0):
                r.setWideTiles(wide)
                for rep in range(2):
                    for n, fr in frames.items():
                        r.updateUniforms(to_vx_frame(vx, fr))
                        r.draw(); r.draw(); r.draw()                   # (the launch order -- which tiles are wide -- settles over frames)
                        got = r.readPixels()
                        bad = int((got != want[n]).any(axis=2).sum())
                        assert bad == 0, (culling, wide, rep, n, bad)
        with pytest.raises(vx.VxrtError):
            r.setWideTiles(65)


def test_remove_sphere_from_a_device_side_command(vx, oracle, default_level):
    """vxrt_edit_remove_sphere_cmd: the 16-byte command {cx, cy, cz, r} lives in device memory (in the multi-GPU split it is the
    target of an NCCL broadcast on the render stream); the grid and its traversal grid must end up as after the host-argument
    edit / the oracle's removeSphere; a radius beyond max_radius is refused and reported"""
    import torch
    level = default_level.copy()
    rs = np.random.RandomState(23)
    with vx.Renderer(grid=gc.DIMS, width=32, height=8) as r:
        r.updateGeometry(level)
        stream = torch.cuda.ExternalStream(r.stream_ptr())
        cmd = torch.zeros(4, dtype=torch.int32, device="cuda")
        for k in range(8):
            c = [int(rs.randint(-2, 514)), int(rs.randint(28, 48)), int(rs.randint(-2, 514)), int(rs.randint(0, 8))]
            with torch.cuda.stream(stream):
                cmd.copy_(torch.tensor(c, dtype=torch.int32), non_blocking=False)
            r.removeSphereCmd(cmd.data_ptr(), 7)
            r.sync()
            oracle.remove_sphere(level, gc.DIMS, *c)
        got = check_traversal_grid(oracle, r, gc.DIMS, "device-side commands")
        assert h64(oracle, got) == h64(oracle, level)
        assert r.editCmdError() == 0
        with torch.cuda.stream(stream):
            cmd.copy_(torch.tensor([100, 40, 100, 9], dtype=torch.int32))
        r.removeSphereCmd(cmd.data_ptr(), 7)                                     # radius 9 > max_radius 7: refused, nothing changes
        assert r.editCmdError() == 1
        assert h64(oracle, r.downloadGrid()) == h64(oracle, level)
        del cmd
        torch.cuda.synchronize()
