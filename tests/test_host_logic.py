"""CPU: the host gameplay code (voxel-rt_b200/csrc/host/vxrt_controls.cpp = the reference's src/controls.cpp:10-98,
112-144 restated without GLM) against golden trajectories recorded from the reference's own controls.o
(tests/golden/golden_controls.json, made by tests/golden/make_golden_controls.py) and, where oracle/_ref is built,
against the reference live, frame by frame, bit for bit."""
import json
import os

import numpy as np
import pytest

import controls_cases as cc
import golden_cases as gc
import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))


def fnv(a):
    h = 1469598103934665603        # the survey's basis, as everywhere in this repo
    for b in np.ascontiguousarray(a).view(np.uint8).tobytes():
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def place_lights(requests):
    """placeLocalLight (render.cpp:375-385) replayed over the T-key requests: first slot with a negative coordinate"""
    tab = np.zeros((16, 4), np.float32)
    tab[:, :3] = -1.0
    for p in requests:
        for i in range(16):
            if (tab[i, :3] < 0).any():
                tab[i, :3] = p
                tab[i, 3] = 0.5
                break
    return tab


@pytest.fixture(scope="module")
def logic(vx, default_level):
    h = ol.HostLogic(default_level, gc.DIMS)
    yield h
    h.close()


@pytest.fixture(scope="module")
def golden_controls():
    with open(os.path.join(HERE, "golden", "golden_controls.json")) as f:
        return json.load(f)


def _bits(vals):
    return np.array([float.fromhex(v) for v in vals], np.float32).view(np.uint32)


def test_sun_and_mouse_look_match_reference_golden(logic, golden_controls):
    """lightUpdate (render.cpp:388-402) and doMouseLook's matrix (controls.cpp:137-142)"""
    for c, want in zip(cc.SUN_CASES, golden_controls["_sun"]):
        assert np.array_equal(logic.light_update(*c).view(np.uint32), _bits(want)), c
    for c, want in zip(cc.LOOK_CASES, golden_controls["_look"]):
        assert np.array_equal(np.concatenate(logic.mouse_look(*c)).view(np.uint32), _bits(want)), c


@pytest.mark.parametrize("name", sorted(cc.cases()))
def test_trajectory_matches_reference_golden(logic, golden_controls, name):
    case = cc.cases()[name]
    g = golden_controls[name]
    states, lights = cc.run_case(logic.reset, logic.step, case, logic.take_light)
    assert len(states) == g["frames"]
    for i, want in g["samples"].items():
        want = np.array([float.fromhex(v) for v in want], np.float32)
        assert np.array_equal(states[int(i)].view(np.uint32), want.view(np.uint32)), (name, i)
    assert "%016x" % fnv(states) == g["fnv"]
    want_lights = np.array([[float.fromhex(v) for v in row] for row in g["lights"]], np.float32)
    assert np.array_equal(place_lights(lights).view(np.uint32), want_lights.view(np.uint32))


def test_player_stays_inside_the_map_and_on_the_ground(logic, default_level):
    """doGravity's clamps (controls.cpp:80-83) and the ground contact rule: after every frame the camera is inside
    [1, size-2] and the column under the player is never penetrated"""
    case = cc.cases()["corner_30"]
    states, _ = cc.run_case(logic.reset, logic.step, case)
    w, h, d = gc.DIMS
    assert (states[:, 0] >= 1).all() and (states[:, 0] <= w - 2).all()
    assert (states[:, 1] >= 11).all() and (states[:, 1] <= h - 2).all()
    assert (states[:, 2] >= 1).all() and (states[:, 2] <= d - 2).all()
    for s in states[::7]:
        assert logic.collided(s[:3]) == 0


def test_collided_counts_the_highest_solid_cell(logic, default_level):
    """controls.cpp:10-19: the return value is the LAST i in 1..9 whose cell (y - 10 + i) is solid"""
    w, h, d = gc.DIMS
    vox = np.asarray(default_level).reshape(d, h, w)
    rng = np.random.default_rng(5)
    for _ in range(300):
        x, z = rng.uniform(1, w - 2), rng.uniform(1, d - 2)
        y = rng.uniform(11, h - 2)
        want = 0
        for i in range(1, 10):
            if vox[int(z), int(y) - 10 + i, int(x)] > -1:
                want = i
        assert logic.collided((x, y, z)) == want


def test_rotate_matches_closed_form(logic):
    """glm::rotate about the principal axes: exact zeros / ones where the formula has them, cos/sin elsewhere"""
    a = np.float32(0.37)
    c, s = np.cos(a, dtype=np.float32), np.sin(a, dtype=np.float32)
    m = logic.rotate(a, (0, 1, 0)).reshape(4, 4)        # columns
    want = np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], np.float32)
    assert np.allclose(m, want, atol=1e-7)
    assert m[3, 3] == 1 and m[1, 1] == 1


ref = pytest.mark.skipif(not ol.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@ref
@pytest.mark.ref
def test_trajectories_match_reference_live(logic):
    rh = ol.RefHost()
    level = rh.level_nodepth()
    mine = ol.HostLogic(level, gc.DIMS)
    try:
        for name, case in cc.cases().items():
            want, _ = cc.run_case(rh.player_reset, rh.player_step, case)
            got, lights = cc.run_case(mine.reset, mine.step, case, mine.take_light)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
            assert np.array_equal(place_lights(lights), rh.lights()), name
    finally:
        mine.close()


@ref
@pytest.mark.ref
def test_rotate_and_collided_match_reference_live(logic):
    rh = ol.RefHost()
    level = rh.level_nodepth()
    mine = ol.HostLogic(level, gc.DIMS)
    try:
        rng = np.random.default_rng(9)
        for _ in range(200):
            rx, ry = rng.uniform(-1.5, 1.5), rng.uniform(-6, 6)
            rot, _ = rh.mouse_look(rx, 0.0)
            assert np.array_equal(mine.rotate(rx, (1, 0, 0)), rot)      # rotY(0) * rotX = rotX exactly
            cam = (rng.uniform(1, 510), rng.uniform(11, 94), rng.uniform(1, 510))
            assert mine.collided(cam) == rh.collided(cam)
            a, b = rh.mouse_look(rx, ry), mine.mouse_look(rx, ry)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            sun = (int(rng.integers(20, 240)), float(rng.uniform(-45, 361)), int(rng.integers(1, 300)))
            assert np.array_equal(rh.light_update(*sun).view(np.uint32), mine.light_update(*sun).view(np.uint32)), sun
    finally:
        mine.close()
