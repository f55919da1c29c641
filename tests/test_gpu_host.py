"""GPU (-m gpu): the C++ host mirror of render.hpp (voxel-rt_b200/csrc/host) driven by the headless game loop
(vxrt_headless, the reference's main.cpp:47-75 without a window), frames compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def run_headless(vx, tmp_path, *args):
    exe = vx.build.build_host()
    raw = os.path.join(str(tmp_path), "frame.rgba")
    ppm = os.path.join(str(tmp_path), "frame.ppm")
    out = subprocess.run([exe, "--raw", raw, "--ppm", ppm] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    return raw, ppm, out.stdout


def test_headless_default_frame_with_lights(vx, oracle, default_level, tmp_path):
    W, H = 320, 180
    raw, ppm, log = run_headless(vx, tmp_path, "--size", W, H, "--lights")
    got = np.fromfile(raw, np.uint8).reshape(H, W, 4)
    fr = ol.make_frame(gc.CAM, aspect=np.float32(W) / np.float32(H), lights=gc.lights_4x4(gc.CAM))
    want = oracle.render(default_level, gc.DIMS, fr, W, H)["rgba8"]
    assert np.array_equal(got, want)
    # PPM: P6, upright (rows flipped), RGB only
    data = open(ppm, "rb").read()
    header = ("P6\n%d %d\n255\n" % (W, H)).encode()
    assert data.startswith(header)
    img = np.frombuffer(data[len(header):], np.uint8).reshape(H, W, 3)
    assert np.array_equal(img, want[::-1, :, :3])


def test_headless_destroy_then_frame(vx, oracle, default_level, tmp_path):
    """controls.cpp:100-110 through the C++ host: right-click looking straight down, then the next frame"""
    W, H = 256, 144
    raw, _, _ = run_headless(vx, tmp_path, "--size", W, H, "--destroy", "--view")
    got = np.fromfile(raw, np.uint8).reshape(H, W, 4)
    level = default_level.copy()
    oracle.do_destroy(level, gc.DIMS, gc.CAM, (0.0, -1.0, 0.0))
    fr = ol.make_frame(gc.CAM, aspect=np.float32(W) / np.float32(H), view=1)
    assert np.array_equal(got, oracle.render(level, gc.DIMS, fr, W, H)["rgba8"])


def test_headless_scripted_walk(vx, oracle, default_level, tmp_path):
    """main.cpp:58-66 with scripted input: movementUpdate / doMouseLook / doGravity on the host mirror, a light dropped
    half way (T), a dig on the last frame (RMB); final pose and frame against the host-logic library + oracle"""
    W, H, N = 320, 180, 90
    raw, _, log = run_headless(vx, tmp_path, "--size", W, H, "--walk", N)
    line = [l for l in log.splitlines() if l.startswith("walk")][0].split()
    pose = np.array([float.fromhex(v) for v in line[3:6] + line[7:10] + line[11:13]], np.float32)
    hl = ol.HostLogic(default_level, gc.DIMS)
    hl.reset(gc.CAM, (0.0, 0.0, 1.0), (0.0, 0.0), 60)
    lights = []
    for f in range(N):
        keys = [0] * 9
        keys[0] = keys[7] = 1
        keys[5] = int(f % 45 == 0)
        keys[4] = int(f == N // 2)
        st, view = hl.step(keys, (W // 2 + W // 20, H // 2 + H // 60, W, H))
        l = hl.take_light()
        if l is not None:
            lights.append(l)
    hl.close()
    assert np.array_equal(pose.view(np.uint32), st[:8].view(np.uint32))
    assert len(lights) == 1
    level = default_level.copy()
    oracle.do_destroy(level, gc.DIMS, st[0:3], st[3:6])
    lt = np.full((16, 4), -1.0, np.float32)
    lt[:, 3] = 0.0
    lt[0] = [lights[0][0], lights[0][1], lights[0][2], 0.5]
    fr = ol.make_frame(st[0:3], rotate=st[8:24], aspect=np.float32(W) / np.float32(H), lights=lt, cam_rotation=(float(st[6]), float(st[7])))
    got = np.fromfile(raw, np.uint8).reshape(H, W, 4)
    assert np.array_equal(got, oracle.render(level, gc.DIMS, fr, W, H)["rgba8"])


def test_headless_sun_moves_between_frames(vx, oracle, default_level, tmp_path):
    """main.cpp:58-61: lightUpdate() after every draw; the frame after N updates uses the reference's sun position"""
    W, H, N = 256, 144, 40
    raw, _, _ = run_headless(vx, tmp_path, "--size", W, H, "--frames", N)
    hl = ol.HostLogic(default_level, gc.DIMS)
    sun = hl.light_update(60, -45.0, N)
    hl.close()
    fr = ol.make_frame(gc.CAM, light_pos=sun[1:4], aspect=np.float32(W) / np.float32(H))
    got = np.fromfile(raw, np.uint8).reshape(H, W, 4)
    assert np.array_equal(got, oracle.render(default_level, gc.DIMS, fr, W, H)["rgba8"])


def test_headless_saves_the_edited_level(vx, oracle, default_level, tmp_path):
    """--destroy then --save: the file holds the reference default level after the reference's right-click edit; a second
    run that --loads it renders the same frame"""
    W, H = 160, 90
    path = os.path.join(str(tmp_path), "edited.vxg")
    raw, _, _ = run_headless(vx, tmp_path, "--size", W, H, "--destroy", "--save", path)
    first = np.fromfile(raw, np.uint8).copy()
    level = default_level.copy()
    oracle.do_destroy(level, gc.DIMS, gc.CAM, (0.0, -1.0, 0.0))
    got, dims = vx.gridfile.read_grid(path)
    assert dims == gc.DIMS and np.array_equal(got, level)
    raw, _, _ = run_headless(vx, tmp_path, "--size", W, H, "--load", path)
    assert np.array_equal(np.fromfile(raw, np.uint8), first)
