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


# ---- the reference's own game on the B200 path (link-level seam) -------------------------------
GAME = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "voxel_rt_on_vxrt")


@pytest.mark.skipif(not os.path.exists(GAME), reason="oracle/_ref/voxel_rt_on_vxrt not built (make -C oracle ref)")
def test_reference_game_on_the_b200_through_the_gl_shim(vx, oracle, default_level, tmp_path):
    """oracle/_ref/voxel_rt_on_vxrt = the reference's six objects, unmodified, linked against libvxrt_glshim.so + libvxrt.so:
    its own main loop (main.cpp:47-75) with scripted input -- look down, place a light, right-click destruction (900
    glBufferSubData calls -> one vxrt_upload_rows), window resize, walking.  Dumped frames and the final device grid
    against the oracle, bit for bit."""
    import test_glshim as tg
    vx.build.build_glshim()
    script = "1:mouse:400,600;1:lmb:down;9:lmb:up;10:key:T:down;11:key:T:up;12:rmb:down;13:rmb:up;15:resize:640x360;18:key:W:down"
    r = tg.run_game(tg.shader_dir(tmp_path), {
        "VXRT_GLSHIM_READY_UPLOADS": "2", "VXRT_GLSHIM_FRAMES": "24", "VXRT_GLSHIM_FPS": "60", "VXRT_GLSHIM_EVENTS": script,
        "VXRT_GLSHIM_DUMP": str(tmp_path / "f%02d.ppm"), "VXRT_GLSHIM_DUMP_FRAMES": "11,12,14,20", "VXRT_GLSHIM_LOG": "1",
        "VXRT_GLSHIM_SAVE_GRID": str(tmp_path / "final.vxg")}, timeout=240)
    assert r.returncode == 0, r.stderr
    assert "24 frames at 640x360" in r.stderr and "900 glBufferSubData calls in 1 batches" in r.stderr

    def dumped(n, w, h):
        data = (tmp_path / ("f%02d.ppm" % n)).read_bytes()
        head = b"P6\n%d %d\n255\n" % (w, h)
        assert data.startswith(head) and len(data) == len(head) + w * h * 3
        fr = ol.Frame.from_buffer_copy((tmp_path / ("f%02d.ppm.frame" % n)).read_bytes())
        return np.frombuffer(data[len(head):], np.uint8).reshape(h, w, 3), fr

    # frames 11 and 12 precede the edit: the reference level with its depth field, the camera at rest looking down
    img11, fr11 = dumped(11, 800, 600)
    img12, fr12 = dumped(12, 800, 600)
    assert list(fr11.cam_pos) == list(fr12.cam_pos) and list(fr11.rotate) == list(fr12.rotate)
    assert np.array_equal(img11, oracle.render(default_level, gc.DIMS, fr11, 800, 600)["rgba8"][::-1, :, :3])
    # doDestroy (controls.cpp:100-110) used camPos and camDir = rotateMatrix * (0,0,1,1) = the matrix's third column
    edited = default_level.copy()
    oracle.do_destroy(edited, gc.DIMS, np.array(fr12.cam_pos, np.float32), np.array(fr12.rotate[8:11], np.float32))
    assert not np.array_equal(edited, default_level)
    saved, dims = vx.gridfile.read_grid(str(tmp_path / "final.vxg"))
    assert dims == gc.DIMS and np.array_equal(saved, edited)
    for n, (w, h) in ((14, (800, 600)), (20, (640, 360))):
        img, fr = dumped(n, w, h)
        assert np.array_equal(img, oracle.render(edited, gc.DIMS, fr, w, h)["rgba8"][::-1, :, :3])
