"""CPU: the link-level seam.  voxel-rt_b200/libvxrt_glshim.so defines the GL / GLEW / GLFW
symbols the reference's objects import and forwards them to the C ABI.  Here the reference's OWN game (its six objects,
unmodified: oracle/_ref/voxel_rt_on_vxrt) runs its own main loop through the shim against a mock libvxrt that sits in the
same process, compares everything it receives with the reference's globals at every draw, and renders dumped frames
with the oracle.  The product side of the same run (real libvxrt.so) needs a GPU: test_gpu_host.py."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "voxel-rt_b200")
GAME = os.path.join(ROOT, "oracle", "_ref", "voxel_rt_on_vxrt")
MOCK_DIR = os.path.join(ROOT, "tests", "mock_vxrt")
DIMS = (512, 96, 512)

# what fshader.glsl:3-10 declares; enough for the reference's shader loader (render.cpp:104-186) where the reference
# tree itself is absent
STUB_FSHADER = ("#version 430\nconst int VOXELS_WIDTH=512;\nconst int VOXELS_HEIGHT=96;\nconst int RENDER_DIST=384;\n"
                "const int MAX_LOCAL_LIGHTS=16;\nconst int LOCAL_LIGHT_DIST=64;\nconst float AMBIENT=0.4f;\n"
                "const float DIFFUSE=0.8f;\nconst float MAX_OVERBRIGHT=1.25f;\nvoid main(){}\n")


@pytest.fixture(scope="module")
def shim(vx):
    return vx.build.build_glshim()


@pytest.fixture(scope="module")
def mock_dir(tmp_path_factory, oracle, shim):
    """a stand-in libvxrt.so (tests/mock_vxrt/mock_vxrt.c, oracle-backed) in a directory of its own"""
    d = tmp_path_factory.mktemp("mock_vxrt")
    subprocess.run(["gcc", "-O2", "-std=c11", "-Wall", "-Wextra", "-fPIC", "-shared", "-o", str(d / "libvxrt.so"),
                    os.path.join(MOCK_DIR, "mock_vxrt.c"), "-L" + ol.ORACLE_DIR, "-lvxo", "-ldl",
                    "-Wl,-rpath," + ol.ORACLE_DIR], check=True)
    subprocess.run(["gcc", "-O1", "-std=c11", "-Wall", "-o", str(d / "shim_driver"), os.path.join(MOCK_DIR, "shim_driver.c"),
                    "-L" + PKG, "-lvxrt_glshim", "-L" + str(d), "-lvxrt", "-Wl,-rpath," + PKG], check=True)
    return d


def shader_dir(tmp_path, fshader=None):
    """CWD for the game: render.cpp:326 opens "vshader.glsl" / "fshader.glsl" relatively"""
    if fshader is None and os.path.isdir(ol.REFERENCE_SRC):
        return ol.REFERENCE_SRC
    d = tmp_path / "shaders"
    d.mkdir(exist_ok=True)
    (d / "vshader.glsl").write_text("#version 430\nvoid main(){}\n")
    (d / "fshader.glsl").write_text(STUB_FSHADER if fshader is None else fshader)
    return str(d)


def run_game(cwd, env_extra, mock=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra)
    if mock is not None:
        env["LD_LIBRARY_PATH"] = str(mock) + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return subprocess.run([GAME], cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)


needs_game = pytest.mark.skipif(not os.path.exists(GAME), reason="oracle/_ref/voxel_rt_on_vxrt not built (make -C oracle ref)")


def test_shim_defines_every_gl_symbol_the_reference_imports(shim):
    out = subprocess.run(["nm", "-D", "--defined-only", shim], capture_output=True, text=True, check=True).stdout
    defined = {l.split()[-1] for l in out.splitlines() if l.strip()}
    glew = ["AttachShader", "BindBuffer", "BindBufferBase", "BindVertexArray", "BufferData", "BufferSubData", "CompileShader",
            "CreateProgram", "CreateShader", "EnableVertexAttribArray", "GenBuffers", "GenVertexArrays", "GetAttribLocation",
            "GetProgramInfoLog", "GetProgramiv", "GetShaderInfoLog", "GetShaderiv", "GetUniformLocation", "LinkProgram",
            "ShaderSource", "Uniform1f", "Uniform1i", "Uniform2f", "Uniform3f", "Uniform4fv", "UniformMatrix4fv", "UseProgram",
            "VertexAttribPointer"]                                   # the 28 variables render.o imports
    glfw = ["CreateWindow", "GetPrimaryMonitor", "Init", "MakeContextCurrent", "PollEvents", "SetCursorPosCallback",
            "SetFramebufferSizeCallback", "SetKeyCallback", "SetMouseButtonCallback", "SetScrollCallback", "SwapBuffers",
            "SwapInterval", "Terminate", "WindowShouldClose"]
    want = {"__glew" + n for n in glew} | {"glfw" + n for n in glfw} | {"glewInit", "glDrawArrays", "glShadeModel", "glViewport"}
    assert want <= defined, sorted(want - defined)
    ldd = subprocess.run(["ldd", shim], capture_output=True, text=True).stdout
    assert "libvxrt.so" in ldd and "libvxo" not in ldd and "libGL" not in ldd


def test_sub_data_batching_keeps_gl_ordering(mock_dir, tmp_path):
    log = tmp_path / "log.jsonl"
    env = dict(os.environ, LD_LIBRARY_PATH=str(mock_dir), MOCK_VXRT_LOG=str(log), VXRT_GLSHIM_FRAMES="2")
    subprocess.run([str(mock_dir / "shim_driver")], env=env, check=True, timeout=60)
    calls = [json.loads(l) for l in log.read_text().splitlines()]
    assert calls[0] == {"call": "create", "grid": [8, 4, 8], "width": 64, "height": 32, "world": 1}
    # replay (tests/mock_vxrt/shim_driver.c)
    g = np.arange(256, dtype=np.int32)
    w = 1000 + np.arange(256, dtype=np.int32)
    g[8:12] = w[8:12]; g[40:44] = w[40:44]; g[100:104] = w[100:104]
    fnv = [ol.Oracle().fnv(g)]
    g[10:14] = [7001, 7002, 7003, 7004]
    fnv.append(ol.Oracle().fnv(g))
    g[200:202] = w[200:202]
    fnv.append(ol.Oracle().fnv(g))
    rows = [c for c in calls if c["call"] == "upload_rows"]
    assert [(c["rows"], c["row_len"]) for c in rows[:3]] == [(3, 4), (1, 4), (1, 2)]
    assert [int(c["grid_fnv"], 16) for c in rows[:3]] == fnv
    g = 2 * np.arange(256, dtype=np.int32)
    g[252:256] = w[252:256]
    assert (rows[3]["rows"], rows[3]["row_len"], rows[3]["first"]) == (1, 4, 252) and len(rows) == 4
    assert int(calls[-1]["grid_fnv"], 16) == ol.Oracle().fnv(g) and calls[-1]["draws"] == 2
    assert [c["call"] for c in calls].count("upload_grid") == 2


def test_a_shader_the_kernels_do_not_implement_is_refused(mock_dir, tmp_path):
    env = dict(os.environ, LD_LIBRARY_PATH=str(mock_dir), MOCK_VXRT_LOG=str(tmp_path / "log.jsonl"))
    r = subprocess.run([str(mock_dir / "shim_driver"), "const int VOXELS_WIDTH=8;\nconst int VOXELS_HEIGHT=4;\nconst int RENDER_DIST = 512;\n"],
                       env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "RENDER_DIST" in r.stderr and "refusing" in r.stderr
    r = subprocess.run([str(mock_dir / "shim_driver"), "const int VOXELS_WIDTH=16;\nconst int VOXELS_HEIGHT=4;\n"],
                       env=env, capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "do not match the shader's grid 16x4x16" in r.stderr


@needs_game
def test_reference_game_without_a_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run_game(shader_dir(tmp_path), {"VXRT_GLSHIM_FRAMES": "2"}, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@needs_game
def test_reference_game_runs_unmodified_through_the_shim(mock_dir, tmp_path, oracle, default_level):
    """the reference's main loop (main.cpp:47-75): start-up, depth threads, 2nd upload, then a scripted session --
    look down (LMB + mouse), place a light (T), destroy (RMB), resize, walk (W)"""
    log = tmp_path / "log.jsonl"
    script = "1:mouse:400,600;1:lmb:down;9:lmb:up;10:key:T:down;11:key:T:up;12:rmb:down;13:rmb:up;15:resize:640x360;18:key:W:down"
    r = run_game(shader_dir(tmp_path), {
        "MOCK_VXRT_LOG": str(log), "VXRT_GLSHIM_READY_UPLOADS": "2", "VXRT_GLSHIM_FRAMES": "24", "VXRT_GLSHIM_FPS": "60",
        "VXRT_GLSHIM_EVENTS": script, "VXRT_GLSHIM_DUMP": str(tmp_path / "f%02d.ppm"), "VXRT_GLSHIM_DUMP_FRAMES": "14,20",
        "VXRT_GLSHIM_LOG": "1", "VXRT_GLSHIM_SAVE_GRID": str(tmp_path / "final.vxg")}, mock=mock_dir)
    assert r.returncode == 0, r.stderr
    assert "24 frames at 640x360" in r.stderr and "900 glBufferSubData calls in 1 batches" in r.stderr
    calls = [json.loads(l) for l in log.read_text().splitlines()]
    assert calls[0] == {"call": "create", "grid": [512, 96, 512], "width": 800, "height": 600, "world": 1}   # main.cpp:11-12
    assert [c["call"] for c in calls].count("upload_grid") == 2          # render.cpp:368 and :298-301
    frames = [c for c in calls if c["call"] == "render"]
    assert [f["frame"] for f in frames] == list(range(24))
    # every draw: uniforms == the reference's globals, frame size == its window, grid == its voxels[]
    assert all(f["uniform_mismatch"] == 0 and f["size_ok"] == 1 and f["grid_diff"] == 0 for f in frames), frames
    assert [f["lights_active"] for f in frames[:10]] == [0] * 10 and all(f["lights_active"] == 1 for f in frames[11:])
    assert frames[0]["rotate0"] == 1.0 and frames[14]["width"] == 800 and frames[16]["width"] == 640
    assert abs(frames[16]["aspect"] - 640 / 360) < 1e-6
    assert frames[23]["cam_pos"] != frames[17]["cam_pos"]               # walking
    # the destruction: 900 rows of 31 voxels in ONE batch, and the grid it leaves == the oracle's doDestroy
    rows = [c for c in calls if c["call"] == "upload_rows"]
    assert len(rows) == 1 and (rows[0]["rows"], rows[0]["row_len"]) == (900, 31)
    edited = default_level.copy()
    oracle.do_destroy(edited, DIMS, np.array(rows[0]["cam_pos"], np.float32), np.array(rows[0]["cam_dir"], np.float32))
    assert oracle.fnv(edited) == int(rows[0]["grid_fnv"], 16) != 0x4c58cc4001a22afa
    assert int(calls[-1]["grid_fnv"], 16) == oracle.fnv(edited)
    import voxel_rt_b200 as vx
    saved, dims = vx.gridfile.read_grid(str(tmp_path / "final.vxg"))
    assert dims == DIMS and np.array_equal(saved, edited)
    # dumped frames == the oracle's rendering of what the shim asked for
    for n, (w, h) in ((14, (800, 600)), (20, (640, 360))):
        path = tmp_path / ("f%02d.ppm" % n)
        data = path.read_bytes()
        head = b"P6\n%d %d\n255\n" % (w, h)
        assert data.startswith(head) and len(data) == len(head) + w * h * 3
        fr = ol.Frame.from_buffer_copy((tmp_path / ("f%02d.ppm.frame" % n)).read_bytes())
        want = oracle.render(edited, DIMS, fr, w, h)["rgba8"][::-1, :, :3]
        assert np.array_equal(np.frombuffer(data[len(head):], np.uint8).reshape(h, w, 3), want)
        assert want.std() > 0
