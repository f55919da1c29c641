"""ctypes bindings for the CPU oracle (oracle/libvxo.so) and, when built, the reference itself
(oracle/_ref/libref_host.so, oracle/_ref/libref_shader.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REFERENCE_SRC = "/root/reference/src"

MAX_LIGHTS = 16


class Dims(C.Structure):
    _fields_ = [("w", C.c_int32), ("h", C.c_int32), ("d", C.c_int32)]


class Frame(C.Structure):
    """== vxo_frame == vxrt_frame: the shader's uniforms (fshader.glsl:20-26)."""
    _fields_ = [("cam_pos", C.c_float * 3), ("cam_rotation", C.c_float * 2), ("light_pos", C.c_float * 3),
                ("aspect", C.c_float), ("rotate", C.c_float * 16), ("view_depth_field", C.c_int32),
                ("lights", (C.c_float * 4) * MAX_LIGHTS)]

    def to89(self):
        """camPos[3] camRotation[2] lightPos[3] aspect rotate[16] lights[64] as float32[89]"""
        return np.concatenate([np.array(self.cam_pos, np.float32), np.array(self.cam_rotation, np.float32),
                               np.array(self.light_pos, np.float32), np.array([self.aspect], np.float32),
                               np.array(self.rotate, np.float32),
                               np.array([list(l) for l in self.lights], np.float32).ravel()])


class RayOut(C.Structure):
    _fields_ = [("hit_pos", C.c_float * 3), ("hit_normal", C.c_float * 3), ("steps", C.c_float), ("hit_set", C.c_int32)]


def make_frame(cam_pos, rotate=None, light_pos=(256.0, 1536.0, 256.0), aspect=16.0 / 9.0, view=0, lights=None,
               cam_rotation=(0.0, 0.0)):
    f = Frame()
    f.cam_pos[:] = [float(np.float32(v)) for v in cam_pos]
    f.cam_rotation[:] = list(cam_rotation)
    f.light_pos[:] = [float(np.float32(v)) for v in light_pos]
    f.aspect = float(np.float32(aspect))
    rot = np.eye(4, dtype=np.float32).ravel() if rotate is None else np.asarray(rotate, np.float32).ravel()
    f.rotate[:] = [float(v) for v in rot]
    f.view_depth_field = int(view)
    for i in range(MAX_LIGHTS):
        f.lights[i][:] = [-1.0, -1.0, -1.0, 0.0]          # render.cpp:304-311
    if lights is not None:
        for i, l in enumerate(lights):
            f.lights[i][:] = [float(np.float32(v)) for v in l]
    return f


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def build_oracle(ref=False):
    args = ["make", "-s", "-C", ORACLE_DIR]
    subprocess.run(args, check=True)
    if ref:
        subprocess.run(args + ["ref"], check=True)


class Oracle:
    """oracle/libvxo.so"""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libvxo.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.L = C.CDLL(path)
        L.vxo_fnv1a64.restype = C.c_uint64
        L.vxo_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        L.vxo_cast_ray.restype = C.c_int32
        L.vxo_shader_index.restype = C.c_int32
        L.vxo_shader_index.argtypes = [Dims, C.c_int32, C.c_int32, C.c_int32]
        L.vxo_host_index.argtypes = [Dims, C.c_int, C.c_int, C.c_int]
        L.vxo_depth_offsets.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int]
        L.vxo_partial_ranges.argtypes = [Dims, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int32), C.c_int]
        L.vxo_remove_sphere.argtypes = [C.POINTER(C.c_int32), Dims, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vxo_fix_depth_field.argtypes = [C.POINTER(C.c_int32), Dims, C.c_int, C.c_int, C.c_int]
        L.vxo_compute_depth_field.argtypes = [C.POINTER(C.c_int32), Dims, C.c_int]
        L.vxo_do_destroy.argtypes = [C.POINTER(C.c_int32), Dims, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]

    def fnv(self, a):
        a = np.ascontiguousarray(a)
        return int(self.L.vxo_fnv1a64(a.ctypes.data, a.nbytes))

    def default_level(self, depth_field=True, nthreads=0):
        vox = np.empty(512 * 96 * 512, np.int32)
        self.L.vxo_init_default_level(_ptr(vox, C.c_int32))
        if depth_field:
            self.L.vxo_compute_depth_field(_ptr(vox, C.c_int32), Dims(512, 96, 512), nthreads)
        return vox

    def depth_offsets(self):
        d = np.zeros(2744, np.float32)
        xyz = np.zeros(3 * 2744, np.int32)
        n = self.L.vxo_depth_offsets(_ptr(d, C.c_float), _ptr(xyz, C.c_int32), 2744)
        return d[:n].copy(), xyz[:3 * n].reshape(n, 3).copy()

    def compute_depth_field(self, vox, dims, nthreads=0):
        self.L.vxo_compute_depth_field(_ptr(vox, C.c_int32), Dims(*dims), nthreads)

    def fix_depth_field(self, vox, dims, x, y, z):
        self.L.vxo_fix_depth_field(_ptr(vox, C.c_int32), Dims(*dims), x, y, z)

    def remove_sphere(self, vox, dims, cx, cy, cz, r):
        self.L.vxo_remove_sphere(_ptr(vox, C.c_int32), Dims(*dims), cx, cy, cz, r)

    def do_destroy(self, vox, dims, cam_pos, cam_dir):
        cp = (C.c_float * 3)(*cam_pos)
        cd = (C.c_float * 3)(*cam_dir)
        out = (C.c_float * 3)()
        self.L.vxo_do_destroy(_ptr(vox, C.c_int32), Dims(*dims), cp, cd, out)
        return np.array(out, np.float32)

    def partial_ranges(self, dims, start, end, max_calls=65536):
        first = np.zeros(max_calls, np.int64)
        count = np.zeros(max_calls, np.int32)
        s = (C.c_float * 3)(*start)
        e = (C.c_float * 3)(*end)
        n = self.L.vxo_partial_ranges(Dims(*dims), s, e, _ptr(first, C.c_int64), _ptr(count, C.c_int32), max_calls)
        return first[:min(n, max_calls)].copy(), count[:min(n, max_calls)].copy(), n

    def cast_ray(self, vox, dims, start, direction, dist):
        out = RayOut()
        s = (C.c_float * 3)(*[float(v) for v in start])
        d = (C.c_float * 3)(*[float(v) for v in direction])
        r = self.L.vxo_cast_ray(_ptr(vox, C.c_int32), Dims(*dims), s, d, C.c_int32(int(dist)), C.byref(out))
        return r, np.array(out.hit_pos, np.float32), np.array(out.hit_normal, np.float32), float(out.steps)

    def render(self, vox, dims, frame, width, height, y0=0, y1=None, want_f32=False, nthreads=0):
        """Returns dict(rgba8, hit_index, steps, occl_mask, cast_mask, counters[, rgba_f32]); full-frame arrays
        with row 0 = bottom; only rows [y0,y1) are filled."""
        y1 = height if y1 is None else y1
        n = width * height
        out = dict(rgba8=np.zeros((height, width, 4), np.uint8), hit_index=np.full((height, width), -2, np.int32),
                   steps=np.zeros((height, width), np.uint16), occl_mask=np.zeros((height, width), np.uint32),
                   cast_mask=np.zeros((height, width), np.uint32), counters=np.zeros(5, np.uint64))
        f32 = np.zeros((height, width, 4), np.float32) if want_f32 else None
        self.L.vxo_render(_ptr(vox, C.c_int32), Dims(*dims), C.byref(frame), C.c_int(width), C.c_int(height),
                          C.c_int(y0), C.c_int(y1), _ptr(f32, C.c_float), _ptr(out["rgba8"], C.c_uint8),
                          _ptr(out["hit_index"], C.c_int32), _ptr(out["steps"], C.c_uint16),
                          _ptr(out["occl_mask"], C.c_uint32), _ptr(out["cast_mask"], C.c_uint32),
                          _ptr(out["counters"], C.c_uint64), C.c_int(nthreads))
        if want_f32:
            out["rgba_f32"] = f32
        assert n == out["hit_index"].size
        return out

    def num_threads(self):
        return int(self.L.vxo_num_threads())

    # ---- host statement of the CUDA path's traversal grid (oracle/vxo_trav.c) ----
    def trav_build(self, vox, dims):
        """(traversal grid, number of values that cannot be encoded)"""
        trav = np.empty(vox.size, np.int32)
        self.L.vxo_trav_build.restype = C.c_int64
        bad = self.L.vxo_trav_build(_ptr(np.ascontiguousarray(vox), C.c_int32), Dims(*dims), _ptr(trav, C.c_int32))
        return trav, int(bad)

    @staticmethod
    def trav_canonical(trav):
        """reference values of traversal words: band words (negative, bit 30 clear) -> -1"""
        t = np.asarray(trav, np.int32)
        return np.where((t < 0) & ((t & 0x40000000) == 0), np.int32(-1), t)

    def trav_cast_ray(self, trav, dims, start, direction, dist):
        out = RayOut()
        s = (C.c_float * 3)(*[float(v) for v in start])
        d = (C.c_float * 3)(*[float(v) for v in direction])
        self.L.vxo_trav_cast_ray.restype = C.c_int32
        r = self.L.vxo_trav_cast_ray(_ptr(trav, C.c_int32), Dims(*dims), s, d, C.c_int32(int(dist)), C.byref(out))
        return r, np.array(out.hit_pos, np.float32), np.array(out.hit_normal, np.float32), float(out.steps)

    def trav_render(self, trav, dims, frame, width, height, y0=0, y1=None):
        """vxo_render on a traversal grid; also 'stats': fast steps / checked steps / jumps per ray kind"""
        y1 = height if y1 is None else y1
        out = dict(rgba8=np.zeros((height, width, 4), np.uint8), hit_index=np.full((height, width), -2, np.int32),
                   steps=np.zeros((height, width), np.uint16), occl_mask=np.zeros((height, width), np.uint32),
                   cast_mask=np.zeros((height, width), np.uint32), counters=np.zeros(5, np.uint64), stats=np.zeros(9, np.uint64))
        self.L.vxo_trav_render.restype = None
        self.L.vxo_trav_render(_ptr(trav, C.c_int32), Dims(*dims), C.byref(frame), C.c_int(width), C.c_int(height),
                               C.c_int(y0), C.c_int(y1), _ptr(out["rgba8"], C.c_uint8), _ptr(out["hit_index"], C.c_int32),
                               _ptr(out["steps"], C.c_uint16), _ptr(out["occl_mask"], C.c_uint32), _ptr(out["cast_mask"], C.c_uint32),
                               _ptr(out["counters"], C.c_uint64), _ptr(out["stats"], C.c_uint64))
        return out


    # ---- host statement of the production kernels' lighting (oracle/vxo_wide.c) ----
    def wide_render(self, vox, dims, frame, width, height, skip_dark, wide):
        """float RGBA frame with the lighting organised like the CUDA path's production kernels (active lights compacted, unlit rays
        skipped, the two-halves evaluation of the wide blocks)"""
        out = np.zeros((height, width, 4), np.float32)
        self.L.vxo_wide_render.restype = None
        self.L.vxo_wide_render(_ptr(vox, C.c_int32), Dims(*dims), C.byref(frame), C.c_int(width), C.c_int(height),
                               C.c_int(1 if skip_dark else 0), C.c_int(1 if wide else 0), _ptr(out, C.c_float))
        return out


def ref_available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("libref_host.so", "libref_shader.so"))


class RefHost:
    """oracle/_ref/libref_host.so: the reference's own host objects + GL shim (process-global state!)."""

    def __init__(self):
        L = self.L = C.CDLL(os.path.join(REF_DIR, "libref_host.so"))
        L.ref_host_voxels.restype = C.POINTER(C.c_int32)
        L.ref_host_init.argtypes = [C.c_char_p]
        self.n = L.ref_host_voxel_count()

    def voxels(self):
        """live numpy view of the reference's global voxels[]"""
        return np.ctypeslib.as_array(self.L.ref_host_voxels(), shape=(self.n,))

    def level_nodepth(self):
        self.L.ref_host_level_nodepth()
        return self.voxels().copy()

    def init_render(self, ref_src=REFERENCE_SRC):
        r = self.L.ref_host_init(ref_src.encode())
        assert r == 2, r
        return self.voxels().copy()

    def do_destroy(self, cam, direction, max_calls=4096):
        off = np.zeros(max_calls, np.int64)
        size = np.zeros(max_calls, np.int64)
        n = self.L.ref_host_do_destroy((C.c_float * 3)(*cam), (C.c_float * 3)(*direction),
                                       _ptr(off, C.c_longlong), _ptr(size, C.c_longlong), max_calls)
        return off[:min(n, max_calls)].copy(), size[:min(n, max_calls)].copy(), n

    def partial(self, start, end, max_calls=65536):
        off = np.zeros(max_calls, np.int64)
        size = np.zeros(max_calls, np.int64)
        n = self.L.ref_host_partial((C.c_float * 3)(*start), (C.c_float * 3)(*end),
                                    _ptr(off, C.c_longlong), _ptr(size, C.c_longlong), max_calls)
        return off[:min(n, max_calls)].copy(), size[:min(n, max_calls)].copy(), n

    def update_uniforms(self, frame):
        out = np.zeros(89, np.float32)
        v = self.L.ref_host_update_uniforms(frame.cam_pos, frame.cam_rotation, frame.light_pos, C.c_float(frame.aspect),
                                            frame.rotate, C.c_int(frame.view_depth_field), _ptr(out, C.c_float))
        return out, v

    def place_light(self, x, y, z, w):
        self.L.ref_host_place_light(C.c_float(x), C.c_float(y), C.c_float(z), C.c_float(w))

    def lights(self):
        out = np.zeros(64, np.float32)
        self.L.ref_host_get_lights(_ptr(out, C.c_float))
        return out.reshape(16, 4)

    def mouse_look(self, rx, ry):
        rot = np.zeros(16, np.float32)
        d = np.zeros(3, np.float32)
        self.L.ref_host_mouse_look_matrix(C.c_float(rx), C.c_float(ry), _ptr(rot, C.c_float), _ptr(d, C.c_float))
        return rot, d

    # ---- gameplay (controls.cpp) ----
    def player_reset(self, cam, direction, camrot, fps):
        self.L.ref_host_player_set((C.c_float * 3)(*cam), (C.c_float * 3)(*direction), (C.c_float * 2)(*camrot), C.c_longlong(fps))
        self.L.ref_host_reset_gravity()
        self.L.ref_host_init_lights()

    def player_step(self, keys9, mouse):
        """keys persist in the reference (T / SHIFT clear themselves), so the caller passes the full state each frame"""
        self.L.ref_host_player_keys((C.c_ubyte * 9)(*keys9))
        self.L.ref_host_player_mouse(*[C.c_int(int(v)) for v in mouse])
        self.L.ref_host_player_step()
        out = np.zeros(24, np.float32)
        view = self.L.ref_host_player_get(_ptr(out, C.c_float))
        return out, view

    def collided(self, cam):
        return self.L.ref_host_collided((C.c_float * 3)(*cam))

    def light_update(self, fps, rotation, n):
        out = np.zeros(4, np.float32)
        self.L.ref_host_light_update(C.c_longlong(fps), C.c_float(rotation), C.c_int(n), _ptr(out, C.c_float))
        return out


class HostLogic:
    """voxel-rt_b200/libvxrt_hostlogic.so: the product's host gameplay code (CPU only, no device calls)."""

    def __init__(self, voxels, dims=(512, 96, 512)):
        import importlib
        build = importlib.import_module("voxel_rt_b200").build
        L = self.L = C.CDLL(build.build_hostlogic())
        L.vxh_player_create.restype = C.c_void_p
        self.voxels = np.ascontiguousarray(voxels, np.int32)
        self.p = C.c_void_p(L.vxh_player_create(_ptr(self.voxels, C.c_int32), *[C.c_int(v) for v in dims]))

    def reset(self, cam, direction, camrot, fps):
        self.L.vxh_player_set(self.p, (C.c_float * 3)(*cam), (C.c_float * 3)(*direction), (C.c_float * 2)(*camrot), C.c_longlong(fps))

    def step(self, keys9, mouse):
        self.L.vxh_player_keys(self.p, (C.c_ubyte * 9)(*keys9))
        self.L.vxh_player_mouse(self.p, *[C.c_int(int(v)) for v in mouse])
        self.L.vxh_player_step(self.p)
        out = np.zeros(24, np.float32)
        view = self.L.vxh_player_get(self.p, _ptr(out, C.c_float))
        return out, view

    def take_light(self):
        out = np.zeros(3, np.float32)
        return out if self.L.vxh_player_take_light(self.p, _ptr(out, C.c_float)) else None

    def collided(self, cam):
        return self.L.vxh_player_collided(self.p, (C.c_float * 3)(*cam))

    def light_update(self, fps, rotation, n, start=(256.0, 1536.0, 256.0)):
        out = np.zeros(4, np.float32)
        self.L.vxh_light_update(C.c_longlong(fps), C.c_float(rotation), (C.c_float * 3)(*start), C.c_int(n), _ptr(out, C.c_float))
        return out

    def mouse_look(self, rx, ry):
        rot = np.zeros(16, np.float32)
        d = np.zeros(3, np.float32)
        self.L.vxh_mouse_look_matrix(C.c_float(rx), C.c_float(ry), _ptr(rot, C.c_float), _ptr(d, C.c_float))
        return rot, d

    def rotate(self, angle, axis):
        out = np.zeros(16, np.float32)
        self.L.vxh_mat4_rotate(C.c_float(angle), *[C.c_float(a) for a in axis], _ptr(out, C.c_float))
        return out

    def close(self):
        if self.p:
            self.L.vxh_player_destroy(self.p)
            self.p = None


class RefShader:
    """oracle/_ref/libref_shader.so: the reference's fshader.glsl compiled as C++ (512x96x512 only)."""

    def __init__(self):
        L = self.L = C.CDLL(os.path.join(REF_DIR, "libref_shader.so"))
        self.n = L.ref_shader_voxel_count()

    def upload(self, vox):
        assert vox.size == self.n and vox.dtype == np.int32
        self.L.ref_shader_upload(_ptr(np.ascontiguousarray(vox), C.c_int32))

    def set_frame(self, frame):
        a = frame.to89()
        self.L.ref_shader_set_uniforms(_ptr(a, C.c_float), C.c_int(frame.view_depth_field))

    def cast_rays(self, starts, dirs, dists):
        starts = np.ascontiguousarray(starts, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        dists = np.ascontiguousarray(dists, np.int32)
        n = len(dists)
        ret = np.zeros(n, np.int32)
        out7 = np.zeros((n, 7), np.float32)
        self.L.ref_shader_cast_rays(C.c_int(n), _ptr(starts, C.c_float), _ptr(dirs, C.c_float), _ptr(dists, C.c_int),
                                    _ptr(ret, C.c_int), _ptr(out7, C.c_float))
        return ret, out7

    def render(self, width, height, y0=0, y1=None, nproc=1):
        y1 = height if y1 is None else y1
        rgba = np.zeros((height, width, 4), np.float32)
        steps = np.zeros((height, width), np.float32)
        r = self.L.ref_shader_render(C.c_int(width), C.c_int(height), C.c_int(y0), C.c_int(y1),
                                     _ptr(rgba, C.c_float), _ptr(steps, C.c_float), C.c_int(nproc))
        assert r == 0, r
        return rgba, steps


def unorm8(rgba_f32):
    """the oracle's UNORM8 store rule: clamp (NaN -> 0), floor(c*255+0.5)"""
    c = np.nan_to_num(rgba_f32.astype(np.float32), nan=0.0)
    c = np.clip(c, np.float32(0), np.float32(1))
    return (c * np.float32(255.0) + np.float32(0.5)).astype(np.int32).astype(np.uint8)
