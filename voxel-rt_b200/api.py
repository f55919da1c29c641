"""Host-side mirror of the reference's render interface (src/render.hpp:30-41) over the C ABI of libvxrt.so.

`Renderer` keeps the reference's function names (updateGeometry, updatePartialGeometry, placeVoxel,
destroyVoxel, removeSphere, updateUniforms, reshape, placeLocalLight, ...) so callers and tests read like
the reference's own host code; every method is a thin ctypes call into include/vxrt.h.  There is no CPU
fallback: constructing a Renderer without an sm_100 GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

MAX_LOCAL_LIGHTS = 16
TILE_W, TILE_H = 32, 8
FLAG_DEBUG_OUTPUTS = 1


class VxrtError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("grid_w", C.c_int32), ("grid_h", C.c_int32), ("grid_d", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("world", C.c_int32), ("flags", C.c_uint32)]


class Frame(C.Structure):
    """== vxrt_frame: the shader's uniforms (fshader.glsl:20-26, render.cpp:289-296)."""
    _fields_ = [("cam_pos", C.c_float * 3), ("cam_rotation", C.c_float * 2), ("light_pos", C.c_float * 3),
                ("aspect", C.c_float), ("rotate", C.c_float * 16), ("view_depth_field", C.c_int32),
                ("lights", (C.c_float * 4) * MAX_LOCAL_LIGHTS)]


class Stats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_global", C.c_uint64), ("rays_local", C.c_uint64),
                ("fetches", C.c_uint64), ("fetches_primary", C.c_uint64), ("rays_dark", C.c_uint64), ("hit_pixels", C.c_uint64),
                ("ms_primary", C.c_float), ("ms_shadow", C.c_float), ("ms_total", C.c_float),
                ("kernel_launches", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/vxrt.h declares (checked by tests/test_cabi.py against the header text)
_SIGNATURES = {
    "vxrt_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "vxrt_destroy": (None, [C.c_void_p]),
    "vxrt_last_error": (C.c_char_p, []),
    "vxrt_device_available": (C.c_int, []),
    "vxrt_host_alloc": (C.c_void_p, [C.c_size_t]),
    "vxrt_host_free": (None, [C.c_void_p]),
    "vxrt_fnv1a64": (C.c_uint64, [C.c_void_p, C.c_size_t]),
    "vxrt_upload_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "vxrt_upload_range": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "vxrt_upload_rows": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "vxrt_update_partial": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_int32)]),
    "vxrt_download_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "vxrt_download_box": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]),
    "vxrt_save_grid": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vxrt_load_grid": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vxrt_place_voxel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int32]),
    "vxrt_destroy_voxel": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "vxrt_place_voxels": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "vxrt_set_traversal": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_traversal_active": (C.c_int, [C.c_void_p]),
    "vxrt_download_traversal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "vxrt_edit_remove_sphere": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vxrt_edit_remove_sphere_cmd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vxrt_edit_cmd_error": (C.c_int, [C.c_void_p]),
    "vxrt_build_depth_field": (C.c_int, [C.c_void_p]),
    "vxrt_generate_default_level": (C.c_int, [C.c_void_p]),
    "vxrt_generate_terrain": (C.c_int, [C.c_void_p, C.c_uint64]),
    "vxrt_terrain_height": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_int]),
    "vxrt_set_frame": (C.c_int, [C.c_void_p, C.POINTER(Frame)]),
    "vxrt_init_local_lights": (C.c_int, [C.c_void_p]),
    "vxrt_place_local_light": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]),
    "vxrt_get_frame": (C.c_int, [C.c_void_p, C.POINTER(Frame)]),
    "vxrt_resize": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vxrt_render": (C.c_int, [C.c_void_p]),
    "vxrt_sync": (C.c_int, [C.c_void_p]),
    "vxrt_set_readback_bands": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_set_l2_prefetch": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_set_culling": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_set_tile_ordering": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_set_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_set_fusion": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_frame_was_fused": (C.c_int, [C.c_void_p]),
    "vxrt_partition_tile": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vxrt_partition_owner": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vxrt_set_wide_tiles": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_set_partition": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_set_stats": (C.c_int, [C.c_void_p, C.c_int]),
    "vxrt_render_frame_host": (C.c_int, [C.c_void_p, C.POINTER(Frame), C.c_void_p]),
    "vxrt_submit_frame_host": (C.c_int, [C.c_void_p, C.POINTER(Frame), C.c_void_p]),
    "vxrt_wait_frames": (C.c_int, [C.c_void_p]),
    "vxrt_read_rgba8": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxrt_read_debug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vxrt_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "vxrt_read_block_costs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "vxrt_cast_rays": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vxrt_selftest_division": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "vxrt_selftest_reciprocal": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "vxrt_write_ppm": (C.c_int, [C.c_void_p, C.c_char_p]),
    "vxrt_local_tiles": (C.c_size_t, [C.c_void_p]),
    "vxrt_local_bytes": (C.c_size_t, [C.c_void_p]),
    "vxrt_device_rgba8": (C.c_void_p, [C.c_void_p]),
    "vxrt_stream": (C.c_void_p, [C.c_void_p]),
    "vxrt_p2p_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxrt_p2p_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxrt_p2p_attach": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxrt_p2p_base": (C.c_void_p, [C.c_void_p]),
    "vxrt_p2p_wait_frame": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "vxrt_p2p_release_frame": (C.c_int, [C.c_void_p]),
    "vxrt_p2p_readback": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxrt_p2p_error": (C.c_int, [C.c_void_p]),
    "vxrt_host_frame_create": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vxrt_host_frame_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "vxrt_host_frame_pixels": (C.c_void_p, [C.c_void_p]),
    "vxrt_render_to_host_frame": (C.c_int, [C.c_void_p, C.POINTER(Frame), C.c_void_p, C.c_uint64]),
    "vxrt_host_frame_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_int]),
    "vxrt_host_frame_release": (C.c_int, [C.c_void_p, C.c_uint64]),
    "vxrt_host_frame_destroy": (None, [C.c_void_p]),
    "vxrt_assemble_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def load_library(build_if_missing=True):
    """dlopen libvxrt.so (building it first if the sources are newer).  Fails loudly if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.lib_path()
    if build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(path):
        raise VxrtError("libvxrt.so is missing (%s): build it with `python -m voxel_rt_b200.build`; there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class HostFrame:
    """One raster RGBA8 frame in page-locked POSIX shared memory that the kernels of every rank store into
    (include/vxrt.h: vxrt_host_frame_*).  create=True on the display rank, False on the ranks that open it."""

    def __init__(self, name, width, height, create):
        self.lib = load_library()
        h = C.c_void_p()
        fn = self.lib.vxrt_host_frame_create if create else self.lib.vxrt_host_frame_open
        rc = fn(name.encode(), int(width), int(height), C.byref(h))
        if rc != 0:
            raise VxrtError("vxrt error %d: %s" % (rc, self.lib.vxrt_last_error().decode()))
        self._h = h
        self.width, self.height = int(width), int(height)

    def pixels(self):
        """[H][W][4] uint8 view of the shared frame, row 0 = bottom"""
        p = self.lib.vxrt_host_frame_pixels(self._h)
        buf = (C.c_uint8 * (self.width * self.height * 4)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(self.height, self.width, 4)

    def wait(self, world, seq, timeout_ms=4000):
        rc = self.lib.vxrt_host_frame_wait(self._h, int(world), C.c_uint64(seq), int(timeout_ms))
        if rc != 0:
            raise VxrtError("vxrt error %d: %s" % (rc, self.lib.vxrt_last_error().decode()))

    def release(self, seq):
        self.lib.vxrt_host_frame_release(self._h, C.c_uint64(seq))

    def close(self):
        if self._h:
            self.lib.vxrt_host_frame_destroy(self._h)
            self._h = None


def make_frame(cam_pos, rotate=None, light_pos=(256.0, 1536.0, 256.0), aspect=16.0 / 9.0, view=0, lights=None,
               cam_rotation=(0.0, 0.0)):
    """Frame parameters as updateUniforms() would send them (render.cpp:289-296); unused light slots are
    (-1,-1,-1,0) like initLocalLights() (render.cpp:304-311)."""
    f = Frame()
    f.cam_pos[:] = [float(np.float32(v)) for v in cam_pos]
    f.cam_rotation[:] = [float(v) for v in cam_rotation]
    f.light_pos[:] = [float(np.float32(v)) for v in light_pos]
    f.aspect = float(np.float32(aspect))
    rot = np.eye(4, dtype=np.float32).ravel() if rotate is None else np.asarray(rotate, np.float32).ravel()
    f.rotate[:] = [float(v) for v in rot]
    f.view_depth_field = int(view)
    for i in range(MAX_LOCAL_LIGHTS):
        f.lights[i][:] = [-1.0, -1.0, -1.0, 0.0]
    if lights is not None:
        for i, l in enumerate(lights):
            f.lights[i][:] = [float(np.float32(v)) for v in l]
    return f


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Renderer:
    """One context = one GPU's replica of the voxel grid plus the frame state (the reference's globals in
    main.cpp:26-37 and GL objects in render.cpp)."""

    def __init__(self, grid=(512, 96, 512), width=800, height=600, device=0, rank=0, world=1, debug=False):
        self.lib = load_library()
        self.grid = tuple(int(v) for v in grid)
        self.width, self.height = int(width), int(height)
        self.rank, self.world = int(rank), int(world)
        self.debug = bool(debug)
        cfg = Config(self.grid[0], self.grid[1], self.grid[2], self.width, self.height, int(device), self.rank, self.world,
                     FLAG_DEBUG_OUTPUTS if debug else 0)
        h = C.c_void_p()
        self._h = None
        self._pinned = []
        self._check(self.lib.vxrt_create(C.byref(cfg), C.byref(h)))
        self._h = h

    # -- plumbing --
    def _check(self, rc):
        if rc < 0:
            raise VxrtError("libvxrt error %d: %s" % (rc, self.lib.vxrt_last_error().decode()))
        return rc

    def close(self):
        if self._h is not None:
            self.lib.vxrt_destroy(self._h)
            self._h = None
            for p in self._pinned:
                self.lib.vxrt_host_free(p)
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def nvox(self):
        return self.grid[0] * self.grid[1] * self.grid[2]

    # -- render.hpp:30-41 mirror --
    def updateGeometry(self, voxels):
        """render.cpp:199-202: whole-grid upload."""
        v = np.ascontiguousarray(voxels, np.int32).ravel()
        self._check(self.lib.vxrt_upload_grid(self._h, _vp(v), v.size))

    def uploadRange(self, first, src):
        """one glBufferSubData (render.cpp:219)."""
        s = np.ascontiguousarray(src, np.int32).ravel()
        self._check(self.lib.vxrt_upload_range(self._h, int(first), s.size, _vp(s)))

    def uploadRows(self, firsts, rows):
        """a batch of equally long glBufferSubData calls (render.cpp:214-221) as one staged copy + scatter kernel;
        rows: [n][row_len] int32, firsts: [n] first voxel index of each row."""
        r = np.ascontiguousarray(rows, np.int32)
        f = np.ascontiguousarray(firsts, np.int64).ravel()
        assert r.ndim == 2 and r.shape[0] == f.size
        self._check(self.lib.vxrt_upload_rows(self._h, r.shape[0], r.shape[1], _vp(f), _vp(r)))

    def updatePartialGeometry(self, start, end, host_voxels):
        """render.cpp:204-223; returns the number of rows uploaded."""
        v = host_voxels
        assert v.dtype == np.int32 and v.flags.c_contiguous and v.size == self.nvox
        rows = C.c_int32(0)
        s = (C.c_float * 3)(*[float(x) for x in start])
        e = (C.c_float * 3)(*[float(x) for x in end])
        self._check(self.lib.vxrt_update_partial(self._h, s, e, _vp(v), C.byref(rows)))
        return rows.value

    def placeVoxel(self, x, y, z, voxel):
        self._check(self.lib.vxrt_place_voxel(self._h, x, y, z, voxel))

    def placeVoxels(self, xyz, voxels):
        """a batch of placeVoxel calls (one staged copy + one kernel); xyz [n][3] ints, voxels [n]"""
        p = np.ascontiguousarray(xyz, np.int32).reshape(-1, 3)
        v = np.ascontiguousarray(voxels, np.int32).ravel()
        assert p.shape[0] == v.size
        self._check(self.lib.vxrt_place_voxels(self._h, v.size, _vp(p), _vp(v)))

    def setTraversal(self, mode):
        """0 / False: rays read the reference-layout grid with the plain kernels; 1 / True: both passes read the traversal grid;
        2 (default): the shade pass always, the primary pass for small shares.  Same pixels."""
        self._check(self.lib.vxrt_set_traversal(self._h, int(mode)))

    def traversalActive(self):
        return bool(self.lib.vxrt_traversal_active(self._h))

    def downloadTraversal(self):
        out = np.empty(self.nvox, np.int32)
        self._check(self.lib.vxrt_download_traversal(self._h, _vp(out), out.size))
        return out

    def destroyVoxel(self, x, y, z):
        self._check(self.lib.vxrt_destroy_voxel(self._h, x, y, z))

    def removeSphere(self, pos, radius):
        """level.cpp:30-56 on the device grid."""
        self._check(self.lib.vxrt_edit_remove_sphere(self._h, int(pos[0]), int(pos[1]), int(pos[2]), int(radius)))

    def removeSphereCmd(self, device_cmd_ptr, max_radius=7):
        """removeSphere from a 16-byte command {cx, cy, cz, r} in device memory (e.g. the target of an NCCL broadcast queued on
        this context's stream); no host synchronisation"""
        self._check(self.lib.vxrt_edit_remove_sphere_cmd(self._h, C.c_void_p(int(device_cmd_ptr)), int(max_radius)))

    def editCmdError(self):
        return self._check(self.lib.vxrt_edit_cmd_error(self._h))

    def doDestroy(self, cam_pos, cam_dir, host_voxels=None):
        """controls.cpp:100-110: centre = camPos + 15*camDir, removeSphere(ivec3(centre), 7); instead of
        updatePartialGeometry the edit already happened on the device; the host mirror (if given) is synced
        from the touched box."""
        c = [np.float32(cam_pos[k]) + np.float32(15.0) * np.float32(cam_dir[k]) for k in range(3)]
        ic = [int(np.trunc(v)) for v in c]
        self.removeSphere(ic, 7)
        if host_voxels is not None:
            self.downloadBox([v - 10 for v in ic], [v + 10 for v in ic], host_voxels)
        return ic

    def buildDepthField(self):
        """computeDepthField sweep (render.cpp:273-286) on the device."""
        self._check(self.lib.vxrt_build_depth_field(self._h))

    def initVoxels(self):
        """render.cpp:349-352 + level.cpp:82-138 generated on the device (no depth field)."""
        self._check(self.lib.vxrt_generate_default_level(self._h))

    def generateTerrain(self, seed=0x5EED):
        """config C4's synthetic terrain, generated on the device (no depth field)."""
        self._check(self.lib.vxrt_generate_terrain(self._h, int(seed)))

    def terrainHeight(self, x, z, seed=0x5EED):
        return int(self.lib.vxrt_terrain_height(int(seed), int(x), int(z), self.grid[1]))

    def downloadGrid(self):
        out = np.empty(self.nvox, np.int32)
        self._check(self.lib.vxrt_download_grid(self._h, _vp(out), out.size))
        return out

    def downloadBox(self, lo, hi, host_voxels):
        assert host_voxels.dtype == np.int32 and host_voxels.flags.c_contiguous and host_voxels.size == self.nvox
        l = (C.c_int32 * 3)(*[int(v) for v in lo])
        h = (C.c_int32 * 3)(*[int(v) for v in hi])
        self._check(self.lib.vxrt_download_box(self._h, l, h, _vp(host_voxels)))

    def saveGrid(self, path):
        """device grid -> VXRTGRD1 file (gridfile.py reads it on the host)"""
        self._check(self.lib.vxrt_save_grid(self._h, os.fsencode(path)))

    def loadGrid(self, path):
        self._check(self.lib.vxrt_load_grid(self._h, os.fsencode(path)))

    def updateUniforms(self, frame):
        self._check(self.lib.vxrt_set_frame(self._h, C.byref(frame)))

    def initLocalLights(self):
        self._check(self.lib.vxrt_init_local_lights(self._h))

    def placeLocalLight(self, x, y, z, diffuse):
        return self._check(self.lib.vxrt_place_local_light(self._h, x, y, z, diffuse))

    def getFrame(self):
        f = Frame()
        self._check(self.lib.vxrt_get_frame(self._h, C.byref(f)))
        return f

    def reshape(self, width, height):
        self._check(self.lib.vxrt_resize(self._h, int(width), int(height)))
        self.width, self.height = int(width), int(height)

    def draw(self):
        """glDrawArrays(GL_TRIANGLES,0,6) main.cpp:59 (asynchronous)."""
        self._check(self.lib.vxrt_render(self._h))

    def setL2Prefetch(self, mode):
        """0 off, 1 on, 2 auto (default)"""
        self._check(self.lib.vxrt_set_l2_prefetch(self._h, int(mode)))

    def setCulling(self, enabled):
        self._check(self.lib.vxrt_set_culling(self._h, 1 if enabled else 0))

    def setFusion(self, mode):
        """0 two kernels per frame, 1 one fused kernel (a block traces its tile's primary rays, then shades its own hits), 2 auto"""
        self._check(self.lib.vxrt_set_fusion(self._h, int(mode)))

    def setPartition(self, mode):
        """0 tiles dealt in groups of `world`, rotated per tile row (default), 1 tile row r -> rank r % world (a rank's pixels are 8-row strips: one strided DMA per frame)"""
        self._check(self.lib.vxrt_set_partition(self._h, int(mode)))

    def setWideTiles(self, tiles):
        """fused frames: the `tiles` heaviest tiles get two blocks / two threads per hit pixel (0 off, default 8)"""
        self._check(self.lib.vxrt_set_wide_tiles(self._h, int(tiles)))

    def frameWasFused(self):
        """the last draw() ran as one fused kernel (setFusion)"""
        return bool(self.lib.vxrt_frame_was_fused(self._h))

    def setOverlap(self, mode):
        """0 off, 1 on, 2 auto (default): the shade pass starts inside the primary pass's tail (programmatic dependent launch)"""
        self._check(self.lib.vxrt_set_overlap(self._h, int(mode)))

    def setTileOrdering(self, enabled):
        self._check(self.lib.vxrt_set_tile_ordering(self._h, 1 if enabled else 0))

    def setStats(self, mode):
        """0 / False (default): production kernels, no per-iteration counters; 1 / True: counted variants by the reference's
        casting rule; 2: counted variants that skip what the production kernels skip (counts of executed work)"""
        self._check(self.lib.vxrt_set_stats(self._h, int(mode)))

    def setReadbackBands(self, n):
        self._check(self.lib.vxrt_set_readback_bands(self._h, int(n)))

    def sync(self):
        self._check(self.lib.vxrt_sync(self._h))

    # -- results --
    def local_bytes(self):
        return int(self.lib.vxrt_local_bytes(self._h))

    def out_shape(self):
        return (self.height, self.width, 4) if self.world == 1 else (self.local_bytes() // (TILE_W * TILE_H * 4), TILE_H, TILE_W, 4)

    def hostFrameBuffer(self, full_frame=False):
        """page-locked numpy frame buffer (vxrt_host_alloc) for renderFrameHost; freed with the renderer"""
        shape = (self.height, self.width, 4) if full_frame else self.out_shape()
        n = int(np.prod(shape))
        p = self.lib.vxrt_host_alloc(n)
        if not p:
            raise VxrtError("vxrt_host_alloc failed")
        self._pinned.append(p)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,)).reshape(shape)

    def renderFrameHost(self, frame, out=None):
        """updateUniforms + draw + read-back to host memory in one call (the end-to-end path)."""
        if out is None:
            out = np.empty(self.out_shape(), np.uint8)
        self._check(self.lib.vxrt_render_frame_host(self._h, C.byref(frame), _vp(out)))
        return out

    def submitFrameHost(self, frame, out):
        """pipelined renderFrameHost: queue the frame, read-back overlaps the next frame's kernels; `out` from hostFrameBuffer()"""
        self._check(self.lib.vxrt_submit_frame_host(self._h, C.byref(frame), _vp(out)))

    def waitFrames(self):
        self._check(self.lib.vxrt_wait_frames(self._h))

    def readPixels(self):
        out = np.empty(self.out_shape(), np.uint8)
        self._check(self.lib.vxrt_read_rgba8(self._h, _vp(out)))
        return out

    def readDebug(self):
        n = (self.height, self.width)
        hit = np.empty(n, np.int32)
        steps = np.empty(n, np.uint16)
        occl = np.empty(n, np.uint32)
        cast = np.empty(n, np.uint32)
        self._check(self.lib.vxrt_read_debug(self._h, _vp(hit), _vp(steps), _vp(occl), _vp(cast)))
        return dict(hit_index=hit, steps=steps, occl_mask=occl, cast_mask=cast)

    def stats(self):
        s = Stats()
        self._check(self.lib.vxrt_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def blockCosts(self, shade_units_per_tile=2):
        """(primary, shade): SM cycles of each local tile's primary block / each shade unit's block in the last whole-frame launch"""
        n = int(self.lib.vxrt_local_tiles(self._h))
        p = np.zeros(n, np.uint32)
        s = np.zeros(n * shade_units_per_tile, np.uint32)
        self._check(self.lib.vxrt_read_block_costs(self._h, _vp(p), _vp(s), s.size))
        return p, s

    def castRays(self, starts, dirs, dists):
        """castRay known-answer hook (fshader.glsl:59-129): returns (ret[n], out7[n,7])."""
        starts = np.ascontiguousarray(starts, np.float32)
        dirs = np.ascontiguousarray(dirs, np.float32)
        dists = np.ascontiguousarray(dists, np.int32)
        n = len(dists)
        ret = np.zeros(n, np.int32)
        out7 = np.zeros((n, 7), np.float32)
        self._check(self.lib.vxrt_cast_rays(self._h, n, _vp(starts), _vp(dirs), _vp(dists), _vp(ret), _vp(out7)))
        return ret, out7

    def selftestDivision(self, n, seed=1):
        bad = C.c_uint64(0)
        self._check(self.lib.vxrt_selftest_division(self._h, int(n), int(seed), C.byref(bad)))
        return int(bad.value)

    def selftestReciprocal(self):
        bad = C.c_uint64(0)
        self._check(self.lib.vxrt_selftest_reciprocal(self._h, C.byref(bad)))
        return int(bad.value)

    def writePPM(self, path):
        self._check(self.lib.vxrt_write_ppm(self._h, str(path).encode()))

    # -- multi-GPU plumbing --
    def device_rgba8_ptr(self):
        return int(self.lib.vxrt_device_rgba8(self._h))

    def stream_ptr(self):
        return int(self.lib.vxrt_stream(self._h) or 0)

    # peer-memory frame target (multi-GPU without a gather)
    def p2pExport(self):
        """owner rank: returns the 64-byte handle (numpy uint8) to send to the other ranks"""
        h = np.zeros(64, np.uint8)
        self._check(self.lib.vxrt_p2p_export(self._h, _vp(h)))
        return h

    def p2pImport(self, handle):
        h = np.ascontiguousarray(handle, np.uint8)
        assert h.size == 64
        self._check(self.lib.vxrt_p2p_import(self._h, _vp(h)))

    def p2pAttach(self, owner):
        """same-process variant: attach to another Renderer's peer-memory target"""
        self._check(self.lib.vxrt_p2p_attach(self._h, C.c_void_p(self.lib.vxrt_p2p_base(owner._h))))

    def p2pWaitFrame(self):
        """owner rank: stream-ordered wait for every rank's tiles; returns the device pointer of the raster frame"""
        p = C.c_void_p()
        self._check(self.lib.vxrt_p2p_wait_frame(self._h, C.byref(p)))
        return int(p.value)

    def p2pReleaseFrame(self):
        self._check(self.lib.vxrt_p2p_release_frame(self._h))

    def p2pReadback(self, out):
        """owner rank: queue acquire -> D2H into the page-locked `out` -> release; returns immediately"""
        self._check(self.lib.vxrt_p2p_readback(self._h, _vp(out)))

    def p2pError(self):
        return self._check(self.lib.vxrt_p2p_error(self._h))

    def renderToHostFrame(self, frame, host_frame, seq):
        """queue this rank's kernels; they store its tiles' pixels straight into the shared host frame"""
        self._check(self.lib.vxrt_render_to_host_frame(self._h, C.byref(frame), host_frame._h, C.c_uint64(seq)))

    def assembleTiles(self, gathered_ptr, dst_ptr, stream_ptr=0):
        self._check(self.lib.vxrt_assemble_tiles(self._h, C.c_void_p(gathered_ptr), C.c_void_p(dst_ptr), C.c_void_p(stream_ptr)))
