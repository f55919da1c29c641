"""The benchmark / parity configurations of SURVEY.md section 8(d) as concrete inputs: camera poses, light
patterns and edit sequences.  Pure host-side description (numpy only); grids come from the caller (the
reference default level) or from the synthetic terrain generator below."""
import numpy as np

from .api import make_frame

DEFAULT_GRID = (512, 96, 512)                    # render.hpp:4-5
DEFAULT_CAM = (195.0, 55.0, 155.0)               # main.cpp:26
DEFAULT_LIGHT = (256.0, 1536.0, 256.0)           # main.cpp:33-34 (before the first lightUpdate)
PITCHED_CAM = (195.0, 60.0, 155.0)

# rotY(0.6) * rotX(0.5) exactly as controls.cpp:138-141 builds it with the reference's GLM (float hex, column-major);
# tests/test_oracle_vs_ref.py re-derives it from the reference build.
PITCHED_ROTATE = [float.fromhex(h) for h in (
    "0x1.a692640000000p-1", "0x0.0p+0", "-0x1.2118d20000000p-1", "0x0.0p+0",
    "0x1.1533700000000p-2", "0x1.c152800000000p-1", "0x1.952ef80000000p-2", "0x0.0p+0",
    "0x1.fb69b20000000p-2", "-0x1.eaee880000000p-2", "0x1.72d7780000000p-1", "0x0.0p+0",
    "0x0.0p+0", "0x0.0p+0", "0x0.0p+0", "0x1.0000000000000p+0")]

RESOLUTIONS = {"720p": (1280, 720), "1080p": (1920, 1080), "4k": (3840, 2160)}


def lights_4x4(cam, y=40.0, weight=0.5):
    """16 lights on a 24-unit 4x4 grid in front of the camera; weight 0.5 is what
    controls.cpp:47 passes to placeLocalLight."""
    return [(cam[0] - 36 + 24 * (i % 4), y, cam[2] + 10 + 24 * (i // 4), weight) for i in range(16)]


def frame_for(name, width, height, grid=DEFAULT_GRID):
    """name in: 'C1' (no local lights), 'C2' (16 lights), 'C3i' (step-count view), 'C3ii' (= C2 lights),
    'C3ii_pitched' (second pose), 'C3i_pitched'."""
    aspect = np.float32(width) / np.float32(height)      # reshape() render.cpp:410
    cam = DEFAULT_CAM
    light = (grid[0] / 2.0, grid[0] * 3.0, grid[0] / 2.0)
    if name == "C1":
        return make_frame(cam, light_pos=light, aspect=aspect)
    if name in ("C2", "C3ii"):
        return make_frame(cam, light_pos=light, aspect=aspect, lights=lights_4x4(cam))
    if name == "C3i":
        return make_frame(cam, light_pos=light, aspect=aspect, lights=lights_4x4(cam), view=1)
    if name == "C3ii_pitched":
        return make_frame(PITCHED_CAM, rotate=PITCHED_ROTATE, light_pos=light, aspect=aspect, lights=lights_4x4(PITCHED_CAM),
                          cam_rotation=(0.5, 0.6))
    if name == "C3i_pitched":
        return make_frame(PITCHED_CAM, rotate=PITCHED_ROTATE, light_pos=light, aspect=aspect, lights=lights_4x4(PITCHED_CAM),
                          cam_rotation=(0.5, 0.6), view=1)
    raise KeyError(name)


TERRAIN_GRID = (1024, 1024, 1024)                # config C4: 4 GiB of int32 voxels, replicated per GPU
TERRAIN_SEED = 0x5EED


def terrain_frame(width, height, surface_at_cam, grid=TERRAIN_GRID, view=0):
    """config C4: camera above the middle of the synthetic terrain, pitched pose, sun at
    (w/2, 3w, w/2), 16 lights on the 4x4 pattern 4 voxels above the surface height under the camera."""
    aspect = np.float32(width) / np.float32(height)
    cam = (grid[0] / 2.0, float(surface_at_cam + 20), grid[2] / 2.0)
    light = (grid[0] / 2.0, grid[0] * 3.0, grid[2] / 2.0)
    return make_frame(cam, rotate=PITCHED_ROTATE, light_pos=light, aspect=aspect, view=view,
                      lights=lights_4x4(cam, y=float(surface_at_cam + 4)), cam_rotation=(0.5, 0.6))


def edit_centres(n, seed=12345):
    """C5: n destruction centres from the MT19937 stream seeded like std::mt19937(seed) (numpy's legacy
    RandomState uses the same init_genrand): x,z in [20,491], y in [30,44]."""
    raw = np.random.RandomState(seed).randint(0, 2 ** 32, size=3 * n, dtype=np.uint64).reshape(n, 3)
    c = np.empty((n, 3), np.int32)
    c[:, 0] = 20 + raw[:, 0] % 472
    c[:, 1] = 30 + raw[:, 1] % 15
    c[:, 2] = 20 + raw[:, 2] % 472
    return c


def fnv1a64(a):
    """FNV-1a-64 of an array's bytes (the grid fingerprints of SURVEY.md 8c); vectorised per byte position is not
    possible for FNV, so this walks 1 MiB chunks through a small C-speed loop via int.from_bytes-free arithmetic."""
    import ctypes
    data = np.ascontiguousarray(a).view(np.uint8).ravel()
    h = 1469598103934665603
    # pure Python would take minutes for 96 MiB; use the product library's helper when loaded
    from . import api
    lib = api.load_library()
    return int(lib.vxrt_fnv1a64(data.ctypes.data_as(ctypes.c_void_p), data.size))


# ---- ray / byte accounting -----------------------------------------------------------
def algorithmic_bytes(stats, width, height):
    """4 B x voxel fetches + 4 B x hit pixels (colour, counted once) + 4 B x pixels (RGBA8 store) + 360 B uniforms."""
    return 4 * int(stats["fetches"]) + 4 * int(stats["hit_pixels"]) + 4 * width * height + 360


def total_rays(stats):
    return int(stats["rays_primary"]) + int(stats["rays_global"]) + int(stats["rays_local"])
