"""voxel-rt hot path for B200 (sm_100a): the reference's fragment-shader work (src/fshader.glsl) as
hand-written CUDA behind a C ABI (include/vxrt.h), plus the host-side mirror of src/render.hpp."""
from .api import (Renderer, HostFrame, Frame, Config, Stats, VxrtError, make_frame, load_library,
                  MAX_LOCAL_LIGHTS, TILE_W, TILE_H)
from . import scenes, tiles, gridfile

__all__ = ["Renderer", "HostFrame", "Frame", "Config", "Stats", "VxrtError", "make_frame", "load_library",
           "MAX_LOCAL_LIGHTS", "TILE_W", "TILE_H", "scenes", "tiles"]
