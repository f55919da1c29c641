"""Shared definitions of the golden cases (used by tests/golden/make_golden.py and by the tests)."""
import numpy as np

import oracle_lib

DIMS = (512, 96, 512)
CAM = (195.0, 55.0, 155.0)
PITCHED_CAM = (195.0, 60.0, 155.0)
PITCHED_ROTATE = [float.fromhex(h) for h in (
    "0x1.a692640000000p-1", "0x0.0p+0", "-0x1.2118d20000000p-1", "0x0.0p+0",
    "0x1.1533700000000p-2", "0x1.c152800000000p-1", "0x1.952ef80000000p-2", "0x0.0p+0",
    "0x1.fb69b20000000p-2", "-0x1.eaee880000000p-2", "0x1.72d7780000000p-1", "0x0.0p+0",
    "0x0.0p+0", "0x0.0p+0", "0x0.0p+0", "0x1.0000000000000p+0")]


def lights_4x4(cam, y=40.0, weight=0.5):
    return [(cam[0] - 36 + 24 * (i % 4), y, cam[2] + 10 + 24 * (i // 4), weight) for i in range(16)]


def frame_cases(width, height):
    """name -> oracle_lib.Frame; the same poses / lights as voxel_rt_b200.scenes."""
    aspect = np.float32(width) / np.float32(height)
    mk = oracle_lib.make_frame
    sparse = [(-1, -1, -1, 0)] * 16
    sparse[2] = (190.0, 40.0, 170.0, 0.9)           # few lights, gaps between slots, strong weights: exercises the
    sparse[3] = (200.0, 42.0, 175.0, 0.9)           # overbright clamp on inactive iterations (fshader.glsl:161-164)
    sparse[9] = (195.0, 39.0, 180.0, 0.7)
    sparse[15] = (185.0, 41.0, 165.0, 0.8)
    return {
        "C1": mk(CAM, aspect=aspect),
        "C2": mk(CAM, aspect=aspect, lights=lights_4x4(CAM)),
        "C3i": mk(CAM, aspect=aspect, lights=lights_4x4(CAM), view=1),
        "C3ii_pitched": mk(PITCHED_CAM, rotate=PITCHED_ROTATE, aspect=aspect, lights=lights_4x4(PITCHED_CAM), cam_rotation=(0.5, 0.6)),
        "sparse_lights": mk((192.0, 50.0, 150.0), rotate=PITCHED_ROTATE, aspect=aspect, lights=sparse),
        "low_sun": mk(CAM, aspect=aspect, light_pos=(900.0, 120.0, 256.0), lights=lights_4x4(CAM)[:5]),
    }


def kat_rays(n, seed=7):
    """seeded castRay inputs incl. degenerate ones (axis-parallel directions, integer starts, starts outside the
    grid, zero direction, tiny components)"""
    rs = np.random.RandomState(seed)
    starts = np.empty((n, 3), np.float32)
    starts[:, 0] = rs.uniform(1, 510, n)
    starts[:, 1] = rs.uniform(30, 80, n)
    starts[:, 2] = rs.uniform(1, 510, n)
    d = rs.normal(size=(n, 3)).astype(np.float32)
    d /= np.sqrt((d * d).sum(1, keepdims=True)).astype(np.float32)
    dists = np.where(rs.rand(n) < 0.5, 384, rs.randint(1, 66, n)).astype(np.int32)
    k = 0
    for axis in range(3):                             # axis-parallel directions (sign() == 0 on two axes)
        for s in (1.0, -1.0):
            d[k] = 0.0
            d[k, axis] = s
            k += 1
    for i in range(6):                                # one zero component
        d[k, i % 3] = 0.0
        k += 1
    d[k] = 0.0; k += 1                                # zero direction
    d[k] = (-0.0, -1.0, 0.0); k += 1                  # negative zero
    starts[k] = (100.0, 60.0, 100.0); k += 1          # integer start
    starts[k] = (100.0, 37.0, 100.0); d[k] = (0.6, -0.8, 0.0); k += 1      # start on a voxel boundary above the grass
    starts[k] = (-5.0, 50.0, 100.0); d[k] = (1.0, 0.0, 0.0); k += 1        # outside the grid
    starts[k] = (600.0, 50.0, 100.0); d[k] = (-0.70710677, -0.1, 0.70003572); k += 1
    starts[k] = (255.5, 200.0, 255.5); d[k] = (0.001, -0.999, 0.002); k += 1
    d[k] = (1e-7, -1.0, 1e-8); k += 1                 # tiny components
    d[k] = (1e-20, -1.0, -1e-30); k += 1
    starts[k] = (0.5, 50.0, 0.5); d[k] = (-0.5, -0.5, -0.70710677); k += 1  # leaves through the low faces
    return starts, d, dists
