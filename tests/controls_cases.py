"""Scripted input sequences for the host gameplay code (controls.cpp): shared by tests/test_host_logic.py and
tests/golden/make_golden_controls.py.  A case = start pose + fps + per-frame (keys[9], mouse) list; keys follow
window.hpp:8-18 (W S A D T SPACE SHIFT LMB RMB).  RMB (destruction) is never pressed here: doDestroy is covered by
the grid tests."""
import numpy as np

W, S, A, D, T, SPACE, SHIFT, LMB, RMB = range(9)
SCREEN = (800, 600)


def _script(seed, frames, fps):
    rng = np.random.default_rng(seed)
    out = []
    held = np.zeros(9, np.uint8)
    mouse = [400, 300]
    for f in range(frames):
        if f % 20 == 0:                       # change the held movement keys every 20 frames
            held[:4] = rng.random(4) < 0.45
            held[LMB] = rng.random() < 0.6
            mouse = [int(rng.integers(0, SCREEN[0])), int(rng.integers(0, SCREEN[1]))]
        keys = held.copy()
        keys[SPACE] = rng.random() < 0.05
        keys[T] = rng.random() < 0.02
        keys[SHIFT] = rng.random() < 0.02
        out.append((keys.tolist(), (mouse[0], mouse[1], SCREEN[0], SCREEN[1])))
    return out


def cases():
    """name -> dict(cam, dir, camrot, fps, frames)"""
    c = {}
    c["spawn_walk_60"] = dict(cam=(195.0, 55.0, 155.0), dir=(0.0, 0.0, 1.0), camrot=(0.0, 0.0), fps=60, frames=_script(1, 900, 60))
    c["spawn_walk_144"] = dict(cam=(195.0, 55.0, 155.0), dir=(0.0, 0.0, 1.0), camrot=(0.0, 0.0), fps=144, frames=_script(2, 1500, 144))
    c["corner_30"] = dict(cam=(3.5, 60.0, 3.5), dir=(0.6, 0.0, 0.8), camrot=(0.2, 0.6), fps=30, frames=_script(3, 600, 30))
    c["trees_60"] = dict(cam=(300.25, 70.0, 260.75), dir=(-0.8, 0.0, 0.6), camrot=(0.0, -0.9), fps=60, frames=_script(4, 900, 60))
    c["far_edge_75"] = dict(cam=(508.0, 90.0, 505.0), dir=(1.0, 0.0, 0.0), camrot=(0.0, 1.5), fps=75, frames=_script(5, 900, 75))
    return c


def run_case(player_reset, player_step, case, take_light=None):
    """drive either implementation; returns (states [frames][25] float32 with the view flag last, lights list)"""
    player_reset(case["cam"], case["dir"], case["camrot"], case["fps"])
    states = np.zeros((len(case["frames"]), 25), np.float32)
    lights = []
    for i, (keys, mouse) in enumerate(case["frames"]):
        st, view = player_step(keys, mouse)
        states[i, :24] = st
        states[i, 24] = view
        if take_light is not None:
            l = take_light()
            if l is not None:
                lights.append(l.copy())
    return states, lights


# lightUpdate (render.cpp:388-402): (fps, start angle in degrees, calls) -- covers day, the 8x night branch and the wrap
SUN_CASES = [(60, -45.0, 1), (60, -45.0, 1000), (30, 179.9, 50), (144, 269.95, 400), (60, 359.99, 10), (75, 200.0, 2000), (240, 90.0, 5000)]
LOOK_CASES = [(0.0, 0.0), (0.5, 0.6), (-1.57, 3.0), (1.2, -6.1), (0.001, 6.28), (-0.7, -2.4)]
