"""CPU: the three ways the CUDA path's production kernels re-organise fshader.glsl's lighting -- only the active light slots are
walked (with the clamp re-applied for the inactive ones behind the last), rays toward lights a surface faces away from are not
traced, and (wide blocks) the lights of a pixel are evaluated in two halves, the second without the early-out, and combined in
slot order -- stated on the host (oracle/vxo_wide.c, test infrastructure) and compared with the oracle's float frame BIT FOR BIT:
the exactness arguments are checked before any GPU runs the kernels (tests/test_gpu_parity.py::test_wide_tiles_... is the GPU side)."""
import numpy as np
import pytest

import golden_cases as gc
import oracle_lib as ol

W, H = 200, 112


def light_sets():
    cam = gc.CAM
    near = [(cam[0] + dx, 40.0 + (k % 3), cam[2] + 20.0 + dz, w) for k, (dx, dz, w) in enumerate(
        [(-12, 0, 0.5), (-8, 6, 0.5), (-4, 12, 0.5), (0, 18, 0.5), (4, 24, 0.5), (8, 30, 0.5), (12, 36, 0.5), (-10, 40, 0.5),
         (-6, 34, 0.5), (-2, 28, 0.5), (2, 22, 0.5), (6, 16, 0.5), (10, 10, 0.5), (14, 4, 0.5), (-14, 8, 0.5), (0, 44, 0.5)])]
    gaps = [(-1, -1, -1, 0)] * 16
    for slot, l in zip((1, 2, 6, 7, 8, 13, 15), near):
        gaps[slot] = l
    tail_gap = [(-1, -1, -1, 0)] * 16                 # the last active slot is not slot 15: the clamp behind it matters
    for slot, l in zip((0, 3, 4, 9), near):
        tail_gap[slot] = (l[0], l[1], l[2], 0.9)
    return {
        "strong_16": near,                                                        # the clamp ends the loop in the first half
        "weak_16": [(x, y, z, 0.07) for x, y, z, _ in near],                      # ... late or never
        "mixed_16": [(x, y, z, 0.02 if k < 8 else 0.6) for k, (x, y, z, _) in enumerate(near)],   # ... in the second half
        "gaps_7": gaps,
        "tail_gap_4": tail_gap,
        "odd_5": near[:5],
        "one": near[:1],
        "weird": [(190.0, 40.0, 170.0, -0.5), (200.0, 45.0, 180.0, float("inf")), (185.0, 39.0, 165.0, float("nan")),
                  (205.0, 38.0, 160.0, 1e38), (195.0, 60.0, 175.0, 0.0), (192.0, 41.0, 172.0, 0.3)],
        "none": None,
    }


@pytest.fixture(scope="module")
def frames():
    aspect = np.float32(W) / np.float32(H)
    out = {n: ol.make_frame(gc.CAM, aspect=aspect, lights=ls) for n, ls in light_sets().items()}
    out["pitched"] = gc.frame_cases(W, H)["C3ii_pitched"]
    out["low_sun"] = gc.frame_cases(W, H)["low_sun"]
    out["step_count_view"] = gc.frame_cases(W, H)["C3i"]
    return out


@pytest.mark.parametrize("skip_dark,wide", [(False, False), (True, False), (False, True), (True, True)],
                         ids=["active_slots_only", "unlit_rays_skipped", "two_halves", "two_halves_unlit_skipped"])
def test_production_lighting_equals_the_oracle_bit_for_bit(oracle, default_level, frames, skip_dark, wide):
    for name, fr in frames.items():
        want = oracle.render(default_level, gc.DIMS, fr, W, H, want_f32=True)["rgba_f32"]
        got = oracle.wide_render(default_level, gc.DIMS, fr, W, H, skip_dark, wide)
        same = got.view(np.uint32) == want.view(np.uint32)
        both_nan = np.isnan(got) & np.isnan(want)                   # (NaN weights: a NaN colour on both sides; payloads may differ)
        assert bool((same | both_nan).all()), (name, int((~(same | both_nan)).any(axis=2).sum()))


def test_the_light_sets_exercise_the_clamp_where_they_claim(oracle, default_level, frames):
    """the early-out really fires in the first half / the second half / never for the sets named so (otherwise the test above
    would prove less than it says): read off the oracle's cast masks -- a light slot that was never cast although it is active
    and in range means the loop had ended before it"""
    def slots_cast(name):
        r = oracle.render(default_level, gc.DIMS, frames[name], W, H)
        hit = r["hit_index"] >= 0
        return r["cast_mask"][hit]
    strong = slots_cast("strong_16")
    assert (((strong >> 1) & 0xFF00) == 0).mean() > 0.5            # mostly nothing cast beyond slot 7: ended in the first half
    weak = slots_cast("weak_16")
    assert (((weak >> 1) & 0xFF00) != 0).mean() > 0.5              # lights of the second half are reached
    mixed = slots_cast("mixed_16")
    reached_second = ((mixed >> 1) & 0xFF00) != 0
    not_all = ((mixed >> 1) & 0x8000) == 0
    assert (reached_second & not_all).mean() > 0.2                 # ... and the loop ends inside the second half for many pixels
