"""Generates tests/golden/golden_controls.json from the REFERENCE's own controls.cpp (oracle/_ref/libref_host.so,
built by `make -C oracle ref`): for every scripted input sequence of tests/controls_cases.py the FNV-1a-64 of all
per-frame states (camPos, camDir, camRotation, rotateMatrix, viewDepthField as float32 bits), a few sampled states
and the final local-light table; plus lightUpdate (sun) and doMouseLook matrix known answers.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_controls.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import controls_cases as cc   # noqa: E402
import oracle_lib as ol       # noqa: E402


def fnv(a):
    h = 1469598103934665603        # the survey's basis, as everywhere in this repo
    for b in np.ascontiguousarray(a).view(np.uint8).tobytes():
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def main():
    rh = ol.RefHost()
    level = rh.level_nodepth()
    assert ol.Oracle().fnv(level) == 0x2f8d49bd81549f5a
    out = {}
    for name, case in cc.cases().items():
        states, _ = cc.run_case(rh.player_reset, rh.player_step, case)
        samples = {str(i): [float(v).hex() for v in states[i]] for i in (0, len(states) // 3, 2 * len(states) // 3, len(states) - 1)}
        out[name] = dict(frames=len(states), fnv="%016x" % fnv(states), samples=samples,
                         lights=[[float(v).hex() for v in row] for row in rh.lights()])
    out["_sun"] = [[float(v).hex() for v in rh.light_update(*c)] for c in cc.SUN_CASES]
    out["_look"] = [[float(v).hex() for v in np.concatenate(rh.mouse_look(*c))] for c in cc.LOOK_CASES]
    with open(os.path.join(HERE, "golden_controls.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
