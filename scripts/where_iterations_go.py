"""CPU analysis (oracle only, no GPU): where do the reference algorithm's castRay iterations go, and how many of them can
NO occupancy structure ever save?  Every ray of a frame is classified by kind (primary / global-light / local-light) and
by how it ended (hit / left the grid / budget exhausted).  A ray that hits must run every iteration (the first-hit voxel
and hitPos depend on the whole float state), so only the iterations of rays that end as misses are avoidable at all;
of those the table shows what the CUDA path already removes (occupancy-summary culling, unlit rays) and what is left for
any finer brick hierarchy; and, with the traversal grid of the CUDA path restated on the host (oracle/vxo_trav.c), how many
iterations become steps of a run that needs no index arithmetic, range test or load.  (Round 1's "clear tube" experiment --
a summed-area table of solids; inexact because of the reference's tie locks -- is recorded in profiles/r1_where_iterations_go.json.)
Usage: python scripts/where_iterations_go.py [--size 3840 2160] [--case C2 C3ii_pitched]
Writes profiles/r2_where_iterations_go.json with --write."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest           # noqa: E402  (cached reference level)
import golden_cases as gc  # noqa: E402
import oracle_lib as ol   # noqa: E402


class Cell(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "iterations", "jumps", "rays_culled", "iterations_after_cull",
                                          "rays_dark", "iterations_dark", "iterations_dark_after_cull")]


class Profile(C.Structure):
    _fields_ = [("cell", (Cell * 3) * 3), ("longest", C.c_uint64 * 3), ("ymin", C.c_int32), ("ymax", C.c_int32)]


KINDS = ("primary", "global_light", "local_light")
OUTCOMES = ("hit", "left_grid", "budget_exhausted")


def profile(o, level, dims, frame, w, h):
    p = Profile()
    o.L.vxo_profile_frame.restype = None
    o.L.vxo_profile_frame(level.ctypes.data_as(C.POINTER(C.c_int32)), ol.Dims(*dims), C.byref(frame), C.c_int(w), C.c_int(h), C.byref(p))
    out = {"ymin": p.ymin, "ymax": p.ymax, "longest": {k: int(p.longest[i]) for i, k in enumerate(KINDS)}, "cells": {}}
    for i, k in enumerate(KINDS):
        for j, oc in enumerate(OUTCOMES):
            c = p.cell[i][j]
            out["cells"]["%s/%s" % (k, oc)] = {n: int(getattr(c, n)) for n, _ in Cell._fields_}
    return out


def summarise(pr):
    cells = pr["cells"]
    total = sum(c["iterations"] for c in cells.values())
    hits = sum(c["iterations"] for k, c in cells.items() if k.endswith("/hit"))
    # shadow / light rays that hit = occluded: they need their iterations unless the surface faces away (dark)
    miss = total - hits
    culled = sum(c["iterations_after_cull"] for k, c in cells.items() if not k.endswith("/hit"))
    dark = sum(c["iterations_dark"] for k, c in cells.items())            # unlit rays, hit or miss: never traced
    dark_miss_culled = sum(c["iterations_dark_after_cull"] for k, c in cells.items() if not k.endswith("/hit"))
    removed = culled + dark - dark_miss_culled
    dark_miss = sum(c["iterations_dark"] for k, c in cells.items() if not k.endswith("/hit"))
    left_miss = miss - culled - (dark_miss - dark_miss_culled)           # misses still marched by the CUDA path
    return {"iterations": total, "by_rays_that_hit": hits, "by_rays_that_miss": miss,
            "removed_by_occupancy_summary": culled, "removed_as_unlit": dark, "removed_total": removed,
            "still_executed": total - removed,
            "still_executed_by_misses": left_miss,
            "upper_bound_gain_of_any_finer_hierarchy_pct": round(100.0 * left_miss / max(1, total - removed), 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs=2, default=[3840, 2160])
    ap.add_argument("--case", nargs="+", default=["C2", "C3ii_pitched", "C3i", "low_sun"])
    ap.add_argument("--write", action="store_true")
    a = ap.parse_args()
    ol.build_oracle()
    o = ol.Oracle()
    level = conftest.load_default_level(o)
    w, h = a.size
    result = {"width": w, "height": h, "grid": list(gc.DIMS), "cases": {}}
    for name in a.case:
        fr = gc.frame_cases(w, h)[name]
        pr = profile(o, level, gc.DIMS, fr, w, h)
        pr["summary"] = summarise(pr)
        result["cases"][name] = pr
        s = pr["summary"]
        print("== %s %dx%d: %d iterations; solid rows %d..%d; longest ray: %s" % (name, w, h, s["iterations"], pr["ymin"], pr["ymax"], pr["longest"]))
        for k, c in pr["cells"].items():
            if c["rays"]:
                print("  %-30s rays %10d  iterations %11d (%.1f/ray, %2.0f%% jumps)  after occupancy cull %10d  unlit %10d" % (
                    k, c["rays"], c["iterations"], c["iterations"] / c["rays"], 100.0 * c["jumps"] / max(1, c["iterations"]),
                    c["iterations_after_cull"], c["iterations_dark"]))
        print("  " + json.dumps(s))
        # the traversal grid of the CUDA path (oracle/vxo_trav.c): how many of the iterations become steps of a run (no index
        # arithmetic, range test or load) -- counted on every ray the reference casts, per ray kind
        trav, bad = o.trav_build(level, gc.DIMS)
        tr = o.trav_render(trav, gc.DIMS, fr, w, h)
        st = [int(v) for v in tr["stats"]]
        pr["traversal_grid"] = {k: {"run_steps": st[i], "checked_steps": st[3 + i], "of_which_jumps": st[6 + i]} for i, k in enumerate(KINDS)}
        pr["traversal_grid"]["unencodable_values"] = bad
        for i, k in enumerate(KINDS):
            print("  traversal grid %-14s run steps %11d  checked steps %11d (jumps %10d)  -> %.1f%% of the iterations need no load" % (
                k, st[i], st[3 + i], st[6 + i], 100.0 * st[i] / max(1, st[i] + st[3 + i])))
    if a.write:
        path = os.path.join(ROOT, "profiles", "r2_where_iterations_go.json")
        with open(path, "w") as f:
            json.dump(result, f, indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main()
