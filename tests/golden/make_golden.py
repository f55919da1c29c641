"""Generates tests/golden/golden.json.  Run in the BUILD CONTAINER (needs /root/reference):

    python tests/golden/make_golden.py

Sources of the vectors:
  "ref_host"   : the reference's own level.cpp / render.cpp / controls.cpp compiled unmodified
                 (oracle/_ref/libref_host.so) -- level fingerprints, the doDestroy sequence and its
                 glBufferSubData call pattern, the uniform upload
  "ref_shader" : the reference's fshader.glsl compiled as C++ through its vendored GLM
                 (oracle/_ref/libref_shader.so) -- castRay known answers and whole-frame fColor hashes
  "oracle"     : quantities only the restatement exposes (hit index / shadow masks / counters), produced by
                 oracle/libvxo.so AFTER it matched ref_host / ref_shader bit-for-bit in this same run
All hashes are FNV-1a-64 over the little-endian array bytes.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol          # noqa: E402
import golden_cases as gc        # noqa: E402

W, H = 160, 90


def h64(o, a):
    return "%016x" % o.fnv(np.ascontiguousarray(a))


def main():
    ol.build_oracle(ref=True)
    o = ol.Oracle()
    rh = ol.RefHost()
    rs = ol.RefShader()
    g = {"width": W, "height": H, "dims": list(gc.DIMS)}

    # ---- host half ---------------------------------------------------------------------------------
    v0 = rh.level_nodepth()
    host = {"nodepth": {"fnv": h64(o, v0), "solid": int((v0 >= 0).sum()), "empty": int((v0 == -1).sum())}}
    v1 = rh.init_render()
    host["depth"] = {"fnv": h64(o, v1), "solid": int((v1 >= 0).sum()), "empty": int((v1 == -1).sum()), "field": int((v1 < -1).sum())}
    vals, cnt = np.unique(v1[v1 < -1], return_counts=True)
    host["depth_values"] = {"%08x" % (int(a) & 0xffffffff): int(c) for a, c in zip(vals, cnt)}
    host["spot"] = {"10,0,10": int(v1[10 + 512 * 0 + 512 * 96 * 10]), "10,30,10": int(v1[10 + 512 * 30 + 512 * 96 * 10]),
                    "10,35,10": int(v1[10 + 512 * 35 + 512 * 96 * 10]), "10,37,10": int(v1[10 + 512 * 37 + 512 * 96 * 10]),
                    "0,0,0": int(v1[0]), "1,0,0": int(v1[1])}
    destroys = []
    for cam, d in (((195.0, 55.0, 155.0), (0.0, -1.0, 0.0)), ((5.0, 55.0, 155.0), (0.0, -1.0, 0.0)),
                   ((300.0, 50.0, 300.0), (0.49552038, -0.47942555, 0.72430015)), ((500.0, 45.0, 500.0), (0.6, -0.5, 0.6244998))):
        off, size, n = rh.do_destroy(cam, d)
        vv = rh.voxels()
        destroys.append({"cam": list(cam), "dir": list(d), "calls": int(n), "bytes": int(size.sum()),
                         "first_offset": int(off[0]) if n else -1, "offsets_fnv": h64(o, off), "sizes_fnv": h64(o, size),
                         "fnv": h64(o, vv), "solid": int((vv >= 0).sum()), "empty": int((vv == -1).sum())})
    host["destroys"] = destroys
    f = ol.make_frame((1.5, 2.5, 3.5), rotate=gc.PITCHED_ROTATE, light_pos=(4, 5, 6), aspect=1.25, view=1, lights=gc.lights_4x4(gc.CAM),
                      cam_rotation=(0.5, 0.6))
    for l in gc.lights_4x4(gc.CAM) + [(1.0, 2.0, 3.0, 4.0)]:      # the 17th is silently dropped (render.cpp:375-385)
        rh.place_light(*l)
    u89, view = rh.update_uniforms(f)
    assert np.array_equal(u89, f.to89()) and view == 1
    host["uniforms_fnv"] = h64(o, u89)
    rot, cdir = rh.mouse_look(0.5, 0.6)
    host["pitched_rotate_hex"] = [float(x).hex() for x in rot]
    host["pitched_dir_hex"] = [float(x).hex() for x in cdir]
    dist, xyz = o.depth_offsets()
    host["depth_offsets"] = {"count": int(len(dist)), "dist_fnv": h64(o, dist), "xyz_fnv": h64(o, xyz)}
    g["ref_host"] = host

    # the oracle must reproduce the host half before anything it says is recorded
    ov = o.default_level(depth_field=False)
    assert h64(o, ov) == host["nodepth"]["fnv"]
    o.compute_depth_field(ov, gc.DIMS)
    assert h64(o, ov) == host["depth"]["fnv"]
    level = ov.copy()
    for dd in destroys:
        o.do_destroy(ov, gc.DIMS, dd["cam"], dd["dir"])
        assert h64(o, ov) == dd["fnv"], dd

    # ---- shader half -------------------------------------------------------------------------------
    rs.upload(level)
    starts, dirs, dists = gc.kat_rays(4096)
    ret, out7 = rs.cast_rays(starts, dirs, dists)
    oret = np.zeros_like(ret)
    oout = np.zeros_like(out7)
    for i in range(len(dists)):
        r, hp, hn, st = o.cast_ray(level, gc.DIMS, starts[i], dirs[i], dists[i])
        oret[i] = r
        oout[i, :3] = hp; oout[i, 3:6] = hn; oout[i, 6] = st
    assert np.array_equal(ret, oret)
    assert np.array_equal(out7.view(np.uint32), oout.view(np.uint32))
    g["ref_shader"] = {"kat": {"n": 4096, "seed": 7, "ret_fnv": h64(o, ret), "out7_fnv": h64(o, out7), "hits": int((ret >= 0).sum()),
                               "first16_ret": [int(x) for x in ret[:16]], "first16_steps": [float(x) for x in out7[:16, 6]]}}
    frames = {}
    oracle_frames = {}
    for name, fr in gc.frame_cases(W, H).items():
        rs.set_frame(fr)
        rgba, steps = rs.render(W, H)
        out = o.render(level, gc.DIMS, fr, W, H, want_f32=True)
        assert np.array_equal(rgba.view(np.uint32), out["rgba_f32"].view(np.uint32)), name
        assert np.array_equal(ol.unorm8(rgba), out["rgba8"]), name
        frames[name] = {"fcolor_fnv": h64(o, rgba), "rgba8_fnv": h64(o, ol.unorm8(rgba)), "total_steps": float(steps.sum(dtype=np.float64))}
        assert float(steps.sum(dtype=np.float64)) == float(out["counters"][3]), name
        oracle_frames[name] = {"hit_fnv": h64(o, out["hit_index"]), "steps_fnv": h64(o, out["steps"]),
                               "occl_fnv": h64(o, out["occl_mask"]), "cast_fnv": h64(o, out["cast_mask"]),
                               "counters": [int(x) for x in out["counters"]]}
    g["ref_shader"]["frames"] = frames
    g["oracle"] = {"frames": oracle_frames}

    with open(os.path.join(HERE, "golden.json"), "w") as fjs:
        json.dump(g, fjs, indent=1, sort_keys=True)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
