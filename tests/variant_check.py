"""Run in a process of its own: frames and known answers of a kernel experiment against the oracle -- a variant library
(VXRT_LIB=..., voxel_rt_b200.build.VARIANTS) or the plain kernels (VXRT_TRAVERSAL=0), chosen by the caller through the environment.  A process of its own because an experiment is new device code: if it faulted, the CUDA
context of the parity tests would be unusable.  Usage: python tests/variant_check.py <label>"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import conftest            # noqa: E402
import golden_cases as gc  # noqa: E402
import oracle_lib as ol    # noqa: E402
import test_oracle_quirks as q  # noqa: E402
import voxel_rt_b200 as vx  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else "default"
    if label == "no_traversal":
        assert os.environ.get("VXRT_TRAVERSAL") == "0"
    elif label != "default":
        assert os.path.basename(os.environ.get("VXRT_LIB", "")) == "libvxrt_exp_%s.so" % label, os.environ.get("VXRT_LIB")
    ol.build_oracle()
    o = ol.Oracle()
    level = conftest.load_default_level(o)
    checked = 0
    with vx.Renderer(grid=gc.DIMS, width=640, height=360, debug=True) as r:
        r.updateGeometry(level)
        for name, W, H in (("C2", 640, 360), ("C3ii_pitched", 416, 240), ("low_sun", 320, 180)):
            fr = gc.frame_cases(W, H)[name]
            r.reshape(W, H)
            vfr = vx.Frame.from_buffer_copy(bytes(fr))
            r.updateUniforms(vfr)
            r.draw()
            rgba, dbg, st = r.readPixels(), r.readDebug(), r.stats()
            ref = o.render(level, gc.DIMS, fr, W, H)
            assert np.array_equal(rgba, ref["rgba8"]), name + ": RGBA8"
            assert np.array_equal(dbg["hit_index"], ref["hit_index"]), name + ": hit index"
            assert np.array_equal(dbg["occl_mask"], ref["occl_mask"]) and np.array_equal(dbg["cast_mask"], ref["cast_mask"]), name + ": shadow masks"
            assert st["fetches"] == int(ref["counters"][3]), name + ": iteration count"
            checked += 1
        ret, out7 = r.castRays(np.array([q.TIE_START] * 2, np.float32), np.array([q.TIE_DIR] * 2, np.float32), np.array([q.TIE_DIST] * 2, np.int32))
        assert [int(v) for v in ret] == [7391987, 7391987] and out7[1][6] == 61.0, "tie-lock ray"
    with vx.Renderer(grid=gc.DIMS, width=640, height=360) as r:          # production variant: counters off, culling on
        r.updateGeometry(level)
        r.setStats(False)
        fr = gc.frame_cases(640, 360)["C2"]
        got = r.renderFrameHost(vx.Frame.from_buffer_copy(bytes(fr)))
        assert np.array_equal(got, o.render(level, gc.DIMS, fr, 640, 360)["rgba8"]), "production frame"
    print("%s ok: %d counted frames, the tie-lock ray, 1 production frame (%s)" % (label, checked, vx.build.lib_path()))


if __name__ == "__main__":
    main()
