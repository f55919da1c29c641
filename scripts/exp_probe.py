"""One kernel experiment, timed and checked in a process of its own (bench.py's "experiments" object; also usable by hand).
The environment selects the build: nothing (the production library), VXRT_TRAVERSAL=0 (the plain kernels on the reference-layout
grid instead of the traversal grid), or VXRT_LIB=<variant library>.  Renders the benchmark workload like bench.py's timed loop (production kernel variants, L2 flushed
between frames, CUDA events inside vxrt_render) and prints ONE JSON line: per-kernel and per-frame milliseconds and the
FNV-1a-64 of the RGBA8 frame, which the caller compares with the production frame (bit-exactness).
    python scripts/exp_probe.py [--workload C3ii_4k] [--frames 40]"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                      # noqa: E402
import voxel_rt_b200 as vx        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C3ii_4k")
ap.add_argument("--frames", type=int, default=40)
a = ap.parse_args()
scene, res = a.workload.rsplit("_", 1)
W, H = vx.scenes.RESOLUTIONS[res]
ren = vx.Renderer(grid=vx.scenes.DEFAULT_GRID, width=W, height=H)
ren.initVoxels(); ren.buildDepthField()
assert vx.scenes.fnv1a64(ren.downloadGrid()) == 0x4c58cc4001a22afa
frame = vx.scenes.frame_for(scene, W, H)
stream = torch.cuda.ExternalStream(ren.stream_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ren.updateUniforms(frame)
ren.setStats(False)
for _ in range(60):
    ren.draw()
ren.sync()
p, s, t = [], [], []
for _ in range(a.frames):
    with torch.cuda.stream(stream):
        flush.fill_(1)
    ren.draw()
    x = ren.stats()
    p.append(x["ms_primary"]); s.append(x["ms_shadow"]); t.append(x["ms_total"])
rgba = ren.renderFrameHost(frame)
# end to end, pipelined (what bench.py's e2e loop does): frame parameters in, RGBA8 frame out to page-locked host memory every frame,
# the read-back of frame k overlapping the kernels of frame k+1; wall clock per frame, L2 flush included
import time                       # noqa: E402
hb = [ren.hostFrameBuffer(), ren.hostFrameBuffer()]
for k in range(6):
    ren.submitFrameHost(frame, hb[k & 1])
ren.waitFrames()
t0 = time.perf_counter()
for k in range(a.frames):
    with torch.cuda.stream(stream):
        flush.fill_(1)
    ren.submitFrameHost(frame, hb[k & 1])
ren.waitFrames()
ms_e2e = (time.perf_counter() - t0) / a.frames * 1e3
xl = ren.stats()                  # the last pipelined frame's kernels ran while the copy engine read the frame before it
pipelined_last = {"ms_primary": round(xl["ms_primary"], 4), "ms_shade": round(xl["ms_shadow"], 4)}
# the same build on one rank's share of an 8-GPU tile split (tiles t with t % 8 == 0): the regime where the critical path of the
# longest rays, not throughput, sets the time
ren8, share = None, None
try:
    ren8 = vx.Renderer(grid=vx.scenes.DEFAULT_GRID, width=W, height=H, rank=0, world=8)
    ren8.initVoxels(); ren8.buildDepthField()
    stream8 = torch.cuda.ExternalStream(ren8.stream_ptr())
    ren8.updateUniforms(frame)
    ren8.setStats(False)
    for _ in range(60):
        ren8.draw()
    ren8.sync()
    p8, s8, t8 = [], [], []
    for _ in range(a.frames):
        with torch.cuda.stream(stream8):
            flush.fill_(1)
        ren8.draw()
        x = ren8.stats()
        p8.append(x["ms_primary"]); s8.append(x["ms_shadow"]); t8.append(x["ms_total"])
    share = {"ms_primary": round(statistics.mean(p8), 4), "ms_shade": round(statistics.mean(s8), 4), "ms_per_frame": round(statistics.mean(t8), 4)}
except Exception as e:               # the full-frame figures above must survive
    share = {"error": repr(e)[:200]}
out = {"lib": os.path.basename(vx.build.lib_path()), "traversal": os.environ.get("VXRT_TRAVERSAL", "1") != "0", "workload": a.workload,
       "frames": a.frames, "ms_primary": round(statistics.mean(p), 4), "ms_shade": round(statistics.mean(s), 4),
       "ms_per_frame": round(statistics.mean(t), 4), "ms_per_frame_min": round(min(t), 4), "ms_e2e_pipelined": round(ms_e2e, 4),
       "kernels_of_the_last_pipelined_frame": pipelined_last,
       "one_of_8_ranks": share,
       "frame_fnv": "%016x" % vx.scenes.fnv1a64(rgba)}
del flush                         # torch tensors used on the renderers' streams go before the streams do
torch.cuda.synchronize()
if ren8 is not None:
    ren8.close()
ren.close()
print(json.dumps(out))
