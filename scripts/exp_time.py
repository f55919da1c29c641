"""Kernel experiment harness: times the frame kernels of one libvxrt build (VXRT_LIB=path selects it).
    VXRT_LIB=/tmp/x.so python scripts/exp_time.py [--frames 30] [--workloads C3ii_4k,C3i_4k]
Prints per workload: primary / shade / total ms (CUDA events inside vxrt_render, counters off, L2 flushed)."""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                      # noqa: E402
import voxel_rt_b200 as vx        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=30)
ap.add_argument("--workloads", default="C3ii_4k,C3i_4k,C2_1080p")
ap.add_argument("--no-cull", action="store_true")
ap.add_argument("--world", type=int, default=1, help="emulate one rank of an N-GPU split on this GPU (rank 0's tiles only)")
ap.add_argument("--prefetch", type=int, default=2)
ap.add_argument("--no-flush", action="store_true")
ap.add_argument("--no-order", action="store_true")
ap.add_argument("--e2e", action="store_true", help="also sweep the read-back band count of the end-to-end call")
ap.add_argument("--gaps", action="store_true", help="back-to-back frames: wall clock per frame against the kernels' own time (launch / sync overhead)")
a = ap.parse_args()
W0, H0 = 3840, 2160
ren = vx.Renderer(grid=vx.scenes.DEFAULT_GRID, width=W0, height=H0, rank=0, world=a.world)
ren.setL2Prefetch(a.prefetch)
ren.setTileOrdering(not a.no_order)
ren.initVoxels(); ren.buildDepthField()
ren.setCulling(not a.no_cull)
assert vx.scenes.fnv1a64(ren.downloadGrid()) == 0x4c58cc4001a22afa
stream = torch.cuda.ExternalStream(ren.stream_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for wl in a.workloads.split(","):
    scene, res = wl.rsplit("_", 1)
    W, H = vx.scenes.RESOLUTIONS[res]
    if (W, H) != (ren.width, ren.height):
        ren.reshape(W, H)
    ren.updateUniforms(vx.scenes.frame_for(scene, W, H))
    ren.setStats(True); ren.draw(); st = ren.stats(); rays = vx.scenes.total_rays(st)
    ren.setStats(False)
    for _ in range(20):
        ren.draw()
    ren.sync()
    p, s, t = [], [], []
    for _ in range(a.frames):
        if not a.no_flush:
            with torch.cuda.stream(stream):
                flush.fill_(1)
        ren.draw()
        x = ren.stats()
        p.append(x["ms_primary"]); s.append(x["ms_shadow"]); t.append(x["ms_total"])
    m = statistics.mean(t)
    print("%-14s primary %.4f  shade %.4f  total %.4f ms  (min %.4f)  %.1f Mrays/s" % (wl, statistics.mean(p), statistics.mean(s), m, min(t), rays / m / 1e3))
import time
if a.gaps:
    # launch overhead: N frames queued back to back (no flush, no read-back) and the pipelined host call, wall clock
    # per frame against the device time of one frame's kernels
    fr = vx.scenes.frame_for("C3ii", ren.width, ren.height)
    ren.updateUniforms(fr)
    for _ in range(20):
        ren.draw()
    ren.sync()
    N = 400
    t0 = time.perf_counter()
    for _ in range(N):
        ren.draw()
    t_queue = time.perf_counter() - t0
    ren.sync()
    t_all = time.perf_counter() - t0
    k = ren.stats()["ms_total"]
    print("gaps: vxrt_render x%d  host queueing %.4f ms/frame  wall %.4f ms/frame  kernels (events, last frame) %.4f ms" % (N, 1e3 * t_queue / N, 1e3 * t_all / N, k))
    bufs = [ren.hostFrameBuffer(), ren.hostFrameBuffer()]
    for i in range(8):
        ren.submitFrameHost(fr, bufs[i & 1])
    ren.waitFrames()
    t0 = time.perf_counter()
    for i in range(N):
        ren.submitFrameHost(fr, bufs[i & 1])
    t_queue = time.perf_counter() - t0
    ren.waitFrames()
    t_all = time.perf_counter() - t0
    print("gaps: vxrt_submit_frame_host x%d  host queueing %.4f ms/frame  wall %.4f ms/frame" % (N, 1e3 * t_queue / N, 1e3 * t_all / N))
# end-to-end (host frame params in, host RGBA8 out) vs read-back band count, 4K / 16 lights
W, H = vx.scenes.RESOLUTIONS["4k"]
if (W, H) != (ren.width, ren.height):
    ren.reshape(W, H)
fr = vx.scenes.frame_for("C3ii", W, H)
out = ren.hostFrameBuffer()
for nb in ((1, 2, 4, 6, 8, 12, 16) if a.e2e else ()):
    ren.setReadbackBands(nb)
    for _ in range(5):
        ren.renderFrameHost(fr, out)
    ts = []
    for _ in range(a.frames):
        with torch.cuda.stream(stream):
            flush.fill_(1)
        ren.sync()
        t0 = time.perf_counter()
        ren.renderFrameHost(fr, out)
        ts.append(time.perf_counter() - t0)
    print("e2e bands=%2d  mean %.4f ms  min %.4f ms" % (nb, 1e3 * statistics.mean(ts), 1e3 * min(ts)))
del flush
torch.cuda.synchronize()
ren.close()
