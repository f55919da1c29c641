"""One rank's share of an N-GPU tile split on ONE GPU (the regime where the serial tail of the longest blocks, not throughput, sets
the frame time): per-pass and per-frame milliseconds under each switch (overlap of the passes, traversal grid, cold-L2 sweep,
launch ordering) and the distribution of per-block SM cycles.  python scripts/share_probe.py [--world 8] [--workload C3ii_4k]"""
import argparse
import json
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                      # noqa: E402
import voxel_rt_b200 as vx        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--workload", default="C3ii_4k")
ap.add_argument("--frames", type=int, default=40)
a = ap.parse_args()
scene, res = a.workload.rsplit("_", 1)
W, H = vx.scenes.RESOLUTIONS[res]
ren = vx.Renderer(grid=vx.scenes.DEFAULT_GRID, width=W, height=H, rank=a.rank, world=a.world)
ren.initVoxels(); ren.buildDepthField()
frame = vx.scenes.frame_for(scene, W, H)
stream = torch.cuda.ExternalStream(ren.stream_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ren.updateUniforms(frame)


def timed(label, **sw):
    ren.setFusion(sw.get("fused", 0)); ren.setOverlap(sw.get("overlap", 0)); ren.setTraversal(sw.get("trav", True)); ren.setL2Prefetch(sw.get("l2", 2))
    ren.setTileOrdering(sw.get("order", True))
    for _ in range(30):
        ren.draw()
    ren.sync()
    p, s, t = [], [], []
    for _ in range(a.frames):
        with torch.cuda.stream(stream):
            flush.fill_(1)
        ren.draw()
        x = ren.stats()
        p.append(x["ms_primary"]); s.append(x["ms_shadow"]); t.append(x["ms_total"])
    out = {"label": label, "ms_primary": round(statistics.mean(p), 4), "ms_shade": round(statistics.mean(s), 4),
           "ms_frame": round(statistics.mean(t), 4), "ms_frame_min": round(min(t), 4)}
    print(json.dumps(out), flush=True)
    return out


res = [timed("fused, L2 sweep auto", fused=1),
       timed("fused, L2 sweep off", fused=1, l2=0),
       timed("fused, L2 sweep off, traversal off", fused=1, l2=0, trav=False),
       timed("fused, L2 sweep off, ordering off", fused=1, l2=0, order=False),
       timed("two passes: overlap off, traversal on, L2 sweep auto, ordering on"),
       timed("overlap on", overlap=1),
       timed("traversal off", trav=False),
       timed("traversal off, overlap on", trav=False, overlap=1),
       timed("L2 sweep off", l2=0),
       timed("ordering off", order=False),
       timed("ordering off, overlap on", order=False, overlap=1)]
ren.setFusion(0); ren.setOverlap(0); ren.setTraversal(True); ren.setL2Prefetch(2); ren.setTileOrdering(True)
for _ in range(10):
    ren.draw()
ren.sync()
pc, sc = ren.blockCosts()
clock_mhz = 1965.0


def dist(c):
    c = np.sort(c[c > 0].astype(np.float64))[::-1]
    if c.size == 0:
        return {}
    us = c / clock_mhz
    return {"blocks": int(c.size), "sum_block_us": round(float(us.sum()), 1), "max_us": round(float(us[0]), 2),
            "top10_us": [round(float(v), 2) for v in us[:10]], "p99_us": round(float(np.percentile(us, 99)), 2),
            "median_us": round(float(np.median(us)), 2)}


print(json.dumps({"primary_blocks": dist(pc), "shade_blocks": dist(sc), "tiles": int(pc.size)}))
ren.setFusion(1); ren.setL2Prefetch(0)
for _ in range(10):
    ren.draw()
ren.sync()
fc, _ = ren.blockCosts()
print(json.dumps({"fused_blocks": dist(fc)}))
del flush
torch.cuda.synchronize()
ren.close()
