"""Every rank's share of an N-GPU split, one after the other on ONE GPU (no exchange): milliseconds per frame of each rank's kernels
under the default switches, for both partitions (tiles dealt in groups of N with a per-row rotation, tile row r -> rank r % N).  A frame of the split takes as long
as its slowest rank; the spread between the ranks is what no protocol work can remove.
python scripts/rank_balance.py [--worlds 2 4 8] [--workload C3ii_4k]"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                      # noqa: E402
import voxel_rt_b200 as vx        # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--worlds", type=int, nargs="+", default=[2, 4, 8])
ap.add_argument("--workload", default="C3ii_4k")
ap.add_argument("--frames", type=int, default=30)
a = ap.parse_args()
scene, res = a.workload.rsplit("_", 1)
W, H = vx.scenes.RESOLUTIONS[res]
frame = vx.scenes.frame_for(scene, W, H)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for world in a.worlds:
    for rows in (0, 1):
        ms = []
        for rank in range(world):
            ren = vx.Renderer(grid=vx.scenes.DEFAULT_GRID, width=W, height=H, rank=rank, world=world)
            ren.initVoxels(); ren.buildDepthField()
            ren.setPartition(rows)
            ren.updateUniforms(frame)
            stream = torch.cuda.ExternalStream(ren.stream_ptr())
            for _ in range(30):
                ren.draw()
            ren.sync()
            t = []
            for _ in range(a.frames):
                with torch.cuda.stream(stream):
                    flush.fill_(1)
                ren.draw()
                t.append(ren.stats()["ms_total"])
            ms.append(round(statistics.mean(t), 4))
            fused = ren.frameWasFused()
            ren.sync(); ren.close()
        print(json.dumps({"workload": a.workload, "world": world, "partition": "tile rows" if rows else "tiles", "fused": fused,
                          "ms_per_rank": ms, "max": max(ms), "mean": round(statistics.mean(ms), 4),
                          "max_over_mean": round(max(ms) / statistics.mean(ms), 3)}), flush=True)
del flush
torch.cuda.synchronize()
