#!/usr/bin/env python
"""bench.py -- Mrays/s and ms/frame of the voxel-rt per-pixel path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3ii_4k|C2_1080p|C3i_4k|C3ii_pitched_4k|C1_720p]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's own shader on the host CPU cores

A "step" is one frame: the per-pixel path (primary DDA + shadow/light rays + shading) over every pixel of
the frame.  `value` counts the rays the timed kernels actually trace (primary + lit global + lit local); the same time by the
REFERENCE's casting rule (W*H primary rays, one global-light ray per hit pixel, one local-light ray per (hit
pixel, light) the reference shader would cast) and with every reference ray marched to its end are printed beside it.
N > 1: sort-first image-tile split (groups of N consecutive tiles dealt one to each rank, rotated per tile row), grid
replicated; the ranks' kernels store their pixels straight into rank 0's frame over NVLink peer memory (--exchange nccl: one
all-gather of the RGBA8 tiles per frame + an un-tile kernel); total work is fixed => "scaling": "strong".  e2e: the frame in
rank 0's HOST memory every step (N = 1 pipelined read-back; N = 2 peer memory + one read-back; N >= 4 tile-row partition, every
rank moves its strips into one shared page-locked host frame by DMA).
--partition frames (opt-in): whole frames are the sharded unit instead (frame f on rank f % N, nothing exchanged) => "weak".
Every line carries "parity": the frame the timed public call delivered against the oracle (checker only, after all timing).

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {                     # name -> (scene, resolution key)
    "C3ii_4k": ("C3ii", "4k"),                    # BASELINE configs[2] (ii): 3840x2160, 16 local lights  <- default
    "C3i_4k": ("C3i", "4k"),                      # configs[2] (i): step-count ("depth field") view
    "C3ii_pitched_4k": ("C3ii_pitched", "4k"),    # second camera pose
    "C2_1080p": ("C2", "1080p"),                  # configs[1]
    "C1_720p": ("C1", "720p"),                    # configs[0] (the reference's CPU-runnable case)
    "C4_terrain_4k": ("C4", "4k"),                # configs[3]: synthetic 1024^3 terrain (4 GiB grid per GPU), 16 lights
    "C5_edits_4k": ("C5", "4k"),                  # configs[4]: one right-click edit (r = 7) before every frame, pitched pose
}
METRIC = "Mrays/sec (primary+shadow)"
L2_FLUSH_BYTES = 256 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3ii_4k", choices=sorted(WORKLOADS))
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: 'p2p' = kernels store straight into rank 0's frame over NVLink (no collective); "
                         "'nccl' = all-gather of tile buffers + un-tile kernel")
    ap.add_argument("--e2e-path", default="auto", choices=["auto", "p2p", "host"],
                    help="N > 1, how the frame reaches rank 0's host memory: p2p = peer stores over NVLink + one read-back on rank 0; "
                         "host = every rank's kernels store into one shared page-locked host frame (auto: host from 4 GPUs on)")
    ap.add_argument("--e2e-host-stores", action="store_true",
                    help="host-frame e2e path: kernels store pixels into the host frame themselves (round 1's form) instead of the strip DMA")
    ap.add_argument("--partition", default="tiles", choices=["tiles", "frames"],
                    help="N > 1: 'tiles' (default, BASELINE's split) = every GPU renders its tiles of the SAME frame (strong scaling, "
                         "shorter frame latency); 'frames' = every GPU renders whole frames of its own (frame f on rank f %% N, no "
                         "exchange at all: weak scaling, N times the frame rate at single-GPU latency)")
    ap.add_argument("--bands", type=int, default=2, help="read-back bands of vxrt_render_frame_host (e2e, N = 1)")
    ap.add_argument("--no-cull", action="store_true", help="disable the occupancy-summary culling of certain misses (vxrt_set_culling)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads / cpu baseline (profiling runs)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- level / oracle helpers (checker-side only: cpu_baseline and --impl reference) -------------------
def oracle_handle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle_lib.build_oracle()
    return oracle_lib, oracle_lib.Oracle()


def default_level():
    """reference default level with depth field (fingerprint 4c58cc4001a22afa), built by the product's own device
    builder when a GPU is present (bench arm) -- see make_level_gpu -- or by the oracle (reference arm)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import conftest
    ol, o = oracle_handle()
    return conftest.load_default_level(o)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.proc = None
        self.lines = []
        self.idx = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# =====================================================================================================
def run_b200(args):
    import torch
    import torch.distributed as dist
    import voxel_rt_b200 as vx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: libvxrt has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene, reskey = WORKLOADS[args.workload]
    W, H = vx.scenes.RESOLUTIONS[reskey]

    # ---- grid: every rank builds its replica; level generation + depth field are outside the timed region ----
    build_info = {}
    if scene == "C4":
        grid = vx.scenes.TERRAIN_GRID
        ren = vx.Renderer(grid=grid, width=W, height=H, device=local_rank, rank=rank, world=world)
        t0 = time.perf_counter(); ren.generateTerrain(vx.scenes.TERRAIN_SEED); t1 = time.perf_counter()
        ren.buildDepthField(); t2 = time.perf_counter()
        build_info = {"terrain_generate_s": round(t1 - t0, 3), "depth_field_build_s": round(t2 - t1, 3)}
        frame = vx.scenes.terrain_frame(W, H, ren.terrainHeight(grid[0] // 2, grid[2] // 2))
        level_arr = None
        level_fnv = "terrain seed 0x5EED (integer fbm, include/vxrt.h)"
    else:
        grid = vx.scenes.DEFAULT_GRID
        frame = vx.scenes.frame_for("C3ii_pitched" if scene == "C5" else scene, W, H)
        ren = vx.Renderer(grid=grid, width=W, height=H, device=local_rank, rank=rank, world=world)
        t0 = time.perf_counter(); ren.initVoxels(); t1 = time.perf_counter()      # device level generator (level.cpp:82-138)
        ren.buildDepthField(); t2 = time.perf_counter()                           # device depth-field builder (render.cpp:273-286)
        build_info = {"level_generate_s": round(t1 - t0, 3), "depth_field_build_s": round(t2 - t1, 3)}
        level_arr = ren.downloadGrid()
        level_fnv = "%016x" % vx.scenes.fnv1a64(level_arr)
        assert level_fnv == "4c58cc4001a22afa", level_fnv        # the reference level, bit for bit
    edits = vx.scenes.edit_centres(1000) if scene == "C5" else None
    edit_state = {"k": 0}

    ren.setReadbackBands(args.bands)
    ren.setCulling(not args.no_cull)
    use_p2p = world > 1 and args.exchange == "p2p"
    stream = torch.cuda.ExternalStream(ren.stream_ptr(), device=torch.device("cuda", local_rank))
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    local_bytes = ren.local_bytes()
    gathered = final = local_t = None
    if use_p2p:
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            handle.copy_(torch.from_numpy(ren.p2pExport()))
        dist.broadcast(handle, src=0)
        if rank != 0:
            ren.p2pImport(handle.cpu().numpy())
    elif world > 1:
        local_t = torch.empty(0)                                  # placeholder; real tensors below
        gathered = torch.empty(world * local_bytes, dtype=torch.uint8, device="cuda")
        final = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
        # zero-copy view of the renderer's device output buffer as a torch tensor
        def bind_local():
            class _Buf:
                __cuda_array_interface__ = {"shape": (local_bytes,), "typestr": "|u1", "data": (ren.device_rgba8_ptr(), False), "version": 3}
            return torch.as_tensor(_Buf(), device="cuda")
        local_t = bind_local()

    def step_device():
        """one frame with inputs resident: kernels (+ gather + un-tile for N > 1) on the renderer's stream"""
        apply_edit()
        ren.draw()
        if use_p2p:
            if rank == 0:                                         # owner: acquire every rank's completion flag, then release
                p2p_state["ptr"] = ren.p2pWaitFrame()
                if not p2p_state["hold"]:
                    ren.p2pReleaseFrame()
        elif world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, nccl_state["local_t"])
            ren.assembleTiles(gathered.data_ptr(), final.data_ptr())

    nccl_state = {"local_t": local_t}
    p2p_state = {"ptr": 0, "hold": False}

    edits_dev = edit_cmd = None
    if edits is not None and world > 1:
        # only rank 0 knows the edits (it is the one with the mouse); every other rank sees them through the broadcast alone
        e4 = np.concatenate([edits, np.full((len(edits), 1), 7, np.int32)], axis=1) if rank == 0 else np.zeros((len(edits), 4), np.int32)
        edits_dev = torch.from_numpy(np.ascontiguousarray(e4)).to("cuda")
        edit_cmd = torch.zeros(4, dtype=torch.int32, device="cuda")

    def apply_edit():
        if edits is not None:                                     # C5: right-click destruction before every frame
            k = edit_state["k"] % len(edits)
            edit_state["k"] += 1
            if world > 1:
                # rank 0 decides; the 16-byte command {cx, cy, cz, r} goes out as ONE NCCL broadcast queued on the render stream and
                # every replica replays it from device memory (vxrt_edit_remove_sphere_cmd): no host synchronisation anywhere
                with torch.cuda.stream(stream):
                    if rank == 0:
                        edit_cmd.copy_(edits_dev[k], non_blocking=True)
                    dist.broadcast(edit_cmd, src=0)
                ren.removeSphereCmd(edit_cmd.data_ptr(), 7)
            else:
                ren.removeSphere([int(edits[k][0]), int(edits[k][1]), int(edits[k][2])], 7)

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)                            # samples through warm-up, timed region and e2e loop
    sampler.start()
    ren.updateUniforms(frame)
    # W warm-up frames as asked, plus a fixed 200 more (same count on every rank: the frames contain collectives)
    # so that clocks settle and the nvidia-smi sampler sees the GPU under load
    for nwarm in range(max(args.warmup, 3) + 200):
        flush_l2(); step_device()
        if nwarm % 16 == 15:
            ren.sync()
    barrier()
    # ray / fetch counts of the frame, from the counted kernel variants (untimed): by the reference's casting rule (mode 1: every
    # ray fshader.glsl casts, marched to its end) and of what the production kernels actually execute (mode 2)
    ren.setStats(1); flush_l2(); step_device(); barrier(); st = ren.stats()
    ren.setStats(2); flush_l2(); step_device(); barrier(); st_exec = ren.stats()
    if args.no_cull:
        st_exec = st
    ren.setStats(0)                                               # production frames: no per-iteration counters
    for nwarm in range(3):
        flush_l2(); step_device()
    barrier()
    # ---- timed region: K frames, each bracketed by CUDA events on the launching stream, L2 flushed in between ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = {"primary": [], "shade": []}
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush_l2()
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        if k % 8 == 7 or k == args.steps - 1:
            ren.sync()
            # per-kernel CUDA-event durations of this frame (events live inside vxrt_render, same stream)
            s = ren.stats(); kern_ms["primary"].append(s["ms_primary"]); kern_ms["shade"].append(s["ms_shadow"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    kern_ms_source = "CUDA events inside vxrt_render, frames of the timed region"
    fused_ms = None
    if frame.view_depth_field != 1 and ren.frameWasFused():
        fused_ms = statistics.mean(kern_ms["primary"]) + statistics.mean(kern_ms["shade"])     # the one kernel of the timed frames
        # the two passes are ONE kernel in this configuration (vxrt_set_fusion, auto): their separate durations come from the same
        # frames rendered once more as two kernels
        kern_ms = {"primary": [], "shade": []}
        ren.setFusion(0)
        for k in range(4 + min(args.steps, 10)):
            flush_l2(); step_device(); ren.sync()
            if k >= 4:
                s = ren.stats(); kern_ms["primary"].append(s["ms_primary"]); kern_ms["shade"].append(s["ms_shadow"])
        ren.setFusion(2)
        kern_ms_source = ("CUDA events inside vxrt_render, the timed region's frames rendered again as two kernels (the timed region runs this share "
                          "of the frame as ONE fused kernel, vxrt_set_fusion auto: primary rays + lighting per tile)")
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    # MAX over ranks of the summed device time; rays / fetches summed over ranks
    rays_local_rank = vx.scenes.total_rays(st)
    per_rank = None
    if world > 1:
        # every rank's own figures, for the reader: mean step (its CUDA events around the whole step) and mean span of its render chain
        # (vxrt_render's events: first to last kernel of the frame, on importers of a peer frame incl. the back-pressure wait)
        mine = torch.tensor([total_ms / args.steps, (fused_ms if fused_ms is not None else
                                                     statistics.mean(kern_ms["primary"]) + statistics.mean(kern_ms["shade"]))],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"step_ms": [round(float(x[0]), 4) for x in allr], "render_chain_ms": [round(float(x[1]), 4) for x in allr],
                    "note": "per rank: mean step of the timed region / mean span of vxrt_render's kernels (two-pass shares: the same frames "
                            "rendered again as two kernels); ms_per_step is the max of step_ms"}
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([rays_local_rank, st["fetches"], st["hit_pixels"], st["rays_primary"], st["rays_global"], st["rays_local"],
                            st["rays_dark"], st_exec["rays_global"], st_exec["rays_local"], st_exec["fetches"], st["fetches_primary"],
                            st_exec["fetches_primary"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        rays, fetches, hits, rp, rg, rl, dark, rg_x, rl_x, fetches_x, fetches_p, fetches_px = (int(x) for x in cnt.tolist())
    else:
        rays, fetches, hits = rays_local_rank, st["fetches"], st["hit_pixels"]
        rp, rg, rl, dark = st["rays_primary"], st["rays_global"], st["rays_local"], st["rays_dark"]
        rg_x, rl_x, fetches_x = st_exec["rays_global"], st_exec["rays_local"], st_exec["fetches"]
        fetches_p, fetches_px = st["fetches_primary"], st_exec["fetches_primary"]
    rays_traced = rp + rg_x + rl_x                               # rays the timed kernels trace (unlit rays are not cast)
    ms_per_step = total_ms / args.steps
    # Mrays/s, whole job.  The headline counts only rays the timed kernels TRACE; the same time with the reference's ray count
    # (every ray fshader.glsl casts for this frame, incl. the unlit ones whose term is exactly 0) is reported beside it
    value = rays_traced / (ms_per_step * 1e-3) / 1e6
    value_ref_rule = rays / (ms_per_step * 1e-3) / 1e6

    # C5: what one edit costs on its own (carve + depth-field repair + traversal-grid repair, and for N > 1 the broadcast of its
    # command), CUDA events around the edit alone, max over ranks
    edit_ms = None
    if edits is not None:
        ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for a, b in ee:
            a.record(stream); apply_edit(); b.record(stream)
        barrier()
        edit_ms = float(sum(a.elapsed_time(b) for a, b in ee)) / len(ee)
        if world > 1:
            t = torch.tensor([edit_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            edit_ms = float(t.item())

    # ---- e2e: the public C-ABI call with HOST buffers (frame params in, RGBA8 frame out), wall clock ----
    host_out = ren.hostFrameBuffer()                              # page-locked (vxrt_host_alloc)
    final_host = torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() if world > 1 else None

    def step_e2e():
        if world == 1:
            apply_edit()
            ren.renderFrameHost(frame, host_out)                  # set_frame + kernels + D2H into page-locked memory + sync
        elif use_p2p:
            ren.updateUniforms(frame)
            p2p_state["hold"] = True                              # keep the frame until it has been copied out
            step_device()
            p2p_state["hold"] = False
            if rank == 0:
                class _F:
                    __cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "|u1", "data": (p2p_state["ptr"], False), "version": 3}
                with torch.cuda.stream(stream):
                    final_host.copy_(torch.as_tensor(_F(), device="cuda"), non_blocking=True)
                ren.p2pReleaseFrame()
            ren.sync()
        else:
            ren.updateUniforms(frame)
            step_device()
            with torch.cuda.stream(stream):
                if rank == 0:
                    final_host.copy_(final, non_blocking=True)
            ren.sync()
    for _ in range(3):
        flush_l2(); step_e2e()
    barrier()
    # (a) synchronous: one frame at a time, wall clock from call to "frame is in host memory"
    e2e_sync_s = 0.0
    for k in range(args.steps):
        flush_l2(); ren.sync()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        step_e2e()
        e2e_sync_s += time.perf_counter() - t0
    # the frame the public call just delivered to host memory (and, where edits / the terrain generator made it, the grid it
    # was rendered from): checked against the oracle after all timing is done ("parity" in the JSON line)
    parity_frame = parity_level = None
    if rank == 0 and not args.no_extra:
        parity_frame = np.array(host_out if world == 1 else final_host.numpy(), copy=True).reshape(H, W, 4)
        if scene in ("C4", "C5"):
            parity_level = ren.downloadGrid()
    # (c) N > 1, frames straight to host memory: no exchange at all -- the kernels of every rank store their tiles' pixels
    # into ONE raster in shared page-locked host memory, each GPU over its own PCIe link; two host frames alternate and
    # rank 0 takes frame k (completion flags of all ranks) while frame k + 1 renders.  Returns False (and the caller
    # falls back to the peer-memory path) if the shared mapping cannot be set up on this box.
    e2e_state = {}

    def host_frame_e2e():
        names = ["/vxrt_bench_%s_%d" % (os.environ.get("MASTER_PORT", "0"), i) for i in range(2)]
        hfs, ok = [], True
        try:
            if rank == 0:
                hfs = [vx.HostFrame(n, W, H, create=True) for n in names]
        except vx.VxrtError as e:
            print("host frames unavailable on rank 0: %s" % e, file=sys.stderr)
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1 and rank != 0:
            try:
                hfs = [vx.HostFrame(n, W, H, create=False) for n in names]
            except vx.VxrtError as e:
                print("host frames unavailable on rank %d: %s" % (rank, e), file=sys.stderr)
                ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            for h in hfs:
                h.close()
            return False
        # tile-row partition for this path: a rank's pixels are 8-row strips of the raster, rendered into a local buffer and moved into
        # the shared host frame by one strided DMA per frame (vxrt_set_partition); --e2e-host-stores keeps the kernels' own stores
        strips = not args.e2e_host_stores
        if strips:
            ren.setPartition(1)
        seq = [0, 0]

        def run(nframes, flush):
            """queue nframes; rank 0 takes each frame as soon as every rank delivered it (while the next one renders)"""
            pending = None
            for k in range(nframes):
                if flush:
                    flush_l2()
                b = k & 1
                seq[b] += 1
                apply_edit()                                          # (C5: the edit command rides the render stream ahead of the frame)
                ren.renderToHostFrame(frame, hfs[b], seq[b])
                if rank == 0 and pending is not None:
                    hfs[pending[0]].wait(world, pending[1]); hfs[pending[0]].release(pending[1])
                pending = (b, seq[b])
            if rank == 0 and pending is not None:
                hfs[pending[0]].wait(world, pending[1]); hfs[pending[0]].release(pending[1])
            ren.sync()
        run(12, False)
        if rank == 0:                                                 # the frame that arrived is the frame
            check = np.array(hfs[1].pixels(), copy=True)
        barrier()
        flush_l2(); ren.sync()
        t0 = time.perf_counter()
        run(args.steps, True)
        e2e_state["s"] = time.perf_counter() - t0
        barrier()
        # the frame that arrived in the shared host frame == the frame the exchange path delivered (sync loop above)
        # (with an edit before every frame no two frames are alike: nothing to compare, the parity object is the check)
        if edits is None:
            e2e_state["frame_check"] = bool(np.array_equal(check, final_host.numpy())) if rank == 0 else True
        e2e_state["api"] = ("every rank: vxrt_render_to_host_frame (host frame params in; " +
                            ("tile-row partition: the rank renders its 8-row strips into a local buffer and ONE strided DMA per frame moves them "
                             "into the shared page-locked host frame over its own PCIe link, overlapped with the next frame's kernels" if strips else
                             "its kernels store its tiles' pixels straight into one shared page-locked host frame over its own PCIe link") +
                            ", no exchange); rank 0: vxrt_host_frame_wait on every rank's "
                            "completion flag, overlapped with the next frame (two host frames alternate); wall clock / K, max over ranks")
        for h in hfs:
            h.close()
        if strips:
            ren.sync()
            ren.setPartition(0)                                       # (re-allocates the local frame: re-bind the all-gather's source)
            if not use_p2p:
                nccl_state["local_t"] = bind_local()
        return True

    # (b) pipelined (N = 1): the same call in its queued form -- every step still passes its frame parameters in and
    # gets its RGBA8 frame out to host memory, but the read-back of frame k overlaps the kernels of frame k+1
    if world == 1:
        host_bufs = [host_out, ren.hostFrameBuffer()]
        for k in range(4):
            apply_edit()
            ren.submitFrameHost(frame, host_bufs[k & 1])
        ren.waitFrames()
        flush_l2(); ren.sync()
        t0 = time.perf_counter()
        for k in range(args.steps):
            flush_l2()
            apply_edit()                                              # (C5: the device-side edit is queued ahead of the frame)
            ren.submitFrameHost(frame, host_bufs[k & 1])
        ren.waitFrames()
        e2e_s = time.perf_counter() - t0
        e2e_api = "vxrt_submit_frame_host x K + vxrt_wait_frames (C ABI): host frame params in, host RGBA8 frame out every step, read-back of frame k overlapped with the kernels of frame k+1; wall clock / K (includes the L2 flush kernels)"
    elif world > 1 and (args.e2e_path == "host" or (args.e2e_path == "auto" and world >= 4)) and host_frame_e2e():
        e2e_s, e2e_api = e2e_state["s"], e2e_state["api"]
    elif use_p2p and edits is None:
        host_bufs = [ren.hostFrameBuffer(full_frame=True), ren.hostFrameBuffer(full_frame=True)] if rank == 0 else None

        def queue_frame(k):
            ren.updateUniforms(frame)
            ren.draw()
            if rank == 0:
                ren.p2pReadback(host_bufs[k & 1])
        for k in range(4):
            queue_frame(k)
        ren.waitFrames()
        barrier()
        flush_l2(); ren.sync()
        t0 = time.perf_counter()
        for k in range(args.steps):
            flush_l2()
            queue_frame(k)
        ren.waitFrames()
        e2e_s = time.perf_counter() - t0
        e2e_api = ("every rank: vxrt_set_frame + vxrt_render (pixels stored into rank 0's frame over NVLink); rank 0: vxrt_p2p_readback "
                   "(acquire flags -> D2H of the whole RGBA8 frame -> release) overlapped with the next frame; wall clock / K, max over ranks")
    else:
        e2e_s = e2e_sync_s
        e2e_api = "per-frame synchronous: host frame params in, host RGBA8 frame out on rank 0; wall clock"
    if world > 1:
        t = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_sync_s = (float(x) for x in t.tolist())
    # the same timed loop with the miss culling switched off, for transparency (every ray marches to its end)
    nocull = None
    if not args.no_cull and not args.no_extra:
        ren.setCulling(False)
        for _ in range(5):
            flush_l2(); step_device()
        barrier()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            flush_l2()
            ev2[k][0].record(stream)
            step_device()
            ev2[k][1].record(stream)
            if k % 8 == 7 or k == args.steps - 1:
                ren.sync()
        barrier()
        t_nc = float(sum(a.elapsed_time(b) for a, b in ev2))
        if world > 1:
            t = torch.tensor([t_nc], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_nc = float(t.item())
        nocull = {"ms_per_step": round(t_nc / args.steps, 4), "value": round(rays / (t_nc / args.steps * 1e-3) / 1e6, 2), "unit": "Mrays/s",
                  "rays_per_frame": rays}
        ren.setCulling(True)
    e2e_value = rays_traced / (e2e_s / args.steps) / 1e6
    clocks = sampler.stop()
    h2d = 360                                                     # the frame parameters (kernel arguments)
    d2h = W * H * 4                                               # the RGBA8 frame (rank 0)

    result = None
    if rank == 0:
        hbm, peak_src = peaks()
        # dominant kernel = the one with the larger share of the frame
        prim_ms, shade_ms = statistics.mean(kern_ms["primary"]), statistics.mean(kern_ms["shade"])
        # rank 0's kernels against rank 0's own work.  HBM bookkeeping: algorithmic bytes = 4 B x castRay iterations +
        # colour read + RGBA8 / hit-record store; stated with the iterations the kernels EXECUTE (achieved / frac) and with the
        # reference's iteration count for the same frame (every ray marched to its end), both named.
        def alg_bytes(stx):
            fpx = stx["fetches_primary"]; fsx = stx["fetches"] - fpx
            return 4 * fpx + 4 * st["rays_primary"] + 360, 4 * fsx + 4 * st["hit_pixels"] + 4 * st["hit_pixels"]
        bytes_primary, bytes_shade = alg_bytes(st_exec)
        bytes_primary_ref, bytes_shade_ref = alg_bytes(st)
        dom = "shade_kernel" if shade_ms >= prim_ms else "primary_kernel"
        dom_ms, dom_bytes, dom_bytes_ref = (shade_ms, bytes_shade, bytes_shade_ref) if dom == "shade_kernel" else (prim_ms, bytes_primary, bytes_primary_ref)
        if fused_ms is not None:                                  # the timed frames run ONE kernel (primary rays + lighting per tile)
            dom, dom_ms, dom_bytes, dom_bytes_ref = "frame_kernel", fused_ms, bytes_primary + bytes_shade, bytes_primary_ref + bytes_shade_ref
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        achieved_ref = dom_bytes_ref / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        # ncu capture of this workload's kernels committed under profiles/ (scripts/ncu_summary.py): DRAM traffic per launch and the
        # issue-slot figures -- the resource that actually binds these kernels
        traffic, issue = None, None
        for prof_name in ("r2_ncu_traffic_%s.json" % args.workload, "r2_ncu_traffic.json", "r1_ncu_traffic.json"):
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", prof_name)))
                if prof.get("workload") == args.workload and world == 1:
                    if dom not in prof["kernels"]:
                        continue
                    kk = prof["kernels"][dom]
                    traffic = int(kk["dram_bytes_read"] + kk["dram_bytes_write"])
                    it_exec = (st_exec["fetches"] - st_exec["fetches_primary"]) if dom == "shade_kernel" else st_exec["fetches_primary"]
                    thread_inst = kk["warp_inst"] * kk["avg_active_threads_per_inst"]
                    issue = {"bound": "issue", "kernel": dom, "source": "profiles/" + prof_name + (" (round-1 kernels: stale)" if prof_name.startswith("r1") else ""),
                             "issue_slot_utilisation_pct": kk["issue_active_pct"], "frac": round(kk["issue_active_pct"] / 100.0, 4),
                             "warp_instructions_per_launch": int(kk["warp_inst"]),
                             "avg_active_threads_per_instruction": kk["avg_active_threads_per_inst"],
                             "executed_iterations_per_launch": int(it_exec),
                             "thread_instructions_per_executed_iteration": round(thread_inst / it_exec, 2) if it_exec else None,
                             "achieved_occupancy_pct": kk["achieved_occupancy_pct"], "l1_hit_pct": kk["l1_hit_pct"], "l2_hit_pct": kk["l2_hit_pct"],
                             "dram_bytes_over_algorithmic_bytes": round(traffic / dom_bytes, 4) if dom_bytes else None}
                    break
            except Exception:
                continue
        roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": hbm, "unit": "GB/s",
                    "frac": round(achieved / hbm, 5), "traffic": traffic, "peak_source": peak_src,
                    "kernel_times": kern_ms_source,
                    "achieved_with_reference_iterations": round(achieved_ref, 2), "frac_with_reference_iterations": round(achieved_ref / hbm, 5),
                    "issue": issue,
                    "note": "HBM bookkeeping figure: algorithmic bytes (4 B x castRay iterations the kernel executes + colour read + RGBA8 store, rank 0's "
                            "tiles) / the kernel's CUDA-event time / measured copy peak.  The gathers are served by L1/L2 (DRAM traffic is a few % of the "
                            "algorithmic bytes): the binding resource is instruction issue, reported in 'issue' from the committed ncu capture",
                    **({"frame_kernel": {"ms": round(fused_ms, 4), "alg_bytes": int(bytes_primary + bytes_shade),
                                         "alg_bytes_reference_iterations": int(bytes_primary_ref + bytes_shade_ref),
                                         "what": "the timed frames: one fused kernel per frame (vxrt_set_fusion auto for this share); 'kernels' below = the "
                                                 "same frames rendered again as two kernels"}} if fused_ms is not None else {}),
                    "kernels": {"primary_kernel": {"ms": round(prim_ms, 4), "alg_bytes": int(bytes_primary), "alg_bytes_reference_iterations": int(bytes_primary_ref)},
                                "shade_kernel": {"ms": round(shade_ms, 4), "alg_bytes": int(bytes_shade), "alg_bytes_reference_iterations": int(bytes_shade_ref)}}}
        result = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "value_counts": "rays the timed kernels trace (primary + lit global + lit local); see value_reference_casting_rule / value_all_rays_marched",
            "value_reference_casting_rule": round(value_ref_rule, 2),
            "value_all_rays_marched": None,
            "ms_per_step": round(ms_per_step, 4),
            "ms_per_step_spread": {"min": round(min(step_ms), 4), "median": round(statistics.median(step_ms), 4), "max": round(max(step_ms), 4),
                                   "of": "rank 0's per-frame CUDA-event times (ms_per_step is the max over ranks of their mean)"},
            **({"per_rank": per_rank} if per_rank else {}),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (reference procedural default level, fnv1a64 %s; fixed camera)" % level_fnv,
            "config": {"workload": args.workload, "grid": list(grid), "width": W, "height": H, "local_lights": 16 if scene != "C1" else 0,
                       "setup": build_info,
                       "view_depth_field": int(frame.view_depth_field),
                       "rays_traced_per_frame": rays_traced, "rays_traced": {"primary": rp, "global": rg_x, "local": rl_x},
                       "iterations_executed_per_frame": fetches_x,
                       "rays_per_frame": rays, "rays_primary": rp, "rays_global": rg,
                       "rays_local": rl, "voxel_fetches_per_frame": fetches, "hit_pixels": hits,
                       "rays_facing_away_from_their_light": dark,
                       "counts": "rays_per_frame / rays_global / rays_local / voxel_fetches_per_frame follow the reference's casting rule (counted kernel "
                                 "variants, every ray marched to its end); rays_traced* / iterations_executed* are what the timed production kernels do; "
                                 "all summed over ranks",
                       "partition": "sort-first 32x8 tiles, every %d consecutive tiles dealt one to each rank (rotated per tile row), grid replicated; frame exchange: %s" % (
                           world, "none (1 GPU)" if world == 1 else ("kernels store into rank 0's frame over NVLink peer memory, release/acquire flags" if use_p2p
                                                                     else "NCCL all-gather of RGBA8 tiles + un-tile kernel")),
                       "miss_culling": ("off" if args.no_cull else "on: a ray ends as a miss once its cell is beyond every grid row holding a solid voxel (occupancy summary) and "
                                        "rays from surfaces facing away from their light (term x max(0,N.L) = 0) are not traced; first-hit voxel and "
                                        "pixels unchanged; ray counts are the reference's; 'without_miss_culling' is the same loop with both off"),
                       "l2": "flushed between timed frames (256 MiB write)", "timing": "CUDA events on the launching stream per frame, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "Mrays/s", "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": e2e_api,
                    "sync_ms_per_step": round(e2e_sync_s / args.steps * 1e3, 4),
                    "sync_note": "vxrt_render_frame_host, one frame at a time (latency figure)",
                    **({"frame_check": e2e_state["frame_check"]} if "frame_check" in e2e_state else {})},
            "gpu_launches": int(args.steps * (st["kernel_launches"] + ((2 if use_p2p else 1) if world > 1 else 0))),
            "roofline": roofline,
            "wall_s_timed_region": round(t_wall, 3),
            "without_miss_culling": nocull,
        }
        if nocull:
            result["value_all_rays_marched"] = nocull["value"]
        if parity_frame is not None:
            result["parity"] = parity_check(frame, parity_frame, level_arr if parity_level is None else parity_level, grid, W, H,
                                            stride=4 if scene != "C4" else 16)
        if scene == "C5":
            result["config"]["edits"] = ("one removeSphere(r=7) per frame, centres from mt19937(12345); rays/fetches are those of the counted frames; "
                                         + ("N > 1: rank 0's 16-byte command is one NCCL broadcast per edit on the render stream, every replica replays it "
                                            "from device memory (vxrt_edit_remove_sphere_cmd), no host synchronisation" if world > 1 else
                                            "vxrt_edit_remove_sphere (device edit: carve + depth-field repair + traversal-grid repair, no upload)"))
            result["config"]["ms_per_edit"] = round(edit_ms, 4) if edit_ms is not None else None
            result["config"]["edit_cmd_error"] = int(ren.editCmdError()) if world > 1 else 0
        if not args.no_extra and world == 1 and scene not in ("C4", "C5"):
            production_fnv = "%016x" % vx.scenes.fnv1a64(host_out)          # the frame the e2e loop delivered
            result["other_workloads"] = extra_workloads(vx, ren, flush_l2, stream, torch)
            result["cpu_baseline"] = cpu_baseline(args.workload, level=level_arr)
            result["experiments"] = experiments(args.workload, production_fnv)
    # teardown order matters: torch tensors that were used on the renderer's stream must be released (their
    # allocator records events on that stream) BEFORE the renderer destroys it
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        nccl_state.clear()
        del gathered, final, local_t
    del flush, final_host, edits_dev, edit_cmd
    p2p_err = ren.p2pError() if use_p2p else 0
    if p2p_err:
        raise SystemExit("peer-memory wait timed out (code %d)" % p2p_err)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ren.close()
    if rank == 0:
        print(json.dumps(result))


def parity_check(frame_vx, got, level, dims, W, H, stride):
    """The frame the timed public call delivered (host memory) against the oracle (oracle/vxo.c, the checker) on every
    `stride`-th block of 8 rows, rendered from the same grid and frame parameters.  Bit-exact or it says how many pixels differ."""
    try:
        import voxel_rt_b200 as vx
        ol, o = oracle_handle()
        fr = ol.Frame()
        C.memmove(C.byref(fr), C.byref(frame_vx), C.sizeof(fr))
        blocks = [b for b in range((H + 7) // 8) if b % stride == 0]
        bad, nrows = 0, 0
        want_rows, got_rows = [], []
        for b in blocks:
            y0, y1 = b * 8, min(H, b * 8 + 8)
            ref = o.render(level, tuple(dims), fr, W, H, y0=y0, y1=y1)["rgba8"][y0:y1]
            bad += int((ref != got[y0:y1]).any(axis=2).sum())
            nrows += y1 - y0
            want_rows.append(ref); got_rows.append(got[y0:y1])
        return {"rows": nrows, "of_rows": H, "pixels_checked": nrows * W, "mismatched_pixels": bad,
                "oracle_rows_fnv": "%016x" % vx.scenes.fnv1a64(np.concatenate(want_rows)),
                "frame_rows_fnv": "%016x" % vx.scenes.fnv1a64(np.concatenate(got_rows)),
                "frame_fnv": "%016x" % vx.scenes.fnv1a64(got),
                "what": "RGBA8 frame delivered to host memory by the e2e call (production kernels) vs oracle/vxo.c on every %d%s block of 8 rows, "
                        "same grid and frame parameters; bit-exact comparison" % (stride, "th")}
    except Exception as e:                                         # the checker must never take the bench line down
        return {"error": repr(e)[:300]}


def extra_workloads(vx, ren, flush_l2, stream, torch, steps=10):
    """secondary single-GPU measurements (same timing discipline), reported beside the headline workload"""
    out = {}
    for name in ("C2_1080p", "C3i_4k", "C3ii_pitched_4k", "C1_720p"):
        scene, reskey = WORKLOADS[name]
        W, H = vx.scenes.RESOLUTIONS[reskey]
        ren.reshape(W, H)
        ren.updateUniforms(vx.scenes.frame_for(scene, W, H))
        ren.setStats(False)
        for _ in range(3):
            flush_l2(); ren.draw()
        ren.sync()
        ms = []
        for _ in range(steps):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); ren.draw(); b.record(stream); ren.sync()
            ms.append(a.elapsed_time(b))
        ren.setStats(True); ren.draw(); st = ren.stats(); ren.setStats(False)
        rays = vx.scenes.total_rays(st)
        m = statistics.mean(ms)
        out[name] = {"ms_per_frame": round(m, 4), "Mrays_per_s": round(rays / (m * 1e-3) / 1e6, 2), "rays_per_frame": rays,
                     "voxel_fetches_per_frame": st["fetches"]}
    return out


def experiments(workload, production_fnv):
    """Kernel experiments that are NOT in the numbers above (separate variant libraries, voxel_rt_b200.build.VARIANTS), each timed
    and checked in a process of its own by scripts/exp_probe.py after the measurements of this line are complete, beside the
    production library through the same probe (the figure to compare with).
    "bit_exact": the probe's frame has the fingerprint of the production frame.  Never raises: a failing experiment is a note."""
    import shutil
    import voxel_rt_b200 as vx
    probe = os.path.join(ROOT, "scripts", "exp_probe.py")
    out = {"note": "not part of value / e2e; same workload, scripts/exp_probe.py, one process each", "production_frame_fnv": production_fnv}

    state = {"timed_out": False}

    def run(label, env):
        if state["timed_out"]:                                     # something systemic: do not spend more of the bench's time
            return {"error": "skipped after an earlier probe timed out"}
        try:
            r = subprocess.run([sys.executable, probe, "--workload", workload], env=dict(os.environ, **env), capture_output=True, text=True, timeout=120)
            if r.returncode != 0:
                return {"error": (r.stderr or r.stdout).strip().splitlines()[-1][:300] if (r.stderr or r.stdout).strip() else "exit %d" % r.returncode}
            d = json.loads(r.stdout.strip().splitlines()[-1])
            d["bit_exact"] = d.pop("frame_fnv") == production_fnv
            return d
        except subprocess.TimeoutExpired:
            state["timed_out"] = True
            return {"error": "probe timed out after 120 s"}
        except Exception as e:
            return {"error": repr(e)[:300]}
    out["production"] = run("production", {})
    out["without_traversal_grid"] = run("without_traversal_grid", {"VXRT_TRAVERSAL": "0"})     # the plain kernels on the reference-layout grid
    import voxel_rt_b200 as vx
    for name in sorted(vx.build.VARIANTS):                         # variant libraries (voxel_rt_b200.build.VARIANTS), built here
        lib = None
        try:
            if not shutil.which(os.environ.get("NVCC", "nvcc")):
                out[name] = {"error": "nvcc unavailable"}
                continue
            lib = vx.build.build_variant(name)
            out[name] = run(name, {"VXRT_LIB": lib})
        except Exception as e:
            out[name] = {"error": repr(e)[:300]}
        finally:
            if lib and os.path.exists(lib):
                os.remove(lib)
    return out


# =====================================================================================================
def run_frame_sharded(args):
    """--partition frames (N > 1, opt-in): whole frames are the sharded unit.  Rank r renders frames r, r + N, r + 2N, ... of
    the stream on its own replica of the grid and reads each back over its own PCIe link; nothing is exchanged between the
    GPUs, so the frame rate scales with N while one frame still takes the single-GPU time.  A step = every rank renders one
    frame (N frames per step): per-GPU work is fixed => "scaling": "weak".  Same kernels, same timing discipline as the
    default tile split (run_b200); NOT yet run on a GPU box (written after round 1's GPU minutes were spent)."""
    import torch
    import torch.distributed as dist
    import voxel_rt_b200 as vx

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus or world < 2:
        raise SystemExit("--partition frames: launch with torch.distributed.run --nproc-per-node N (N = --gpus >= 2)")
    scene, reskey = WORKLOADS[args.workload]
    if scene in ("C4", "C5"):
        raise SystemExit("--partition frames supports the default-level workloads only")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: libvxrt has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, H = vx.scenes.RESOLUTIONS[reskey]
    grid = vx.scenes.DEFAULT_GRID
    frame = vx.scenes.frame_for(scene, W, H)
    ren = vx.Renderer(grid=grid, width=W, height=H, device=local_rank)          # world = 1: this replica renders whole frames
    ren.initVoxels()
    ren.buildDepthField()
    level_fnv = "%016x" % vx.scenes.fnv1a64(ren.downloadGrid())
    assert level_fnv == "4c58cc4001a22afa", level_fnv
    ren.setCulling(not args.no_cull)
    stream = torch.cuda.ExternalStream(ren.stream_ptr(), device=torch.device("cuda", local_rank))
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.fill_(1)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ren.updateUniforms(frame)
    for nwarm in range(max(args.warmup, 3) + 200):
        flush_l2(); ren.draw()
        if nwarm % 16 == 15:
            ren.sync()
    barrier()
    ren.setStats(1); flush_l2(); ren.draw(); ren.sync(); st = ren.stats()       # counts by the reference's casting rule (untimed)
    ren.setStats(2); flush_l2(); ren.draw(); ren.sync(); st_exec = ren.stats()  # what the production kernels trace
    ren.setStats(0)
    for _ in range(3):
        flush_l2(); ren.draw()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = {"primary": [], "shade": []}
    for k in range(args.steps):
        flush_l2()
        ev[k][0].record(stream)
        ren.draw()
        ev[k][1].record(stream)
        if k % 8 == 7 or k == args.steps - 1:
            ren.sync()
            s = ren.stats(); kern_ms["primary"].append(s["ms_primary"]); kern_ms["shade"].append(s["ms_shadow"])
    barrier()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    # e2e: every rank pipelines its own frames to its own page-locked buffers (vxrt_submit_frame_host)
    host_bufs = [ren.hostFrameBuffer(), ren.hostFrameBuffer()]
    for k in range(4):
        ren.submitFrameHost(frame, host_bufs[k & 1])
    ren.waitFrames()
    barrier()
    flush_l2(); ren.sync()
    t0 = time.perf_counter()
    for k in range(args.steps):
        flush_l2()
        ren.submitFrameHost(frame, host_bufs[k & 1])
    ren.waitFrames()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = (float(x) for x in t.tolist())
    clocks = sampler.stop()
    rays_ref_rule = vx.scenes.total_rays(st)                       # per frame; every rank renders the same benchmark frame
    rays = vx.scenes.total_rays(st_exec)                           # rays the timed kernels trace
    ms_per_step = total_ms / args.steps                            # one step = N frames, one per rank
    result = None
    if rank == 0:
        hbm, peak_src = peaks()
        prim_ms, shade_ms = statistics.mean(kern_ms["primary"]), statistics.mean(kern_ms["shade"])
        fs = st_exec["fetches"] - st_exec["fetches_primary"]
        bytes_shade = 4 * fs + 8 * st["hit_pixels"]
        achieved = bytes_shade / (shade_ms * 1e-3) / 1e9 if shade_ms > 0 else 0.0
        result = {
            "metric": METRIC, "value": round(world * rays / (ms_per_step * 1e-3) / 1e6, 2), "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (reference procedural default level, fnv1a64 %s; fixed camera)" % level_fnv,
            "config": {"workload": args.workload, "grid": list(grid), "width": W, "height": H, "local_lights": 16 if scene != "C1" else 0,
                       "rays_traced_per_frame": rays, "rays_per_frame": rays_ref_rule, "frames_per_step": world, "voxel_fetches_per_frame": st["fetches"],
                       "iterations_executed_per_frame": st_exec["fetches"],
                       "partition": "frame-sharded: rank r renders whole frames r, r+N, ... on its own grid replica, no exchange "
                                    "(opt-in; the default is BASELINE's sort-first tile split of one frame)",
                       "frame_latency_ms": round(ms_per_step, 4), "ms_per_frame_throughput": round(ms_per_step / world, 4),
                       "l2": "flushed between timed frames (256 MiB write)", "timing": "CUDA events on the launching stream per frame, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": round(world * rays / (e2e_s / args.steps) / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(e2e_s / args.steps * 1e3, 4),
                    "h2d_bytes_per_step": 360 * world, "d2h_bytes_per_step": W * H * 4 * world,
                    "api": "every rank: vxrt_submit_frame_host x K + vxrt_wait_frames (host frame params in, host RGBA8 frame out, own PCIe link); wall clock / K, max over ranks"},
            "gpu_launches": int(args.steps * st["kernel_launches"] * world),
            "roofline": {"bound": "hbm", "kernel": "shade_kernel", "achieved": round(achieved, 2), "peak": hbm, "unit": "GB/s",
                         "frac": round(achieved / hbm, 5), "traffic": None, "peak_source": peak_src,
                         "kernels": {"primary_kernel": {"ms": round(prim_ms, 4)}, "shade_kernel": {"ms": round(shade_ms, 4), "alg_bytes": int(bytes_shade)}}},
        }
    dist.barrier()
    dist.destroy_process_group()
    del flush
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ren.close()
    if rank == 0:
        print(json.dumps(result))


# =====================================================================================================
def reference_frame_runner(workload, stride, level=None):
    """Returns (run_once() -> seconds, rays_in_sample, kind, cores, sample_description).  Uses the reference's own
    shader compiled for the CPU (oracle/_ref/ref_shader_cli, all host cores via fork) when that build travelled
    with the repo, else the oracle port (OpenMP, all cores)."""
    import voxel_rt_b200 as vx
    ol, o = oracle_handle()
    scene, reskey = WORKLOADS[workload]
    W, H = vx.scenes.RESOLUTIONS[reskey]
    fr_vx = vx.scenes.frame_for(scene, W, H)
    fr = ol.Frame()
    C.memmove(C.byref(fr), C.byref(fr_vx), C.sizeof(fr))
    if level is None:
        level = default_level()
    cores = os.cpu_count() or 1
    rows = [y for y in range(H) if ((y >> 3) % stride) == 0]
    # rays of the sample, counted by the oracle (untimed)
    counters = np.zeros(5, np.uint64)
    blocks = sorted(set(y >> 3 for y in rows))
    for b in blocks:
        out = o.render(level, (512, 96, 512), fr, W, H, y0=b * 8, y1=min(H, b * 8 + 8))
        counters += out["counters"]
    rays = int(counters[0] + counters[1] + counters[2])
    sample = "%s: %d of %d rows (8-row blocks, every %d%s block), %d rays" % (workload, len(rows), H, stride, "th" if stride > 1 else "", rays)
    cli = os.path.join(ROOT, "oracle", "_ref", "ref_shader_cli")
    if os.path.exists(cli) and os.access(cli, os.X_OK):
        tmp = tempfile.mkdtemp(prefix="vxrt_ref_")
        gpath, fpath = os.path.join(tmp, "grid.i32"), os.path.join(tmp, "frame.bin")
        level.tofile(gpath)
        buf = np.zeros(90, np.float32)
        buf[:89] = fr.to89()
        buf[89:90].view(np.int32)[0] = fr.view_depth_field
        buf.tofile(fpath)

        def run(reps):
            r = subprocess.run([cli, gpath, fpath, str(W), str(H), str(cores), str(reps), str(stride)], capture_output=True, text=True, check=True)
            return [float(x) for x in r.stdout.split()]
        return run, rays, "reference", cores, sample + ("; unmodified fshader.glsl compiled for the CPU via the reference's GLM (g++ -O2 -ffp-contract=off, no -march=native: the binary is "
                                                     "built in the build container and travels), %d persistent forked workers, timed between barriers" % cores)

    def run(reps):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            for b in blocks:
                o.render(level, (512, 96, 512), fr, W, H, y0=b * 8, y1=min(H, b * 8 + 8))
            ts.append(time.perf_counter() - t0)
        return ts
    return run, rays, "port", o.num_threads(), sample + "; C restatement (oracle/vxo.c), OpenMP"


def cpu_baseline(workload, level=None):
    """bounded CPU sample of the same workload on this box's host cores (reported, not targeted)"""
    try:
        run, rays, kind, cores, sample = reference_frame_runner(workload, stride=4, level=level)
        ts = run(2)
        t = min(ts)
        return {"value": round(rays / t / 1e6, 3), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample,
                "seconds_per_sample": round(t, 4)}
    except Exception as e:                                         # the baseline must never take the bench line down
        return {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import voxel_rt_b200 as vx
    scene, reskey = WORKLOADS[args.workload]
    W, H = vx.scenes.RESOLUTIONS[reskey]
    # size the per-step sample so that (warmup + steps) samples end within ~2 minutes
    level = default_level()
    run, rays_full, kind, cores, _ = reference_frame_runner(args.workload, stride=16, level=level)
    t_probe = min(run(1)) * 16.0                                  # ~ full-frame seconds
    budget = 120.0
    n = args.steps + args.warmup
    stride = 1
    while stride < 64 and t_probe / stride * n > budget:
        stride *= 2
    run, rays, kind, cores, sample = reference_frame_runner(args.workload, stride=stride, level=level)
    if args.warmup:
        run(args.warmup)
    ts = run(args.steps)
    ms = statistics.mean(ts) * 1e3
    value = rays / (ms * 1e-3) / 1e6
    frame = vx.scenes.frame_for(scene, W, H)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (reference procedural default level, fnv1a64 4c58cc4001a22afa; fixed camera)",
        "config": {"workload": args.workload, "grid": [512, 96, 512], "width": W, "height": H, "local_lights": 16 if scene != "C1" else 0,
                   "view_depth_field": int(frame.view_depth_field), "rays_per_step_sample": rays},
        "cpu_baseline": {"value": round(value, 3), "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.partition == "frames" and a.gpus > 1:
        run_frame_sharded(a)
    else:
        run_b200(a)
