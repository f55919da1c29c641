"""Multi-GPU parity check (run under torchrun on N GPUs): every rank renders its image tiles of the same frame
from its own replica of the grid, one NCCL all-gather collects the RGBA8 tiles, the un-tile kernel assembles the
raster frame on every rank, and rank 0 compares it bit-for-bit with (a) the oracle's full frame and (b) a
single-context render.  Also replays a broadcast edit on every replica and compares grid fingerprints.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import voxel_rt_b200 as vx          # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = 1920, 1080
    frame = vx.scenes.frame_for("C3ii_pitched", W, H)
    ren = vx.Renderer(grid=vx.scenes.DEFAULT_GRID, width=W, height=H, device=local, rank=rank, world=world)
    ren.initVoxels()
    ren.buildDepthField()
    stream = torch.cuda.ExternalStream(ren.stream_ptr(), device=torch.device("cuda", local))
    nbytes = ren.local_bytes()

    class _Buf:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ren.device_rgba8_ptr(), False), "version": 3}
    local_t = torch.as_tensor(_Buf(), device="cuda")
    gathered = torch.empty(world * nbytes, dtype=torch.uint8, device="cuda")
    final = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")

    def render_frame():
        ren.updateUniforms(frame)
        ren.draw()
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(gathered, local_t)
        ren.assembleTiles(gathered.data_ptr(), final.data_ptr())
        ren.sync()
        return final.cpu().numpy()

    # ---- gather-free variant: a second context per rank whose kernels store straight into rank 0's frame ----
    ren2 = vx.Renderer(grid=vx.scenes.DEFAULT_GRID, width=W, height=H, device=local, rank=rank, world=world)
    ren2.initVoxels()
    ren2.buildDepthField()
    handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
    if rank == 0:
        handle.copy_(torch.from_numpy(ren2.p2pExport()))
    dist.broadcast(handle, src=0)
    if rank != 0:
        ren2.p2pImport(handle.cpu().numpy())
    stream2 = torch.cuda.ExternalStream(ren2.stream_ptr(), device=torch.device("cuda", local))
    final2 = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")

    def render_frame_p2p(nframes=1):
        out = None
        for _ in range(nframes):
            ren2.updateUniforms(frame)
            ren2.draw()
            if rank == 0:
                ptr = ren2.p2pWaitFrame()

                class _F:
                    __cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "|u1", "data": (ptr, False), "version": 3}
                with torch.cuda.stream(stream2):
                    final2.copy_(torch.as_tensor(_F(), device="cuda"))
                ren2.p2pReleaseFrame()
        ren2.sync()
        if rank == 0:
            out = final2.cpu().numpy()
        dist.barrier()
        return out

    ok = True
    got = render_frame()
    got_p2p = render_frame_p2p(5)                  # 5 frames > 2 buffers: exercises the back-pressure across processes
    # edit broadcast: rank 0 picks the edits, every replica applies the same commands
    # ren: every edit is a 16-byte command {cx, cy, cz, r} known to rank 0 only, ONE NCCL broadcast per edit queued on the render
    # stream, replayed from device memory on every replica (vxrt_edit_remove_sphere_cmd; no host synchronisation);
    # ren2: the whole list broadcast ahead, host-argument edits
    e4 = np.concatenate([vx.scenes.edit_centres(8), np.full((8, 1), 7, np.int32)], axis=1)
    edits = torch.tensor(e4 if rank == 0 else np.zeros((8, 4), np.int32), dtype=torch.int32, device="cuda")
    cmd = torch.zeros(4, dtype=torch.int32, device="cuda")
    for k in range(8):
        with torch.cuda.stream(stream):
            if rank == 0:
                cmd.copy_(edits[k], non_blocking=True)
            dist.broadcast(cmd, src=0)
        ren.removeSphereCmd(cmd.data_ptr(), 7)
    ren.sync()
    cmd_err = ren.editCmdError()
    dist.broadcast(edits, src=0)
    for c in edits.cpu().numpy():
        ren2.removeSphere(c[:3], int(c[3]))
    got_edit = render_frame()
    got_edit_p2p = render_frame_p2p(3)
    # pipelined owner read-back: 6 frames queued back to back, alternating host buffers
    bufs = [ren2.hostFrameBuffer(full_frame=True), ren2.hostFrameBuffer(full_frame=True)] if rank == 0 else None
    for k in range(6):
        ren2.updateUniforms(frame)
        ren2.draw()
        if rank == 0:
            ren2.p2pReadback(bufs[k & 1])
    ren2.waitFrames()
    dist.barrier()
    readback_ok = bool(np.array_equal(bufs[0], got_edit_p2p) and np.array_equal(bufs[1], got_edit_p2p)) if rank == 0 else True
    # frames straight to host memory: every rank's kernels store into one shared page-locked raster (no exchange)
    names = ["/vxrt_mgc_%s_%d" % (os.environ.get("MASTER_PORT", "0"), i) for i in range(2)]
    if rank == 0:
        hfs = [vx.HostFrame(n, W, H, create=True) for n in names]
    dist.barrier()
    if rank != 0:
        hfs = [vx.HostFrame(n, W, H, create=False) for n in names]
    host_ok = True
    for k in range(6):
        ren.renderToHostFrame(frame, hfs[k & 1], k // 2 + 1)
        if rank == 0 and k >= 1:
            j = k - 1
            hfs[j & 1].wait(world, j // 2 + 1)
            host_ok &= bool(np.array_equal(hfs[j & 1].pixels(), got_edit))
            hfs[j & 1].release(j // 2 + 1)
    if rank == 0:
        hfs[1].wait(world, 3)
        host_ok &= bool(np.array_equal(hfs[1].pixels(), got_edit))
        hfs[1].release(3)
    ren.sync()
    dist.barrier()
    # the same with the tile-row partition: every rank renders its 8-row strips into a local buffer, one strided DMA per frame
    # moves them into the shared host frame (sequence numbers continue); ren2 (peer-memory context) renders a peer frame in that
    # partition too
    ren.setPartition(1)
    strips_ok = True
    for k in range(6, 12):
        ren.renderToHostFrame(frame, hfs[k & 1], k // 2 + 1)
        if rank == 0 and k >= 7:
            j = k - 1
            hfs[j & 1].wait(world, j // 2 + 1)
            strips_ok &= bool(np.array_equal(hfs[j & 1].pixels(), got_edit))
            hfs[j & 1].release(j // 2 + 1)
    if rank == 0:
        hfs[1].wait(world, 6)
        strips_ok &= bool(np.array_equal(hfs[1].pixels(), got_edit))
        hfs[1].release(6)
    ren.sync()
    dist.barrier()
    ren.setPartition(0)
    ren2.setPartition(1)
    got_rows_p2p = render_frame_p2p(3)
    ren2.setPartition(0)
    if rank == 0:
        strips_ok &= bool(np.array_equal(got_rows_p2p, got_edit_p2p))
    for h in hfs:
        h.close()
    p2p_err = ren2.p2pError()
    fnv = vx.scenes.fnv1a64(ren.downloadGrid())
    fnvs = [None] * world
    dist.all_gather_object(fnvs, fnv)
    tfnvs = [None] * world
    dist.all_gather_object(tfnvs, (vx.scenes.fnv1a64(ren.downloadTraversal()), vx.scenes.fnv1a64(ren2.downloadTraversal()), cmd_err))
    if rank == 0:
        import ctypes as C
        import conftest
        import oracle_lib as ol
        o = ol.Oracle()
        level = conftest.load_default_level(o).copy()
        fr = ol.Frame()
        C.memmove(C.byref(fr), C.byref(frame), C.sizeof(fr))
        want = o.render(level, (512, 96, 512), fr, W, H)["rgba8"]
        ok &= bool(np.array_equal(got, want))
        print("frame vs oracle:", np.array_equal(got, want))
        ok &= bool(np.array_equal(got_p2p, want))
        print("peer-memory frame vs oracle:", np.array_equal(got_p2p, want))
        for c in vx.scenes.edit_centres(8):
            o.remove_sphere(level, (512, 96, 512), int(c[0]), int(c[1]), int(c[2]), 7)
        want2 = o.render(level, (512, 96, 512), fr, W, H)["rgba8"]
        ok &= bool(np.array_equal(got_edit, want2))
        print("frame after broadcast edits vs oracle:", np.array_equal(got_edit, want2))
        ok &= bool(np.array_equal(got_edit_p2p, want2)) and p2p_err == 0
        print("peer-memory frame after edits vs oracle:", np.array_equal(got_edit_p2p, want2), "p2p_err", p2p_err)
        ok &= readback_ok
        print("pipelined peer-memory read-back frames intact:", readback_ok)
        ok &= host_ok
        print("frames stored straight into the shared host frame vs gathered frame:", host_ok)
        ok &= strips_ok
        print("tile-row partition: strips moved into the shared host frame by DMA, and the peer-memory frame, equal the gathered frame:", strips_ok)
        ok &= all(f == o.fnv(level) for f in fnvs)
        print("replica fingerprints equal oracle:", all(f == o.fnv(level) for f in fnvs), ["%016x" % f for f in fnvs])
        want_trav, bad = o.trav_build(level, (512, 96, 512))
        tf = o.fnv(want_trav)
        trav_ok = bad == 0 and all(t[0] == tf and t[1] == tf and t[2] == 0 for t in tfnvs)
        ok &= trav_ok
        print("traversal grids of every replica (device-command edits and host-argument edits) equal the host rebuild:", trav_ok)
    dist.barrier()
    dist.destroy_process_group()
    del gathered, final, local_t, final2, handle, cmd, edits
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ren.close()
    ren2.close()
    if rank == 0:
        print("MULTIGPU CHECK", "PASS" if ok else "FAIL")
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
