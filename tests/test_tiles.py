"""CPU: the sort-first image-tile partition and its gather layout, incl. a world_size-2 gloo run where each
rank renders its own tiles (with the ORACLE standing in for the GPU) and the all-gathered, assembled frame must
equal the single-rank frame."""
import os
import sys

import numpy as np
import pytest

import golden_cases as gc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("size", [(1280, 720), (100, 37), (32, 8), (33, 9), (3840, 2160)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_every_pixel_once(vx, size, world):
    t = vx.tiles
    w, h = size
    tx, ty, n = t.tile_counts(w, h)
    assert tx * t.TILE_W >= w and ty * t.TILE_H >= h
    allt = np.concatenate([t.tiles_of_rank(w, h, r, world) for r in range(world)])
    assert sorted(allt.tolist()) == list(range(n))
    assert all(len(t.tiles_of_rank(w, h, r, world)) <= t.local_tiles(w, h, world) for r in range(world))
    owner = t.pixel_owner(w, h, world)
    counts = np.bincount(owner.ravel(), minlength=world)
    assert counts.sum() == w * h
    if n >= 64 * world:
        assert counts.max() - counts.min() <= 2 * t.TILE_W * t.TILE_H * max(1, tx // world + 1)


@pytest.mark.parametrize("size", [(100, 37), (256, 64)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_extract_assemble_round_trip(vx, size, world):
    t = vx.tiles
    w, h = size
    frame = np.random.RandomState(1).randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    gathered = np.stack([t.extract_local(frame, r, world) for r in range(world)])
    assert gathered.shape[1] == t.local_tiles(w, h, world)
    assert np.array_equal(t.assemble(gathered, w, h), frame)


@pytest.mark.parametrize("size", [(100, 37), (256, 64), (3840, 2160), (33, 9)])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_tile_row_partition(vx, size, world):
    """vxrt_set_partition 1: tile rows interleaved, a rank's buffer = its 8-row strips, raster inside"""
    t = vx.tiles
    w, h = size
    tx, ty, n = t.tile_counts(w, h)
    allt = np.concatenate([t.tiles_of_rank(w, h, r, world, rows=True) for r in range(world)])
    assert sorted(allt.tolist()) == list(range(n))
    assert all(len(t.tiles_of_rank(w, h, r, world, rows=True)) <= t.local_tiles(w, h, world, rows=True) for r in range(world))
    owner = t.pixel_owner(w, h, world, rows=True)
    assert all((owner[y] == owner[y, 0]).all() for y in range(h))           # whole rows
    if w * h <= 256 * 64:
        frame = np.random.RandomState(2).randint(0, 256, size=(h, w, 4)).astype(np.uint8)
        gathered = np.stack([t.extract_local(frame, r, world, rows=True) for r in range(world)])
        assert gathered.shape[1] == t.local_tiles(w, h, world, rows=True) * 256
        assert np.array_equal(t.assemble(gathered, w, h, rows=True), frame)


@pytest.mark.parametrize("size", [(3840, 2160), (1920, 1080), (416, 236), (100, 37), (33, 9), (32, 8)])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("rows", [False, True], ids=["tiles", "tile_rows"])
def test_library_partition_equals_the_host_statement(vx, size, world, rows):
    """the library's own partition arithmetic (kernels.cuh tile_of / tile_owner, the table the render kernels read; exported host-only
    as vxrt_partition_tile / vxrt_partition_owner) against tiles.py, for every local tile of every rank and every tile's owner"""
    lib = vx.load_library()
    t = vx.tiles
    w, h = size
    tx, ty, n = t.tile_counts(w, h)
    nl = t.local_tiles(w, h, world, rows)
    seen = []
    for rank in range(world):
        mine = [lib.vxrt_partition_tile(w, h, rank, world, int(rows), j) for j in range(nl)]
        assert lib.vxrt_partition_tile(w, h, rank, world, int(rows), nl) == -1           # beyond the rank's local tiles
        want = t.tiles_of_rank(w, h, rank, world, rows).tolist()
        assert [g for g in mine if g >= 0] == want, (rank, mine[:8], want[:8])
        if not (rows and world > 1):
            # tile partition: local tile j is the rank's tile of group j (padding only in the last group)
            assert all((g == -1) or (g // world == j) for j, g in enumerate(mine))
        seen += [g for g in mine if g >= 0]
    assert sorted(seen) == list(range(n))
    step = max(1, n // 500)
    tiles = list(range(0, n, step)) + [n - 1]
    owner = t.pixel_owner(w, h, world, rows)
    for g in tiles:
        x0, y0 = (g % tx) * t.TILE_W, (g // tx) * t.TILE_H
        assert lib.vxrt_partition_owner(w, h, world, int(rows), g) == int(owner[y0, x0]), g
    assert lib.vxrt_partition_owner(w, h, world, int(rows), n) == -1


def _gloo_worker(rank, world, port, level_path, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib as ol
    import voxel_rt_b200 as vx
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = ol.Oracle()
        level = np.load(level_path)
        W, H = 160, 90
        fr = gc.frame_cases(W, H)["C2"]
        # this rank's pixels only: rows are rendered in bands, then the rank's tiles are cut out
        owner = vx.tiles.pixel_owner(W, H, world)
        full = o.render(level, gc.DIMS, fr, W, H, nthreads=2)["rgba8"]
        mine = np.where((owner == rank)[..., None], full, 0).astype(np.uint8)      # what this rank "rendered"
        local = torch.from_numpy(vx.tiles.extract_local(mine, rank, world).copy())
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        # edit broadcast: rank 0 decides the edit, every replica applies the same command
        cmd = torch.tensor([150, 36, 150, 7] if rank == 0 else [0, 0, 0, 0], dtype=torch.int32)
        dist.broadcast(cmd, src=0)
        if rank == 0:
            frame = vx.tiles.assemble(np.stack([g.numpy() for g in gathered]), W, H)
            np.save(os.path.join(tmpdir, "frame.npy"), frame)
            np.save(os.path.join(tmpdir, "full.npy"), full)
        np.save(os.path.join(tmpdir, "cmd%d.npy" % rank), cmd.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather_reassembles_the_frame(oracle, default_level, tmp_path):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    level_path = os.path.join(os.environ.get("VXRT_CACHE", "/tmp/vxrt_cache"), "default_level_depth.npy")
    assert os.path.exists(level_path)
    mp.spawn(_gloo_worker, args=(2, port, level_path, str(tmp_path)), nprocs=2, join=True)
    frame = np.load(tmp_path / "frame.npy")
    full = np.load(tmp_path / "full.npy")
    assert np.array_equal(frame, full)
    assert np.array_equal(np.load(tmp_path / "cmd0.npy"), np.load(tmp_path / "cmd1.npy"))
    assert np.load(tmp_path / "cmd1.npy").tolist() == [150, 36, 150, 7]
