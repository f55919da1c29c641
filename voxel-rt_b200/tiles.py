"""Sort-first image-tile partition: the frame is cut into 32x8-pixel tiles, dealt in groups of `world`
consecutive tiles, one to each rank (rotated per tile row, see tile_owner), each rank renders its tiles into a compact [nlocal][8][32] RGBA8 buffer (the gather layout),
one all-gather collects them and `assemble` un-tiles into the raster frame.  This module is the host-side
(numpy) statement of that mapping; libvxrt's TileMap / assemble_kernel implement the same arithmetic.

rows=True is the other partition (vxrt_set_partition 1): whole tile ROWS are interleaved (tile row r belongs to rank
r % world), a rank's buffer is its 8-row strips back to back, raster inside: [local rows * 8][width] pixels, padded to
local_tiles * 256 entries."""
import numpy as np

TILE_W, TILE_H = 32, 8


def tile_counts(width, height):
    tx = (width + TILE_W - 1) // TILE_W
    ty = (height + TILE_H - 1) // TILE_H
    return tx, ty, tx * ty


def local_tiles(width, height, world, rows=False):
    tx, ty, n = tile_counts(width, height)
    if rows and world > 1:
        return ((ty + world - 1) // world) * tx
    return (n + world - 1) // world


def tile_owner(t, tx, world):
    """rank of global tile t: groups of `world` consecutive tiles are dealt one to each rank, rotated by the tile row the group
    starts in (a rank's tiles then do not line up in columns when tx is a multiple of world)"""
    t = np.asarray(t)
    j = t // world
    return (t % world - ((j * world) // tx) % world) % world


def tiles_of_rank(width, height, rank, world, rows=False):
    tx, ty, ntiles = tile_counts(width, height)
    if rows and world > 1:
        return np.concatenate([np.arange(r * tx, (r + 1) * tx) for r in range(rank, ty, world)] or [np.zeros(0, np.int64)])
    t = np.arange(ntiles)
    return t[tile_owner(t, tx, world) == rank]


def pixel_owner(width, height, world, rows=False):
    """[height][width] array: rank that renders each pixel."""
    tx = tile_counts(width, height)[0]
    py, px = np.mgrid[0:height, 0:width]
    if rows and world > 1:
        return (py // TILE_H) % world
    return tile_owner((py // TILE_H) * tx + px // TILE_W, tx, world)


def extract_local(frame, rank, world, rows=False):
    """raster [H][W][C] -> this rank's gather-layout buffer [nlocal][8][32][C] (padding zero-filled); rows: the rank's strips,
    [nlocal * 256][C] with the strips' raster [local rows * 8][W] at its start."""
    h, w = frame.shape[:2]
    tx, ty, _ = tile_counts(w, h)
    nl = local_tiles(w, h, world, rows)
    if rows and world > 1:
        out = np.zeros((nl * TILE_W * TILE_H,) + frame.shape[2:], frame.dtype)
        for j, r in enumerate(range(rank, ty, world)):
            blk = frame[r * TILE_H:(r + 1) * TILE_H].reshape((-1,) + frame.shape[2:])
            out[j * TILE_H * w:j * TILE_H * w + len(blk)] = blk
        return out
    out = np.zeros((nl, TILE_H, TILE_W) + frame.shape[2:], frame.dtype)
    for j, t in enumerate(tiles_of_rank(w, h, rank, world)):
        x0, y0 = (t % tx) * TILE_W, (t // tx) * TILE_H
        blk = frame[y0:y0 + TILE_H, x0:x0 + TILE_W]
        out[j, :blk.shape[0], :blk.shape[1]] = blk
    return out


def assemble(gathered, width, height, rows=False):
    """[world][nlocal][8][32][C] -> raster [H][W][C]; rows: [world][nlocal * 256][C] (strips)."""
    world = gathered.shape[0]
    tx = tile_counts(width, height)[0]
    py, px = np.mgrid[0:height, 0:width]
    if rows and world > 1:
        g = gathered.reshape((world, -1) + gathered.shape[-1:]) if gathered.ndim > 2 else gathered
        row = py // TILE_H
        return g[row % world, ((row // world) * TILE_H + py % TILE_H) * width + px]
    t = (py // TILE_H) * tx + px // TILE_W
    return gathered[tile_owner(t, tx, world), t // world, py % TILE_H, px % TILE_W]
