"""Sort-first image-tile partition: the frame is cut into 32x8-pixel tiles, tile t belongs to
rank t % world, each rank renders its tiles into a compact [nlocal][8][32] RGBA8 buffer (the gather layout),
one all-gather collects them and `assemble` un-tiles into the raster frame.  This module is the host-side
(numpy) statement of that mapping; libvxrt's TileMap / assemble_kernel implement the same arithmetic."""
import numpy as np

TILE_W, TILE_H = 32, 8


def tile_counts(width, height):
    tx = (width + TILE_W - 1) // TILE_W
    ty = (height + TILE_H - 1) // TILE_H
    return tx, ty, tx * ty


def local_tiles(width, height, world):
    return (tile_counts(width, height)[2] + world - 1) // world


def tiles_of_rank(width, height, rank, world):
    ntiles = tile_counts(width, height)[2]
    return np.arange(rank, ntiles, world)


def pixel_owner(width, height, world):
    """[height][width] array: rank that renders each pixel."""
    tx = tile_counts(width, height)[0]
    py, px = np.mgrid[0:height, 0:width]
    return ((py // TILE_H) * tx + px // TILE_W) % world


def extract_local(frame, rank, world):
    """raster [H][W][C] -> this rank's gather-layout buffer [nlocal][8][32][C] (padding zero-filled)."""
    h, w = frame.shape[:2]
    tx = tile_counts(w, h)[0]
    nl = local_tiles(w, h, world)
    out = np.zeros((nl, TILE_H, TILE_W) + frame.shape[2:], frame.dtype)
    for j, t in enumerate(tiles_of_rank(w, h, rank, world)):
        x0, y0 = (t % tx) * TILE_W, (t // tx) * TILE_H
        blk = frame[y0:y0 + TILE_H, x0:x0 + TILE_W]
        out[j, :blk.shape[0], :blk.shape[1]] = blk
    return out


def assemble(gathered, width, height):
    """[world][nlocal][8][32][C] -> raster [H][W][C]."""
    world = gathered.shape[0]
    tx = tile_counts(width, height)[0]
    py, px = np.mgrid[0:height, 0:width]
    t = (py // TILE_H) * tx + px // TILE_W
    return gathered[t % world, t // world, py % TILE_H, px % TILE_W]
