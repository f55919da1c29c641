"""Host-side (numpy / pure Python) statement of the synthetic C4 terrain that libvxrt generates on the device
(include/vxrt.h vxrt_generate_terrain): integer-only, so it must match bit for bit.  Test infrastructure."""
import numpy as np

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def lattice16(seed, ix, iz, octv):
    k = seed ^ ((ix & 0xFFFFFFFF) * 73856093) ^ ((iz & 0xFFFFFFFF) * 19349663) ^ (octv * 83492791)
    return splitmix64(k & M64) >> 48


def fbm16(seed, x, z):
    acc = 0
    for o in range(5):
        S = 256 >> o
        ix, iz, fx, fz = x // S, z // S, x % S, z % S
        v00, v10 = lattice16(seed, ix, iz, o), lattice16(seed, ix + 1, iz, o)
        v01, v11 = lattice16(seed, ix, iz + 1, o), lattice16(seed, ix + 1, iz + 1, o)
        top = v00 * (S - fx) + v10 * fx
        bot = v01 * (S - fx) + v11 * fx
        acc += ((top * (S - fz) + bot * fz) // (S * S)) >> (o + 1)
    return min(acc, 65535)


def height(seed, x, z, h):
    return h // 4 + ((fbm16(seed, x, z) * (h * 3 // 8)) >> 16)


def crem(a, b):
    """C remainder (sign of the dividend)"""
    return int(np.fmod(a, b))


def generate(dims, seed):
    w, h, d = dims
    surf = np.array([[height(seed, x, z, h) for x in range(w)] for z in range(d)], np.int32)      # [z][x]
    z, y, x = np.mgrid[0:d, 0:h, 0:w]
    cv = 5 * ((x + y + z) % 3)
    s = surf[:, None, :]
    vox = np.full((d, h, w), -1, np.int64)
    stone = ((90 + cv) << 16) | ((90 + cv) << 8) | (90 + cv)
    dirt = ((120 + cv) << 16) | ((100 + cv) << 8)
    grass = (10 << 16) | ((130 + cv) << 8) | 10
    vox = np.where(y <= s - 11, stone, np.where(y <= s - 3, dirt, np.where(y <= s, grass, -1)))
    vox = vox.astype(np.int32)

    def place(px, py, pz, v):
        if 0 <= px < w and 0 <= py < h and 0 <= pz < d:
            vox[pz, py, px] = v
    for tz in range(10, d - 10):
        for tx in range(10, w - 10):
            if tx % 30 == 0 and tz % 25 == 0:
                by = int(surf[tz, tx])
                px, py, pz = tx + 1 + tz % 7, by, tz                       # trunk, level.cpp:59-79
                for rz in range(-2, 1):
                    for ry in range(0, 6):
                        for rx in range(-2, 1):
                            if rx + px < w and ry + py < h and rz + pz < d:
                                m = crem(rx + rz, 2) * 10
                                place(rx + px, ry + py, rz + pz, ((128 - m) << 16) | ((100 - m) << 8) | 15)
                px, py, pz = tx + tz % 7, by + 10, tz                      # bush, level.cpp:4-27
                for rz in range(-6, 6):
                    for ry in range(-6, 6):
                        for rx in range(-6, 6):
                            if (rx + px < w and ry + py < h and rz + pz < d and rx + px >= 0 and ry + py >= 0 and rz + pz >= 0
                                    and rx * rx + ry * ry + rz * rz < 36):
                                place(rx + px, ry + py, rz + pz, (15 << 16) | ((128 - crem(rx + ry + rz, 3) * 20) << 8) | 15)
    return np.ascontiguousarray(vox.ravel())
