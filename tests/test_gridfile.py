"""CPU: the VXRTGRD1 grid-file format on the host side (voxel-rt_b200/gridfile.py).  The device side
(vxrt_save_grid / vxrt_load_grid) is checked against it in tests/test_gpu_parity.py."""
import struct

import numpy as np
import pytest


def small_grid(seed=0, dims=(12, 5, 7)):
    rng = np.random.default_rng(seed)
    v = rng.integers(-1, 1 << 24, size=dims[0] * dims[1] * dims[2], dtype=np.int32)
    v[rng.random(v.size) < 0.5] = np.float32(-3.5).view(np.int32)     # depth-field style negative float bits
    return v, dims


def test_round_trip_and_header_layout(vx, tmp_path):
    v, dims = small_grid()
    p = str(tmp_path / "g.vxg")
    vx.gridfile.write_grid(p, v, dims)
    raw = open(p, "rb").read()
    assert len(raw) == 64 + 4 * v.size
    assert raw[:8] == b"VXRTGRD1"
    w, h, d, flags = struct.unpack_from("<IIII", raw, 8)
    count, fnv = struct.unpack_from("<QQ", raw, 24)
    assert (w, h, d, flags, count) == (*dims, 0, v.size)
    assert fnv == vx.scenes.fnv1a64(v)
    assert raw[40:64] == b"\0" * 24
    assert np.array_equal(np.frombuffer(raw, "<i4", offset=64), v)
    got, gdims = vx.gridfile.read_grid(p)
    assert gdims == dims and got.dtype == np.int32 and np.array_equal(got, v)
    assert vx.gridfile.read_header(p)["fnv"] == fnv


def test_default_level_fingerprint_survives_the_file(vx, oracle, default_level, tmp_path):
    p = str(tmp_path / "level.vxg")
    vx.gridfile.write_grid(p, default_level, (512, 96, 512))
    assert vx.gridfile.read_header(p)["fnv"] == 0x4c58cc4001a22afa
    got, _ = vx.gridfile.read_grid(p)
    assert oracle.fnv(got) == 0x4c58cc4001a22afa


def test_rejects_damaged_files(vx, tmp_path):
    v, dims = small_grid(1)
    p = str(tmp_path / "g.vxg")
    vx.gridfile.write_grid(p, v, dims)
    raw = bytearray(open(p, "rb").read())
    E = vx.gridfile.GridFileError

    def variant(name, data):
        q = str(tmp_path / name)
        open(q, "wb").write(bytes(data))
        return q
    with pytest.raises(E, match="VXRTGRD1"):
        vx.gridfile.read_grid(variant("magic", b"VXRTGRD2" + raw[8:]))
    with pytest.raises(E, match="header"):
        vx.gridfile.read_grid(variant("short", raw[:40]))
    with pytest.raises(E, match="payload has"):
        vx.gridfile.read_grid(variant("trunc", raw[:-8]))
    flipped = bytearray(raw)
    flipped[64 + 17] ^= 0x40
    with pytest.raises(E, match="fingerprint"):
        vx.gridfile.read_grid(variant("flip", flipped))
    assert vx.gridfile.read_grid(variant("flip2", flipped), verify=False)[0].size == v.size
    bad_count = bytearray(raw)
    struct.pack_into("<Q", bad_count, 24, v.size + 1)
    with pytest.raises(E, match="count"):
        vx.gridfile.read_grid(variant("count", bad_count))
    with pytest.raises(E, match="extents"):
        vx.gridfile.write_grid(str(tmp_path / "x"), v[:-1], dims)
