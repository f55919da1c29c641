"""VXRTGRD1 grid files on the host (the same format vxrt_save_grid / vxrt_load_grid stream to and from the device,
include/vxrt.h): 64-byte little-endian header + int32 voxels in the reference's order (x + w*y + w*h*z).  The
reference has no on-disk format (its level exists only as level.cpp's generator); this one exists so that benchmark
inputs and edited levels can be reproduced exactly.  No device is needed; the payload fingerprint is
computed by the FNV helper that libvxrt.so exports (host code; importing it loads -- and, if its sources are newer, builds -- the
library, which itself needs no GPU to load)."""
import struct

import numpy as np

from .scenes import fnv1a64

MAGIC = b"VXRTGRD1"
HEADER = struct.Struct("<8sIIIIQQ24x")        # magic, w, h, d, flags, count, fnv


class GridFileError(ValueError):
    pass


def write_grid(path, voxels, dims):
    w, h, d = (int(v) for v in dims)
    v = np.ascontiguousarray(voxels, dtype="<i4").ravel()
    if v.size != w * h * d:
        raise GridFileError("voxel count %d does not match extents %dx%dx%d" % (v.size, w, h, d))
    with open(path, "wb") as f:
        f.write(HEADER.pack(MAGIC, w, h, d, 0, v.size, fnv1a64(v)))
        f.write(v.tobytes())


def read_header(path):
    with open(path, "rb") as f:
        raw = f.read(HEADER.size)
    if len(raw) != HEADER.size:
        raise GridFileError("file shorter than the 64-byte header")
    magic, w, h, d, flags, count, fnv = HEADER.unpack(raw)
    if magic != MAGIC:
        raise GridFileError("not a VXRTGRD1 file")
    if count != w * h * d:
        raise GridFileError("header count %d does not match extents %dx%dx%d" % (count, w, h, d))
    return dict(dims=(w, h, d), flags=flags, count=count, fnv=fnv)


def read_grid(path, verify=True):
    """-> (int32 voxels, (w, h, d)); verify=True recomputes the FNV-1a-64 fingerprint of the payload"""
    hd = read_header(path)
    v = np.fromfile(path, dtype="<i4", offset=HEADER.size)
    if v.size != hd["count"]:
        raise GridFileError("payload has %d voxels, header says %d" % (v.size, hd["count"]))
    if verify and fnv1a64(v) != hd["fnv"]:
        raise GridFileError("payload fingerprint does not match the header")
    return v.astype(np.int32, copy=False), hd["dims"]
