"""CPU: the C-ABI shared library builds, loads and exports every symbol include/vxrt.h declares; without a GPU
the product fails loudly (no CPU fallback, nothing under oracle/ is ever loaded by it)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vxrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vxrt_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ("vxrt_create", "vxrt_upload_grid", "vxrt_upload_range", "vxrt_update_partial", "vxrt_edit_remove_sphere",
                 "vxrt_build_depth_field", "vxrt_set_frame", "vxrt_render", "vxrt_read_rgba8", "vxrt_read_debug", "vxrt_get_stats"):
        assert must in syms


def test_library_exports_every_declared_symbol(vx):
    lib = vx.load_library()
    for s in header_symbols():
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", vx.build.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vxrt_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported
    from voxel_rt_b200 import api
    assert set(api._SIGNATURES) == set(header_symbols())


def test_library_is_built_for_sm_100a(vx):
    out = subprocess.run(["cuobjdump", "--list-elf", vx.build.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "voxel-rt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libvxo" not in text and "oracle_lib" not in text and "vxo_" not in text, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libvxrt.so")], capture_output=True, text=True).stdout
    assert "libvxo" not in out


def test_no_gpu_means_loud_failure(vx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert vx.load_library().vxrt_device_available() == 0
    with pytest.raises(vx.VxrtError, match="no CPU fallback"):
        vx.Renderer()


def test_host_frame_needs_a_device_and_leaves_nothing_behind(vx):
    """the shared host frame is page-locked through CUDA: without a device its creation fails loudly and unlinks the name"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    name = "/vxrt_cabi_%d" % os.getpid()
    with pytest.raises(vx.VxrtError, match="cudaHostRegister"):
        vx.HostFrame(name, 64, 32, create=True)
    assert not os.path.exists("/dev/shm" + name)
    with pytest.raises(vx.VxrtError, match="shm_open"):
        vx.HostFrame(name, 64, 32, create=False)


def test_frame_struct_layout(vx):
    import ctypes as C
    import oracle_lib as ol
    assert C.sizeof(vx.Frame) == C.sizeof(ol.Frame) == 360          # 90 scalars, SURVEY.md a2
    assert [f[0] for f in vx.Frame._fields_] == [f[0] for f in ol.Frame._fields_]


def test_documents_name_only_entry_points_that_exist():
    """INTEGRATION.md / DESIGN.md / README.md show reference-side bindings: every vxrt_* name in them is declared in include/vxrt.h
    (or is one of the repo's file / binary names)"""
    text = open(os.path.join(ROOT, "include", "vxrt.h")).read()
    declared = set(re.findall(r"\b(vxrt_[a-z0-9_]+)\b", text))
    files = {"vxrt_controls", "vxrt_headless", "vxrt_render", "vxrt_glshim", "vxrt_host_frame_", "vxrt_p2p_"}
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md", os.path.join("profiles", "README.md"), os.path.join("scripts", "README.md")):
        names = set(re.findall(r"\b(vxrt_[a-z0-9_]+)\b", open(os.path.join(ROOT, doc)).read()))
        assert names <= declared | files, (doc, sorted(names - declared - files))
