"""CPU: the C-ABI libraries load, export every symbol the headers declare, and the ctypes mirrors
match the compiled struct sizes. No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from slr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"SLRGPU_API[^;{]*?\b(" + prefix + r"\w+)\s*\(", text)))


def test_gpu_library_exports_every_declared_symbol():
    syms = declared_symbols("slrgpu.h", "slrgpu_")
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(capi.gpu, s)]
    assert not missing, f"libslrgpu.so lacks {missing}"


def test_host_library_exports_every_declared_symbol():
    syms = declared_symbols("slrhost.h", "slrhost_")
    assert len(syms) >= 8
    missing = [s for s in syms if not hasattr(capi.host, s)]
    assert not missing, f"libslrhost.so lacks {missing}"


def test_struct_sizes_match():
    assert capi.check_abi()
    assert C.sizeof(capi.BvhNode) == 128 and C.sizeof(capi.LeafRecord) == 48
    assert C.sizeof(capi.Triangle) == 32 and C.sizeof(capi.Vertex) == 48


def test_abi_version_and_device_count_do_not_need_a_gpu():
    assert capi.gpu.slrgpu_abi_version() >> 16 == 1
    assert capi.gpu.slrgpu_device_count() >= 0


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    if capi.gpu.slrgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    from slr_b200 import synth
    b = capi.SceneBuilder()
    pos, idx = synth.cube()
    b.place_mesh(b.add_mesh(pos, idx))
    hs = b.finish()
    with pytest.raises(capi.SlrError, match="no CUDA device|failed"):
        capi.GpuScene(hs)


def test_scene_create_rejects_bad_arguments():
    out = C.c_void_p()
    assert capi.gpu.slrgpu_scene_create(None, 0, C.byref(out)) == -1
    d = capi.SceneDesc()
    d.struct_size = 12
    assert capi.gpu.slrgpu_scene_create(C.byref(d), 0, C.byref(out)) == -1
    assert b"struct_size" in capi.gpu.slrgpu_last_error()
