"""CPU-side checks of the boundary: the C-ABI library builds, loads without a GPU, exports every
symbol include/frg.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from facerecognition_infrenceengine_b200 import build
    build.build()
    from facerecognition_infrenceengine_b200 import _native
    return _native


def header_symbols():
    text = open(os.path.join(ROOT, "include", "frg.h")).read()
    return sorted(set(re.findall(r"FRG_API\s+[\w\s\*]+?\b(frg_\w+)\s*\(", text)))


def test_header_and_binding_agree(native):
    syms = header_symbols()
    assert len(syms) >= 19
    assert syms == sorted(native.SIGNATURES)


def test_library_exports_every_declared_symbol(native):
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (frg_\w+)", out))
    assert set(header_symbols()) <= exported
    # nothing but the ABI leaks out of the library
    leaked = [s for s in re.findall(r" T (\S+)", out) if not s.startswith("frg_")]
    assert leaked == []


def test_struct_layouts_match_header(native):
    assert ctypes.sizeof(native.StoreStats) == 56
    assert ctypes.sizeof(native.MatchParams) == 32
    assert native.lib.frg_abi_version() == 1


def test_sass_is_sm100a(native):
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(native):
    """Without a CUDA device compute entry points fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeError):
        native.device_count()
    from facerecognition_infrenceengine_b200 import GalleryStore
    with pytest.raises(native.NativeError):
        GalleryStore(dim=512, capacity=16)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "facerecognition_infrenceengine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f
