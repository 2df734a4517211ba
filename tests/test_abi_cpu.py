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
    assert ctypes.sizeof(native.Exchange) == 32
    assert ctypes.sizeof(native.ExchangeStatus) == 40
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


C_CONSUMER = r'''
#include <stdio.h>
#include <string.h>
#include "frg.h"

int main(void) {
  frg_match_params_t p;
  frg_exchange_t x;
  frg_store* s = NULL;
  int32_t n = -1;
  int rc;
  memset(&p, 0, sizeof p);
  memset(&x, 0, sizeof x);
  printf("abi=%d params=%u stats=%u exchange=%u xstatus=%u\n", frg_abi_version(), (unsigned)sizeof p,
         (unsigned)sizeof(frg_store_stats_t), (unsigned)sizeof x, (unsigned)sizeof(frg_exchange_status_t));
  rc = frg_device_count(&n);                 /* no GPU here: an error code and a message, not a crash */
  printf("device_count rc=%d n=%d err=%s\n", rc, (int)n, frg_last_error());
  rc = frg_store_create(0, 512, 16, FRG_STORE_BF16_PLANE, &s);
  printf("store_create rc=%d\n", rc);
  if (rc == FRG_OK) frg_store_destroy(s);
  return 0;
}
'''


def test_header_is_plain_c_and_links_from_c(native, tmp_path):
    """include/frg.h is the boundary a non-Python host would bind: it must compile as strict C99 and a
    C program must link against libfrg.so alone (no Python, no torch, no libcuda at link time)."""
    src = tmp_path / "consumer.c"
    src.write_text(C_CONSUMER)
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(native.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", libdir, "-lfrg", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "abi=1 params=32 stats=56 exchange=32 xstatus=40" in out.stdout
    import torch
    if not torch.cuda.is_available():
        assert "device_count rc=2" in out.stdout and "store_create rc=0" not in out.stdout
