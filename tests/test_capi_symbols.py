"""The C-ABI library loads and exports exactly what include/drr_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

from deepdrr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "drr_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drr_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} missing from libdrr_b200.so"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (drr_[a-z_0-9]+)$", out, flags=re.M)))
    assert exported == _declared(), "exported drr_* symbols differ from the header"


def test_version_and_sm100a_code():
    lib = _lib.load()
    assert lib.drr_version().decode().endswith("sm_100a")
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.drr_create(0, ctypes.byref(h))
    assert rc == _lib.E_CUDA
    assert b"no CPU fallback" in lib.drr_last_error(None)
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deepdrr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f
