"""CPU-side checks of the C-ABI boundary: the CUDA library builds for sm_100a,
loads, exports every symbol include/melvin_b200.h declares, and the Python layer
refuses to run without a device (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__ as ge
    if ge._stale():
        if shutil.which("nvcc") is None:
            pytest.skip("nvcc not available and libmelvin_b200.so not prebuilt")
        ge.build()
    return ge.LIB


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "melvin_b200.h")).read()
    return sorted(set(re.findall(r"\b(mlv_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_mirror_agree():
    from melvin import _capi
    assert sorted(_capi.EXPORTS) == declared_symbols()


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym
    lib.mlv_abi_version.restype = ctypes.c_int
    from melvin import _capi
    assert lib.mlv_abi_version() == _capi.ABI_VERSION
    lib.mlv_last_error.restype = ctypes.c_char_p
    assert lib.mlv_last_error() is not None


def test_library_contains_sm100a_code(libpath):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    code = (
        "import sys; sys.path.insert(0, %r);"
        "import numpy as np; import melvin; from melvin import b200 as xp\n"
        "try:\n    xp.zeros((4, 4))\n    print('ALLOCATED')\n"
        "except Exception as e:\n    print(type(e).__name__)\n"
        "try:\n    melvin.ArrayFactory(None, np)\n    print('ACCEPTED')\n"
        "except Exception as e:\n    print(type(e).__name__)\n" % os.path.join(ROOT, "melvin.py_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout.split()
    assert out == ["BackendUnavailable", "BackendUnavailable"], out
