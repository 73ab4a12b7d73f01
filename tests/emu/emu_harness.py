"""Loads the CPU *emulation* build of the kernels (tests/emu/build_emu.sh).

DEVELOPMENT HARNESS ONLY: it lets the C ABI and the kernels' index logic be
exercised in a container without a GPU.  It is not importable from the product
package and is never used as a fallback.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))

from melvin import _capi  # noqa: E402

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "_build", "libmelvin_emu.so")
        srcs = [os.path.join(ROOT, "melvin.py_b200", "csrc", f)
                for f in os.listdir(os.path.join(ROOT, "melvin.py_b200", "csrc"))]
        srcs.append(os.path.join(HERE, "emu_rt.cpp"))
        if (not os.path.exists(so)
                or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)):
            subprocess.check_call(["sh", os.path.join(HERE, "build_emu.sh")])
        _LIB = _capi.declare(ctypes.CDLL(so))
    return _LIB


def ptr(a):
    return a.ctypes.data


class EmuCtx:
    def __init__(self, nx, nz, lx, lz, fdm_z=False, fd_order=2):
        self.lib = lib()
        p = _capi.Params()
        p.nx, p.nz, p.fdm_z, p.fd_order = nx, nz, int(fdm_z), fd_order
        p.lx, p.lz = lx, lz
        p.kx0 = float(np.abs(1j * 2 * np.pi / lx))
        p.kz0 = float(np.abs(1j * 2 * np.pi / lz))
        p.d2x = float(-np.abs(1j * 2 * np.pi) ** 2 / lx ** 2)
        p.d2z = float(-np.abs(1j * 2 * np.pi) ** 2 / lz ** 2)
        h = ctypes.c_void_p()
        _capi.check(self.lib, self.lib.mlv_create(ctypes.byref(p), ctypes.byref(h)))
        self.h = h
        info = _capi.Info()
        _capi.check(self.lib, self.lib.mlv_get_info(h, ctypes.byref(info)))
        self.info = info
        self.nx, self.nz = nx, nz

    def call(self, name, *args):
        _capi.check(self.lib, getattr(self.lib, name)(self.h, *args))

    def ibuf(self):
        return np.full((self.nx, self.info.ipitch), np.nan + 1j * np.nan, dtype=np.complex128)

    def close(self):
        self.lib.mlv_destroy(self.h)


Ctx = EmuCtx
