"""Harness that drives libmelvin_b200.so on a CUDA device with NumPy test data.

``ptr(a)`` uploads the NumPy array ``a`` into a mirrored device tensor and returns
its device pointer; after every ``Ctx.call`` all mirrors are copied back into
their NumPy arrays (synchronising), so the shared cases in abi_cases.py read the
same on the emulation build and on the GPU."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))

from melvin import _backend, _capi  # noqa: E402

_mirrors = {}


def ptr(a):
    a_c = np.ascontiguousarray(a)
    old = _mirrors.get(id(a))
    if old is not None and old[0] is a and tuple(old[1].shape) == a_c.shape:
        old[1].copy_(torch.from_numpy(a_c))      # keep the device pointer stable
        return old[1].data_ptr()
    t = torch.from_numpy(a_c).cuda()
    _mirrors[id(a)] = (a, t)
    return t.data_ptr()


def _download():
    torch.cuda.synchronize()
    for a, t in _mirrors.values():
        if a.flags.writeable:
            a[...] = t.cpu().numpy()


class Ctx:
    def __init__(self, nx, nz, lx, lz, fdm_z=False, fd_order=2):
        self.lib = _backend.lib()
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
        _capi.check(self.lib, self.lib.mlv_set_stream(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def call(self, name, *args):
        _capi.check(self.lib, getattr(self.lib, name)(self.h, *args))
        _download()

    def ibuf(self):
        return np.full((self.nx, self.info.ipitch), np.nan + 1j * np.nan, dtype=np.complex128)

    def close(self):
        torch.cuda.synchronize()
        self.lib.mlv_destroy(self.h)
        _mirrors.clear()
