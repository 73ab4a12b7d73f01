#!/usr/bin/env python3
"""Micro-benchmark of the three step kernels through the C ABI on random data
(no time stepping): prints ms per launch.  usage: kbench.py [nx nz] [reps]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))
from melvin import _backend, _capi  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 2 else 4096
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ctx = _backend.Context(nx, nz, 16.0 / 9.0, 1.0, False, 2)
dev = "cuda"
S = lambda: torch.randn(ctx.spec_shape, dtype=torch.complex128, device=dev)  # noqa: E731
I = lambda: torch.randn((nx, ctx.ipitch), dtype=torch.complex128, device=dev)  # noqa: E731
w, q_out, f0, fm1 = S(), S(), S(), S()
iux, iuz, iq, ia, ib = I(), I(), I(), I(), I()
red4 = torch.empty(4, dtype=torch.float64, device=dev)
vp = ctypes.c_void_p


def timed(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:16s} {a.elapsed_time(b) / reps:8.4f} ms")


srcs = (vp * 3)(w.data_ptr(), w.data_ptr(), w.data_ptr())
order = os.environ.get("KBENCH_XINV_ORDER", "api")       # "api": scalar first, as Variable.vec_dot_nabla issues it
if order == "api":
    ops = (ctypes.c_int32 * 3)(_capi.OP_IDENT, _capi.OP_UX, _capi.OP_UZ)
    dsts = (vp * 3)(iq.data_ptr(), iux.data_ptr(), iuz.data_ptr())
else:
    ops = (ctypes.c_int32 * 3)(_capi.OP_UX, _capi.OP_UZ, _capi.OP_IDENT)
    dsts = (vp * 3)(iux.data_ptr(), iuz.data_ptr(), iq.data_ptr())
timed("x_inverse(3)", lambda: ctx.call("mlv_x_inverse", 3, srcs, ops, dsts))
timed("advect_z", lambda: ctx.call("mlv_advect_z", vp(iux.data_ptr()), vp(iuz.data_ptr()), vp(iq.data_ptr()),
                                   vp(ia.data_ptr()), vp(ib.data_ptr()), vp(red4.data_ptr())))
d = _capi.XFwd()
d.nf, d.mode = 2, 1
d.src[0], d.src[1] = ia.data_ptr(), ib.data_ptr()
d.sym[0], d.sym[1] = _capi.SYM_FDX, _capi.SYM_FDZ
d.coef[0] = d.coef[1] = -1.0
d.lin = _capi.make_lin_terms([])
d.integ.ab_order, d.integ.scheme = 2, _capi.SCHEME_SI_LAP
d.integ.dt, d.integ.alpha, d.integ.lcoef = 1e-5, 0.51, 1e-5
d.integ.q_in, d.integ.q_out = w.data_ptr(), q_out.data_ptr()
d.integ.f0, d.integ.fm1 = f0.data_ptr(), fm1.data_ptr()
timed("x_forward(2)", lambda: ctx.call("mlv_x_forward", ctypes.byref(d)))
phys = torch.randn((nx, nz), dtype=torch.float64, device=dev)
timed("z_inverse", lambda: ctx.call("mlv_z_inverse", vp(iq.data_ptr()), vp(phys.data_ptr())))
timed("z_forward", lambda: ctx.call("mlv_z_forward", vp(phys.data_ptr()), vp(ia.data_ptr())))
a_, b_ = torch.empty(1 << 28, dtype=torch.float64, device=dev), torch.empty(1 << 28, dtype=torch.float64, device=dev)
timed("torch copy 2GiB", lambda: b_.copy_(a_))
