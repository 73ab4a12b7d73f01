"""Process-wide binding to libmelvin_b200.so and to the device allocator.

torch is used for exactly three things: device memory (tensors own every field
buffer), the CUDA stream handle and host<->device copies.  All arithmetic on
those buffers goes through the C ABI in ``include/melvin_b200.h``.

There is NO CPU path: if the shared library or a CUDA device is missing the
first operation raises.  (``_install`` lets the test-suite point the binding at
the host *emulation* build of the kernels; product code never does that.)
"""
import ctypes
import os
import weakref

import numpy as np
import torch

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
# MLV_LIB: development switch to A/B-measure another build of the same sources (tools/build_variant.sh)
LIB_PATH = os.environ.get("MLV_LIB") or os.path.join(_HERE, "_lib", "libmelvin_b200.so")

_state = {"lib": None, "device": None, "launches": 0, "calls": {}}


class BackendUnavailable(RuntimeError):
    pass


def _install(lib, device):
    """Bind an already loaded library handle and a torch device (tests only)."""
    _state["lib"] = _capi.declare(lib)
    _state["device"] = torch.device(device)


def _load_default():
    if not os.path.exists(LIB_PATH):
        raise BackendUnavailable(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). melvin-b200 has no CPU fallback.")
    if not torch.cuda.is_available():
        raise BackendUnavailable("melvin-b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _install(ctypes.CDLL(LIB_PATH), "cuda")


def lib():
    if _state["lib"] is None:
        _load_default()
    return _state["lib"]


def device():
    if _state["device"] is None:
        _load_default()
    return _state["device"]


def is_cuda():
    return device().type == "cuda"


def count_launch(n=1):
    _state["launches"] += n


def launches():
    """Number of kernels the library has launched since process start
    (counted inside libmelvin_b200.so at every launch site)."""
    return int(lib().mlv_launch_count())


def call_counts():
    """Per-entry-point call counts (copy)."""
    return dict(_state["calls"])


def current_stream_handle():
    if is_cuda():
        return torch.cuda.current_stream().cuda_stream
    return 0


# --------------------------------------------------------------- allocation
_TORCH_DTYPE = {np.dtype(np.float64): torch.float64, np.dtype(np.complex128): torch.complex128,
                np.dtype(np.int64): torch.int64}


def torch_dtype(dtype):
    dt = np.dtype(dtype)
    if dt not in _TORCH_DTYPE:
        raise TypeError(f"melvin-b200 arrays are float64 / complex128 (got {dt}); "
                        "single precision is not implemented")
    return _TORCH_DTYPE[dt]


def empty(shape, dtype):
    return torch.empty(tuple(shape), dtype=torch_dtype(dtype), device=device())


def zeros(shape, dtype):
    """Zero-filled device buffer; the fill is the library's own kernel (mlv_elementwise)."""
    t = empty(shape, dtype)
    if t.numel():
        if t.is_complex() or t.is_floating_point():
            from . import b200
            b200._elementwise(_capi.EW_COPY, 0.0, None, out=t)
        else:
            t.zero_()
    return t


def _host_tensor(arr, dtype=None):
    a = np.ascontiguousarray(arr, dtype=dtype)
    if a.dtype not in _TORCH_DTYPE:
        a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    return torch.from_numpy(a)


def from_host(arr, dtype=None):
    """Host array -> new device buffer.  A source in PINNED host memory is copied
    asynchronously on the current stream (the caller keeps the source unchanged until the
    stream has passed the copy, as with any pinned-memory transfer); pageable sources are
    copied synchronously."""
    t = _host_tensor(arr, dtype)
    if is_cuda():
        return t.to(device(), non_blocking=t.is_pinned())
    return t.clone()


def copy_from_host(dst, arr):
    """Host array -> existing device buffer of the same shape (no staging buffer); asynchronous
    for pinned sources, see from_host."""
    t = _host_tensor(arr, np.complex128 if dst.is_complex() else np.float64)
    dst.copy_(t, non_blocking=is_cuda() and t.is_pinned())


def copy_to_host(out, t):
    """Device buffer -> caller-provided host array (numpy array or torch tensor).  Into PINNED
    memory the copy is asynchronous on the current stream: the data is valid once the stream
    (or an event recorded after this call) has been synchronised."""
    h = out if isinstance(out, torch.Tensor) else torch.from_numpy(out)
    if tuple(h.shape) != tuple(t.shape) or h.dtype != t.dtype or not h.is_contiguous():
        raise ValueError("on_host(out=...): the host buffer must be contiguous and match the shape and dtype")
    h.copy_(t.detach(), non_blocking=is_cuda() and h.is_pinned())
    return out


def to_host(t):
    return t.detach().cpu().numpy() if t.device.type != "cpu" else t.detach().numpy().copy()


def synchronize():
    if is_cuda():
        torch.cuda.synchronize()


# ------------------------------------------------------------------ context
class Context:
    """One mlv_ctx per Parameters object (plans for one grid)."""

    def __init__(self, nx, nz, lx, lz, fdm_z, fd_order, shard=None):
        self.lib = lib()
        p = _capi.Params()
        p.nx, p.nz, p.fdm_z, p.fd_order = int(nx), int(nz), int(bool(fdm_z)), int(fd_order)
        p.lx, p.lz = float(lx), float(lz)
        # reference: melvin/BasisFunctions.py:26-59 (COMPLEX_EXP factors)
        p.kx0 = float(np.abs(1j * 2 * np.pi / lx))
        p.kz0 = float(np.abs(1j * 2 * np.pi / lz))
        p.d2x = float(-np.abs(1j * 2 * np.pi) ** 2 / lx ** 2)
        p.d2z = float(-np.abs(1j * 2 * np.pi) ** 2 / lz ** 2)
        h = ctypes.c_void_p()
        _capi.check(self.lib, self.lib.mlv_create(ctypes.byref(p), ctypes.byref(h)))
        self.handle = h
        # slab decomposition behind the public API (melvin/_dist.py): local shapes from here on
        self.rank, self.world = shard if shard else (0, 1)
        if self.world > 1:
            _capi.check(self.lib, self.lib.mlv_set_sharding(h, self.rank, self.world, 1, 1))
        info = _capi.Info()
        _capi.check(self.lib, self.lib.mlv_get_info(h, ctypes.byref(info)))
        self.nn, self.nm = info.nn, info.nm
        self.spec_shape = (info.spec_rows, info.spec_cols)           # local (rows, columns of this rank)
        self.global_spec_shape = (info.spec_rows, info.nm if self.world > 1 else info.spec_cols)
        self.ipitch = info.ipitch
        self.nm_local = info.nm_local
        self.m_off = self.rank * info.spec_cols if self.world > 1 else 0
        self.nxl = int(nx) // self.world
        self.x_off = self.rank * self.nxl
        self.phys_shape = (self.nxl, int(nz))                        # local rows
        self.ifield = int(info.ibytes) // 16                         # elements of one exchange field
        self.red_doubles = int(info.red_doubles)
        self.nx, self.nz = int(nx), int(nz)
        self.fdm_z = bool(fdm_z)
        self.fd_order = int(fd_order)
        self.dx, self.dz = lx / nx, lz / nz
        self._ipool = []
        self._scratch_i = None
        self._red4 = None
        self._lazy_exprs = weakref.WeakSet()     # live deferred expressions (b200.SpecExpr)
        # CFL / kinetic-energy partials of the fused z stage: Simulation.end_loop clears this for
        # steps no ticker will read (melvin/simulation.py); readers fall back to explicit reductions
        self.want_reductions = True
        self._reductions_on = True
        self._stream = None
        self.set_stream(current_stream_handle())

    def set_stream(self, handle):
        _capi.check(self.lib, self.lib.mlv_set_stream(self.handle, ctypes.c_void_p(handle)))
        self._stream = handle

    def call(self, name, *args, count=True):
        # kernels run on torch's *current* stream, like the allocations and copies around them
        stream = current_stream_handle()
        if stream != self._stream:
            self.set_stream(stream)
        _capi.check(self.lib, getattr(self.lib, name)(self.handle, *args))
        if count:
            _state["launches"] += 1
            calls = _state["calls"]
            calls[name] = calls.get(name, 0) + 1

    def sync_reduction_mode(self):
        """Tell the library whether the next fused z stage computes its reduction partials."""
        on = bool(self.want_reductions)
        if on != self._reductions_on:
            _capi.check(self.lib, self.lib.mlv_set_reductions(self.handle, int(on)))
            self._reductions_on = on
        return on

    # pool of x-transformed intermediates: (nx, ipitch) complex128, FDM-z mode: (nn, nz) x spectra
    def take_i(self):
        if self._ipool:
            return self._ipool.pop()
        if self.world > 1:
            return empty((self.ifield,), np.complex128)
        return empty(self.spec_shape if self.fdm_z else (self.nx, self.ipitch), np.complex128)

    # ---- slab decomposition: the all-to-all between the two passes of a transform
    def exchange(self, send, forward):
        """send: field in [peer][block] layout -> the blocks received from the peers.  Inverse
        blocks are (nx/G rows, nml columns), forward blocks (column tiles of a rank, nx/G rows)."""
        from . import _dist
        n = self.world * self.nxl * self.ipitch if not forward else self._fwd_len()
        recv = self.take_i()
        _dist.all_to_all(recv[:n], send[:n])
        return recv

    def _fwd_len(self):
        tiles = self.ifield // (self.world * self.nxl)               # tpr * ct
        return self.world * tiles * self.nxl

    def give_i(self, t):
        if t is not None and len(self._ipool) < 8:
            self._ipool.append(t)

    def scratch_i(self):
        if self._scratch_i is None:
            self._scratch_i = (empty((self.ifield,), np.complex128) if self.world > 1
                               else empty((self.nx, max(self.ipitch, 1)), np.complex128))
        return self._scratch_i

    def __del__(self):
        try:
            if self.handle:
                self.lib.mlv_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_contexts = {}


def context_for(params):
    """Shared context for a Parameters-like object (nx, nz, lx, lz, discretisation, order)."""
    fdm_z = params.discretisation[1] == "fdm"
    if params.discretisation[0] == "fdm":
        raise NotImplementedError("Finite difference not implemented in x direction")
    if getattr(params, "precision", "double") not in ("double", "single"):
        raise NotImplementedError("precision must be 'double' or 'single' (promoted to double)")
    from . import _dist
    shard = None
    if _dist.world() > 1:
        if fdm_z:
            raise NotImplementedError(
                "Fourier-x / FDM-z runs are not slab-decomposed (replicas only): set MLV_SHARD=0")
        if int(params.nx) % (2 * _dist.world()):
            raise NotImplementedError("slab decomposition needs nx divisible by twice the rank count")
        shard = (_dist.rank(), _dist.world())
    # one context (plans + scratch pool) per grid AND per stream it is created under: simulations
    # built under different `torch.cuda.stream(...)` blocks run concurrently (melvin/ensemble.py)
    key = (int(params.nx), int(params.nz), float(params.lx), float(params.lz), fdm_z,
           int(params.spatial_derivative_order), str(device()), shard, current_stream_handle())
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = Context(params.nx, params.nz, params.lx, params.lz, fdm_z,
                      params.spatial_derivative_order, shard)
        _contexts[key] = ctx
    return ctx
