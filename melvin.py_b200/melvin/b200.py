"""``xp`` array namespace of the B200 backend.

The reference injects an array module (``numpy`` or ``cupy``) into every class
(e.g. reference ``melvin/Simulation.py:19-21``).  This module is the
drop-in for that slot: the subset of the NumPy surface the reference and its
example scripts touch, implemented on device buffers through the C ABI
(``mlv_elementwise``, ``mlv_reduce``, ``mlv_spec_lincomb`` ...).

Three array-like types live here:

``DeviceArray``  eager float64 / complex128 array on the GPU (views, slicing,
                 arithmetic, reductions).
``SpecExpr``     a *deferred* linear combination of spectral arrays and
                 nonlinear-term handles, ``sum_i c_i op_i(a_i) + sum_k c_k N_k``.
                 The example scripts build their right-hand sides with ``-``,
                 ``+``, ``c *`` and ``/ c``; keeping them symbolic lets
                 ``Integrator.integrate`` evaluate the whole RHS, the history
                 write and the update in one kernel (the forward-x epilogue).
``LazyLap``      ``coef * lap`` (``Variable.lap()``) kept symbolic for the same
                 reason; materialises to a real DeviceArray on any other use.
"""
import builtins as _bi
import ctypes
import numbers

import numpy as np

from . import _backend, _capi

float64 = np.float64
complex128 = np.complex128
int64 = np.int64
pi = np.pi

_util_ctx = None


def _ctx():
    """Context used for shape-agnostic array kernels (stream + reduction scratch)."""
    global _util_ctx
    if _util_ctx is None:
        _util_ctx = _backend.Context(16, 16, 1.0, 1.0, False, 2)
    return _util_ctx


def _is_scalar(x):
    return isinstance(x, (numbers.Number, np.generic))


def _is_complex_scalar(x):
    return isinstance(x, (complex, np.complexfloating))


# ================================================================== slabs
class Dist:
    """One axis of an array is distributed over the ranks (melvin/_dist.py): the array
    stands for the GLOBAL index range [start, start + count) of that axis -- spectral columns m
    ("cols") or physical rows x ("rows") -- and the local tensor holds the part of it this rank
    owns.  Indexing uses global indices; host reads gather."""

    def __init__(self, ctx, kind, axis, start, count, squeeze=False):
        self.ctx, self.kind, self.axis, self.start, self.count, self.squeeze = ctx, kind, axis, start, count, squeeze

    def owned(self):
        c = self.ctx
        return (c.m_off, c.nm_local) if self.kind == "cols" else (c.x_off, c.nxl)

    def local_range(self):
        own0, ownn = self.owned()
        lo = _bi.max(self.start, own0)
        return lo, _bi.max(lo, _bi.min(self.start + self.count, own0 + ownn))

    def global_size(self):
        return self.ctx.nm if self.kind == "cols" else self.ctx.nx

    def root(self, axis):
        return Dist(self.ctx, self.kind, axis, 0, self.global_size())


def _expand_index(idx, ndim):
    idx = idx if isinstance(idx, tuple) else (idx,)
    if any(i is Ellipsis for i in idx):
        k = idx.index(Ellipsis)
        idx = idx[:k] + (slice(None),) * (ndim - (len(idx) - 1)) + idx[k + 1:]
    return tuple(idx) + (slice(None),) * (ndim - len(idx))


def _dist_index(d, idx, ndim):
    """Global index of an array distributed by `d` -> (local index, Dist of the result or None,
    (lo, hi) position of the local part inside the indexed global range)."""
    idx = _expand_index(idx, ndim)
    if len(idx) != ndim or any(not isinstance(i, (int, np.integer, slice)) for i in idx):
        raise NotImplementedError("slab-decomposed arrays support basic (int / slice) indexing only")
    comp = idx[d.axis]
    L0, L1 = d.local_range()
    squeeze = False
    if isinstance(comp, (int, np.integer)):
        g = int(comp) + (d.count if comp < 0 else 0)
        if not 0 <= g < d.count:
            raise IndexError("index out of range")
        a, b, squeeze = g, g + 1, True
    else:
        a, b, step = comp.indices(d.count)
        if step != 1:
            raise NotImplementedError("slab-decomposed axis: unit-stride slices only")
        b = _bi.max(a, b)
    lo, hi = _bi.max(d.start + a, L0), _bi.min(d.start + b, L1)
    hi = _bi.max(lo, hi)
    local = list(idx)
    local[d.axis] = slice(lo - L0, hi - L0)
    new_axis = d.axis - _bi.sum(1 for i in idx[:d.axis] if isinstance(i, (int, np.integer)))
    nd = Dist(d.ctx, d.kind, new_axis, d.start + a, b - a, squeeze or d.squeeze)
    return tuple(local), nd, (lo - (d.start + a), hi - (d.start + a))


# ============================================================= DeviceArray
class DeviceArray:
    __array_priority__ = 1000
    __array_ufunc__ = None          # make NumPy scalars defer to our reflected operators

    def __init__(self, tensor, base=None, idx=None, dist=None, gidx=None):
        self._t = tensor
        # a view keeps its parent: reads and in-place writes through the view go through the
        # parent's hooks (deferred definitions, dependants), and the view follows the parent
        # when a Variable re-points its state buffer (double buffering)
        self._base = base
        self._idx = idx
        self._dist = dist            # slab decomposition of one axis (Dist) or None
        self._gidx = gidx            # view of a slab-decomposed array: its global index

    # -- hooks for lazily materialised subclasses
    def _touch(self):
        """Called before the data is read or written in place."""
        if self._base is not None:
            self._t = self._base._touch()._t[self._idx]
        return self

    def _pre_write(self):
        """Called before the data is modified in place."""
        if self._base is not None:
            self._base._pre_write()

    # -- metadata
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def ndim(self):
        return self._t.dim()

    @property
    def size(self):
        return self._t.numel()

    @property
    def dtype(self):
        return np.dtype(np.complex128) if self._t.is_complex() else (
            np.dtype(np.float64) if self._t.is_floating_point() else np.dtype(np.int64))

    @property
    def is_complex(self):
        return self._t.is_complex()

    def __len__(self):
        return self.shape[0]

    # -- host interop
    def get(self):
        self._touch()
        if self._dist is not None:
            return self._gather()
        return _backend.to_host(self._t)

    def _gather(self):
        """Host copy of the GLOBAL array (collective: every rank calls it)."""
        from . import _dist
        if self._base is not None and self._gidx is not None:
            full = self._base.get()
            return full[self._gidx]
        d = self._dist
        full = _dist.gather(self._t, d.axis)
        if d.kind == "cols":                       # drop the padding columns of the last slab
            full = np.take(full, range(d.ctx.nm), axis=d.axis)
        if not (d.start == 0 and d.count == d.global_size()):
            full = np.take(full, range(d.start, d.start + d.count), axis=d.axis)
        return full

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a.astype(dtype) if dtype is not None else a

    def item(self):
        return self.get().reshape(-1)[0].item()

    def __float__(self):
        return float(self.item())

    def __complex__(self):
        return complex(self.item())

    def __repr__(self):
        return f"DeviceArray({self.get()!r})"

    def copy(self):
        out = DeviceArray(_backend.empty(self.shape, self.dtype), dist=self._dist)
        DeviceArray(out._t)[...] = DeviceArray(self._touch()._t)
        return out

    def take(self, indices, axis=None):
        """numpy.ndarray.take (the reference's stencils use it for the periodic wrap,
        SpatialDifferentiator.py:145-180; here a convenience of the namespace, gathered on the host)"""
        return DeviceArray(_backend.from_host(np.take(self.get(), indices, axis=axis)))

    # -- views
    def __getitem__(self, idx):
        self._touch()
        if self._dist is not None:
            local, nd, _ = _dist_index(self._dist, idx, self._t.dim())
            return DeviceArray(self._t[local], self, local, dist=nd, gidx=idx)
        tracked = self._base is not None or type(self) is not DeviceArray
        return DeviceArray(self._t[idx], self if tracked else None, idx if tracked else None)

    def __setitem__(self, idx, value):
        self._touch()
        self._pre_write()
        if isinstance(value, (SpecExpr, LazyLap)):
            value = value.materialize()
        if self._dist is not None:
            local, nd, (lo, hi) = _dist_index(self._dist, idx, self._t.dim())
            target = self._t[local]
            if isinstance(value, np.ndarray) and value.ndim == target.dim():
                # a global-shaped host block: keep the part this rank owns
                value = np.take(value, range(lo, hi), axis=nd.axis) if value.shape[nd.axis] != 1 else value
            if isinstance(value, DeviceArray) and value._dist is None and value._t.dim() == target.dim() \
                    and value.shape[nd.axis] not in (1, target.shape[nd.axis]):
                value = DeviceArray(value._touch()._t.narrow(nd.axis, lo, hi - lo))
        else:
            target = self._t[idx]
        if isinstance(value, np.ndarray):
            value = DeviceArray(_backend.from_host(value))
        _elementwise(_capi.EW_COPY, value, None, out=target)

    def reshape(self, *shape):
        self._touch()
        return DeviceArray(self._t.reshape(*shape))

    # -- arithmetic (eager)
    def _bin(self, op, other, reflected=False):
        if isinstance(other, (SpecExpr, LazyLap)):
            return NotImplemented
        self._touch()
        if isinstance(other, np.ndarray):
            other = DeviceArray(_backend.from_host(other))
        if not (isinstance(other, DeviceArray) or _is_scalar(other)):
            return NotImplemented
        a, b = (other, self) if reflected else (self, other)
        dist = self._dist if self._dist is not None else getattr(other, "_dist", None)
        if dist is not None and (self._base is not None or dist.squeeze):
            dist = None if dist.squeeze else Dist(dist.ctx, dist.kind, dist.axis, dist.start, dist.count)
        return DeviceArray(_elementwise(op, a, b), dist=dist)

    def __add__(self, o): return self._bin(_capi.EW_ADD, o)
    def __radd__(self, o): return self._bin(_capi.EW_ADD, o, True)
    def __sub__(self, o): return self._bin(_capi.EW_SUB, o)
    def __rsub__(self, o): return self._bin(_capi.EW_SUB, o, True)
    def __mul__(self, o): return self._bin(_capi.EW_MUL, o)
    def __rmul__(self, o): return self._bin(_capi.EW_MUL, o, True)
    def __truediv__(self, o): return self._bin(_capi.EW_DIV, o)
    def __rtruediv__(self, o): return self._bin(_capi.EW_DIV, o, True)
    def __neg__(self): return self._bin(_capi.EW_MUL, -1.0)
    def __pos__(self): return self

    def __pow__(self, e):
        if self.is_complex or not _is_scalar(e) or int(e) != e or e < 0 or e > 16:
            raise NotImplementedError("only real ** small non-negative integer is implemented")
        return self._bin(_capi.EW_POW, float(e))

    def __iadd__(self, o):
        self[...] = self + o
        return self

    def __isub__(self, o):
        self[...] = self - o
        return self

    def __imul__(self, o):
        self[...] = self * o
        return self

    def __itruediv__(self, o):
        self[...] = self / o
        return self

    # -- reductions (host scalars: every reduction is a sync, as in the reference
    #    where Python min()/if consume device scalars, Integrator.py:37-44)
    def max(self): return max(self)
    def min(self): return min(self)
    def sum(self): return sum(self)
    def mean(self): return mean(self)


def _view2(t):
    """(ptr, rows, cols, row_stride, col_stride) of a <=2-D (or contiguous n-D) tensor."""
    if t.dim() > 2:
        if not t.is_contiguous():
            raise NotImplementedError("only contiguous arrays with more than 2 dimensions")
        t = t.reshape(-1, t.shape[-1])
    if t.dim() == 0:
        return t.data_ptr(), 1, 1, 0, 0
    if t.dim() == 1:
        return t.data_ptr(), 1, t.shape[0], 0, t.stride(0)
    return t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), t.stride(1)


def _operand(x, shape):
    """-> (kind, View, re, im, keepalive) broadcast to `shape` (<= 2-D)."""
    if isinstance(x, DeviceArray):
        t = x._touch()._t
        if not (t.is_complex() or t.is_floating_point()):
            raise TypeError("integer device arrays do not support arithmetic")
        if tuple(t.shape) != tuple(shape):
            t = t.expand(tuple(shape))
        ptr, _, _, rs, cs = _view2(t)
        kind = _capi.KIND_CPLX if t.is_complex() else _capi.KIND_REAL
        return kind, _capi.View(ptr, rs, cs), 0.0, 0.0, t
    z = complex(x)
    return _capi.KIND_SCALAR, _capi.View(None, 0, 0), z.real, z.imag, None


def _result_shape(a, b):
    sa = a.shape if isinstance(a, DeviceArray) else ()
    sb = b.shape if isinstance(b, DeviceArray) else ()
    return tuple(np.broadcast_shapes(sa, sb))


def _elementwise(op, a, b, out=None):
    """out = a (op) b ; returns the torch tensor holding the result."""
    if out is None:
        shape = _result_shape(a, b)
        cplx = any((isinstance(x, DeviceArray) and x.is_complex) or _is_complex_scalar(x)
                   for x in (a, b) if x is not None)
        out = _backend.empty(shape, np.complex128 if cplx else np.float64)
    result = out
    if out.numel() == 0:
        return result
    if out.dim() > 2:
        shape = tuple(out.shape)
        flat = out.reshape(-1, shape[-1])
        if flat.data_ptr() != out.data_ptr() or not out.is_contiguous():
            raise NotImplementedError("only contiguous targets with more than 2 dimensions")

        def prep(x):
            if isinstance(x, DeviceArray):
                return DeviceArray(x._touch()._t.expand(shape).reshape(flat.shape))
            return x
        a, b, out = prep(a), (prep(b) if b is not None else None), flat
    shape = tuple(out.shape)
    d = _capi.Ew()
    d.op = op
    optr, rows, cols, ors, ocs = _view2(out)
    d.rows, d.cols = rows, cols
    d.out = _capi.View(optr, ors, ocs)
    d.out_kind = _capi.KIND_CPLX if out.is_complex() else _capi.KIND_REAL
    ka = _operand(a, shape)
    d.a_kind, d.a, d.a_re, d.a_im = ka[0], ka[1], ka[2], ka[3]
    kb = None
    if b is not None:
        kb = _operand(b, shape)
        d.b_kind, d.b, d.b_re, d.b_im = kb[0], kb[1], kb[2], kb[3]
    else:
        d.b_kind = _capi.KIND_SCALAR
    _ctx().call("mlv_elementwise", ctypes.byref(d))
    del ka, kb
    return result


def _reduce(op, a, b=None):
    a._touch()
    if a.is_complex:
        raise NotImplementedError("reductions are implemented for real arrays")
    t = a._t if a._t.dim() <= 2 else a._t.contiguous()
    ptr, rows, cols, rs, cs = _view2(t)
    if rows * cols == 0:
        raise ValueError("zero-size array to reduction operation")
    va = _capi.View(ptr, rs, cs)
    vb = None
    if b is not None:
        b._touch()
        bp, _, _, brs, bcs = _view2(b._t.expand(tuple(a.shape)))
        vb = _capi.View(bp, brs, bcs)
    out = _backend.empty((1,), np.float64)
    _ctx().call("mlv_reduce", op, rows, cols, ctypes.byref(va),
                ctypes.byref(vb) if vb is not None else None, ctypes.c_void_p(out.data_ptr()))
    val = float(_backend.to_host(out)[0])
    if a._dist is not None:                        # slabs: combine over the ranks
        from . import _dist
        kind = {_capi.RED_MAX: "max", _capi.RED_MIN: "min"}.get(op, "sum")
        val = float(_dist.all_reduce_host([val], kind)[0])
    return val


# ------------------------------------------------------ namespace functions
def _as_dev(x):
    if isinstance(x, DeviceArray):
        return x
    if isinstance(x, (SpecExpr, LazyLap)):
        return x.materialize()
    return DeviceArray(_backend.from_host(np.asarray(x)))


def zeros(shape, dtype=np.float64):
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    return DeviceArray(_backend.zeros(shape, dtype))


def empty(shape, dtype=np.float64):
    if isinstance(shape, (int, np.integer)):
        shape = (int(shape),)
    return DeviceArray(_backend.empty(shape, dtype))


def ones(shape, dtype=np.float64):
    out = empty(shape, dtype)
    out[...] = 1.0
    return out


def zeros_like(a, dtype=None):
    out = zeros(a.shape, dtype or a.dtype)
    out._dist = getattr(a, "_dist", None)
    return out


def empty_like(a, dtype=None):
    out = empty(a.shape, dtype or a.dtype)
    out._dist = getattr(a, "_dist", None)
    return out


def array(obj, dtype=None):
    if isinstance(obj, DeviceArray):
        return obj.copy()
    return DeviceArray(_backend.from_host(np.array(obj, dtype=dtype)))


asarray = array


def asnumpy(a):
    return a.get() if isinstance(a, (DeviceArray, SpecExpr, LazyLap)) else np.asarray(a)


def arange(*args, **kw):
    return DeviceArray(_backend.from_host(np.arange(*args, **kw)))


def meshgrid(*xi, **kw):
    """Set-up helper of the namespace (ArrayFactory.py:26 builds its mode-number matrices with it):
    composed on the host, one upload per result."""
    return [DeviceArray(_backend.from_host(g)) for g in np.meshgrid(*[asnumpy(x) for x in xi], **kw)]


def concatenate(arrays, axis=0):
    """Set-up helper of the namespace (ArrayFactory.py:12, SpectralTransformer.py:170-186)."""
    return DeviceArray(_backend.from_host(np.concatenate([asnumpy(a) for a in arrays], axis=axis)))


def max(a):   # noqa: A001  (NumPy name)
    return _reduce(_capi.RED_MAX, _as_dev(a))


def min(a):   # noqa: A001
    return _reduce(_capi.RED_MIN, _as_dev(a))


def sum(a):   # noqa: A001
    return _reduce(_capi.RED_SUM, _as_dev(a))


def mean(a):
    a = _as_dev(a)
    size = a.size
    if a._dist is not None:                        # global element count
        size = int(np.prod([a._dist.count if i == a._dist.axis else n for i, n in enumerate(a.shape)]))
    return _reduce(_capi.RED_SUM, a) / size


def sum_of_squares(a):
    return _reduce(_capi.RED_SUMSQ, _as_dev(a))


def sum_of_product(a, b):
    return _reduce(_capi.RED_SUMPROD, _as_dev(a), _as_dev(b))


class _OutputPipeline:
    """Asynchronous frame output (Variable.save -> xp.save, reference melvin/Variable.py:130-133):
    the array is snapshotted into a device staging buffer by the library's copy kernel on the
    compute stream, travels to pinned host memory on a separate copy stream, and is written to
    disk by a background thread -- the time step never waits for PCIe or the file system.
    Files are complete after flush() (called at interpreter exit, by load() and by
    synchronize()); MLV_SYNC_OUTPUT=1 restores the reference's blocking behaviour."""
    SLOTS = 2

    def __init__(self):
        import queue
        import threading
        self._queue = queue.Queue()
        self._free = []                      # (device staging, pinned host) pairs by byte size
        self._busy = 0
        self._cv = threading.Condition()
        self._stream = None
        self._thread = threading.Thread(target=self._writer, daemon=True)
        self._thread.start()
        self.frames_written = 0

    def _writer(self):
        while True:
            item = self._queue.get()
            if item is None:
                return
            fname, host, shape, dtype, event, slot = item
            try:
                event.synchronize()
                np.save(fname, host.numpy().view(dtype).reshape(shape))
                self.frames_written += 1
            finally:
                with self._cv:
                    self._free.append(slot)
                    self._busy -= 1
                    self._cv.notify_all()

    def submit(self, fname, a):
        import torch
        t = a._touch()._t
        nbytes = t.numel() * t.element_size()
        with self._cv:
            while self._busy >= self.SLOTS:              # both staging pairs in flight: wait for one
                self._cv.wait()
            slot = next((s for s in self._free if s[0].numel() == nbytes), None)
            if slot is not None:
                self._free.remove(slot)
            self._busy += 1
        if slot is None:
            slot = (torch.empty(nbytes, dtype=torch.uint8, device=t.device),
                    torch.empty(nbytes, dtype=torch.uint8, pin_memory=True))
        stage, host = slot
        dtype = torch.complex128 if t.is_complex() else torch.float64
        snap = stage.view(dtype).reshape(tuple(t.shape))
        _elementwise(_capi.EW_COPY, a, None, out=snap)          # snapshot on the compute stream
        if self._stream is None:
            self._stream = torch.cuda.Stream()
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        done = torch.cuda.Event()
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            host.copy_(stage, non_blocking=True)
            done.record(self._stream)
        self._queue.put((fname, host, tuple(t.shape), np.complex128 if t.is_complex() else np.float64,
                         done, slot))

    def flush(self):
        with self._cv:
            while self._busy:
                self._cv.wait()


_output = None


def flush_output():
    """Block until every frame handed to save() is on disk."""
    if _output is not None:
        _output.flush()


def save(fname, a):
    """numpy.save of a device array (reference: xp.save in Variable.save).  Device arrays on a
    GPU go through the asynchronous output pipeline."""
    import os
    global _output
    fname = fname if str(fname).endswith(".npy") else str(fname) + ".npy"
    if isinstance(a, DeviceArray) and a._dist is not None:
        from . import _dist
        full = a.get()                               # collective gather; one writer
        if _dist.rank() == 0:
            np.save(fname, full)
        return
    if (isinstance(a, DeviceArray) and _backend.is_cuda() and not os.environ.get("MLV_SYNC_OUTPUT")
            and a._touch()._t.is_contiguous()):
        if _output is None:
            import atexit
            _output = _OutputPipeline()
            atexit.register(flush_output)
        _output.submit(fname, a)
        return
    np.save(fname, asnumpy(a))


def savez(fname, **arrays):
    from . import _dist
    host = {k: asnumpy(v) for k, v in arrays.items()}
    if _dist.rank() == 0:
        np.savez(fname, **host)


def load(fname, **kw):
    flush_output()
    return np.load(fname, **kw)


def synchronize():
    _backend.synchronize()
    flush_output()


class _FFT:
    """The reference library calls xp.fft.* itself; this backend performs the
    transforms inside SpectralTransformer, so the namespace has nothing to offer."""

    def __getattr__(self, name):
        raise NotImplementedError(
            f"xp.fft.{name}: transforms are provided by melvin.SpectralTransformer on this backend")


fft = _FFT()


# ================================================================ SpecExpr
def _frozen(a):
    """Plain array on the buffer `a` occupies now (materialising a deferred definition)."""
    if type(a) is DeviceArray and a._base is None:
        return a
    return DeviceArray(a._touch()._t)


def _eval_terms(ctx, terms, out_t, accumulate):
    """out (+)= sum_i c_i op_i(a_i), four operands per launch."""
    terms = list(terms)
    first = not accumulate
    while terms:
        room = _capi.MAXLIN if first else _capi.MAXLIN - 1
        chunk, terms = terms[:room], terms[room:]
        packed = [(c, op, a._touch()._t.data_ptr()) for c, op, a in chunk]
        if any(p[2] == out_t.data_ptr() and p[1] >= _capi.OP_FDM_D2DZ2 for p in packed):
            raise NotImplementedError("a row stencil cannot be evaluated in place")
        if not first:
            packed.append((1.0, _capi.OP_IDENT, out_t.data_ptr()))
        lt = _capi.make_lin_terms(packed)
        ctx.call("mlv_spec_lincomb", ctypes.byref(lt), ctypes.c_void_p(out_t.data_ptr()))
        first = False


class NLTerm:
    """Handle on the z-spectra (IA, IB) of the products ux*q, uz*q produced by the
    fused physical-space stage; consumed by the forward x pass."""

    def __init__(self, ctx, ia, ib):
        self.ctx, self.ia, self.ib = ctx, ia, ib

    def lin_terms(self, coef):
        """FDM-z mode: ia, ib are the x spectra of ux*q, uz*q and the nonlinear term is
        d/dx(ux q) + d/dz(uz q) = (i symx) ia + pddz(ib), two row-wise linear terms."""
        return [(complex(coef), _capi.OP_FDX_SYM, DeviceArray(self.ia)),
                (complex(coef), _capi.OP_FDM_DDZ, DeviceArray(self.ib))]

    def __del__(self):
        try:
            self.ctx.give_i(self.ia)
            self.ctx.give_i(self.ib)
        except Exception:
            pass


class SpecExpr:
    """sum_i c_i * op_i(a_i)  +  sum_k c_k * NL_k   (spectral-shaped, complex128)."""
    __array_priority__ = 2000
    __array_ufunc__ = None

    def __init__(self, ctx, terms=(), nls=()):
        self.ctx = ctx
        # (complex coef, op code, DeviceArray).  Operands are captured as plain arrays on the
        # buffer they live in *now*: the expression is registered with the context so that
        # whoever is about to overwrite that buffer evaluates the affected terms first
        # (the reference evaluates right-hand sides eagerly; see Variable._flush_dependants)
        self.terms = [(c, op, _frozen(a)) for c, op, a in terms]
        self.nls = list(nls)          # (float coef, NLTerm)
        if self.terms:
            ctx._lazy_exprs.add(self)

    def _reads(self, tensor):
        ptr = tensor.data_ptr()
        return any(a._t.data_ptr() == ptr for _, _, a in self.terms)

    def _detach_from(self, tensor):
        """`tensor` is about to change: evaluate the terms that read it into a private buffer."""
        ptr = tensor.data_ptr()
        hit = [t for t in self.terms if t[2]._t.data_ptr() == ptr]
        if not hit:
            return
        tmp = _backend.empty(self.ctx.spec_shape, np.complex128)
        _eval_terms(self.ctx, hit, tmp, False)
        self.terms = [t for t in self.terms if t[2]._t.data_ptr() != ptr]
        self.terms.append((1.0 + 0j, _capi.OP_IDENT, DeviceArray(tmp)))

    # -- metadata
    @property
    def shape(self):
        return self.ctx.spec_shape

    @property
    def dtype(self):
        return np.dtype(np.complex128)

    # -- algebra
    def _scaled(self, c):
        c = complex(c)
        if self.nls and c.imag != 0.0:
            return SpecExpr(self.ctx, [(c, _capi.OP_IDENT, self.materialize())])
        return SpecExpr(self.ctx, [(k * c, op, a) for k, op, a in self.terms],
                        [(k * c.real, n) for k, n in self.nls])

    @staticmethod
    def _lift(ctx, x):
        if isinstance(x, SpecExpr):
            return x
        if isinstance(x, DeviceArray) and x.is_complex and x.shape == ctx.spec_shape \
                and x._touch()._t.is_contiguous():
            return SpecExpr(ctx, [(1.0 + 0j, _capi.OP_IDENT, x)])
        return None

    def __neg__(self): return self._scaled(-1.0)
    def __pos__(self): return self

    def __mul__(self, o):
        if _is_scalar(o):
            return self._scaled(o)
        return self.materialize() * o

    __rmul__ = __mul__

    def __truediv__(self, o):
        if _is_scalar(o):
            return self._scaled(1.0 / complex(o) if isinstance(o, (complex, np.complexfloating))
                                else 1.0 / float(o))
        return self.materialize() / o

    def __rtruediv__(self, o):
        return o / self.materialize()

    def _add(self, o, sign):
        other = self._lift(self.ctx, o)
        if other is None:
            m = self.materialize()
            return m + o if sign > 0 else m - o
        if sign < 0:
            other = other._scaled(-1.0)
        return SpecExpr(self.ctx, self.terms + other.terms, self.nls + other.nls)

    def __add__(self, o): return self._add(o, +1)
    def __radd__(self, o): return self._add(o, +1)
    def __sub__(self, o): return self._add(o, -1)

    def __rsub__(self, o):
        return self._scaled(-1.0)._add(o, +1)

    # -- evaluation
    def materialize(self, out=None):
        """Evaluate into a (new) spectral DeviceArray."""
        ctx = self.ctx
        if out is None:
            out = DeviceArray(_backend.empty(ctx.spec_shape, np.complex128))
            if ctx.world > 1:
                out._dist = Dist(ctx, "cols", 1, 0, ctx.nm)
        out_t = out._t
        terms = list(self.terms)
        first = True
        nls = list(self.nls)
        if ctx.fdm_z:
            for coef, nl in nls:
                terms += nl.lin_terms(coef)
            nls = []
        while nls:
            chunk, nls = nls[:2], nls[2:]
            d = _capi.XFwd()
            d.nf, d.mode = 2 * len(chunk), 0
            for i, (coef, nl) in enumerate(chunk):
                d.src[2 * i], d.src[2 * i + 1] = nl.ia.data_ptr(), nl.ib.data_ptr()
                d.sym[2 * i], d.sym[2 * i + 1] = _capi.SYM_FDX, _capi.SYM_FDZ
                d.coef[2 * i] = d.coef[2 * i + 1] = coef
            if first:
                d.dst = out_t.data_ptr()
                ctx.call("mlv_x_forward", ctypes.byref(d))
                first = False
            else:
                tmp = _backend.empty(ctx.spec_shape, np.complex128)
                d.dst = tmp.data_ptr()
                ctx.call("mlv_x_forward", ctypes.byref(d))
                terms.append((1.0 + 0j, _capi.OP_IDENT, DeviceArray(tmp)))
        if not terms and first:
            out[...] = 0.0
            return out
        _eval_terms(ctx, terms, out_t, not first)
        return out

    # -- array-like fallbacks
    def __getitem__(self, idx):
        return self.materialize()[idx]

    def __array__(self, dtype=None, copy=None):
        return self.materialize().__array__(dtype)

    def get(self):
        return self.materialize().get()


# ================================================================= LazyLap
class LazyLap:
    """``coef * (d2x n^2 + d2z m^2)`` -- Variable.lap() / SpatialDifferentiator.calc_lap
    (reference melvin/SpatialDifferentiator.py:70-74) kept symbolic."""
    __array_priority__ = 2000
    __array_ufunc__ = None

    def __init__(self, ctx, coef=1.0):
        self.ctx, self.coef = ctx, float(coef)
        self._mat = None

    @property
    def shape(self):
        return self.ctx.spec_shape

    @property
    def dtype(self):
        return np.dtype(np.float64)

    def __mul__(self, o):
        if _is_scalar(o) and complex(o).imag == 0.0:
            return LazyLap(self.ctx, self.coef * float(np.real(o)))
        return self.materialize() * o

    __rmul__ = __mul__

    def __truediv__(self, o):
        if _is_scalar(o) and complex(o).imag == 0.0:
            return LazyLap(self.ctx, self.coef / float(np.real(o)))
        return self.materialize() / o

    def __neg__(self):
        return LazyLap(self.ctx, -self.coef)

    def materialize(self):
        if self._mat is None:
            out = _backend.empty(self.ctx.spec_shape, np.float64)
            self.ctx.call("mlv_lap_array", ctypes.c_double(self.coef), ctypes.c_void_p(out.data_ptr()))
            self._mat = DeviceArray(out)
            if self.ctx.world > 1:
                self._mat._dist = Dist(self.ctx, "cols", 1, 0, self.ctx.nm)
        return self._mat

    def __add__(self, o): return self.materialize() + o
    def __radd__(self, o): return o + self.materialize()
    def __sub__(self, o): return self.materialize() - o
    def __rsub__(self, o): return o - self.materialize()
    def __rtruediv__(self, o): return o / self.materialize()
    def __getitem__(self, idx): return self.materialize()[idx]

    def __setitem__(self, idx, v):
        self.materialize()[idx] = v

    def __array__(self, dtype=None, copy=None):
        return self.materialize().__array__(dtype)

    def get(self):
        return self.materialize().get()
