"""Field objects: ``Variable`` (spectral + physical pair) and ``TimeDerivative``
(ring buffer of right-hand-side history).

API mirror of the reference's ``melvin/Variable.py`` and
``melvin/TimeDerivative.py``; the storage is device memory and the operators are
the C-ABI kernels.  On top of the reference semantics a Variable keeps a
*private* x-transformed intermediate (``_i``) so that the physical field never
has to be written to HBM inside the time step:

    spectral  --x pass-->  I (nx, nm)  --z pass-->  physical

``to_physical()`` only records the request; the x pass is issued when the
intermediate is first needed (batched with the other fields of the same
``vec_dot_nabla``), the z pass only when somebody really reads ``getp()``.
Whatever is read is always what the reference would have returned: pending
work captures the buffer it was defined on, in-place writes flush dependants
first, and ``Integrator`` double-buffers a state that still has dependants.
"""
import ctypes
import weakref

import numpy as np

from . import _backend, _capi
from .b200 import DeviceArray, Dist, LazyLap, NLTerm, SpecExpr, _frozen
from .basis import BasisFunctions


def _require_device_namespace(xp):
    from .operators import _require_device_namespace as f      # (operators imports this module)
    return f(xp)

_I_NONE, _I_PENDING, _I_VALID = 0, 1, 2


class _SpecHandle(DeviceArray):
    """Spectral buffer of a Variable: reads materialise a deferred definition,
    in-place writes first flush everything that still depends on the old data."""

    def __init__(self, tensor, owner):
        super().__init__(tensor)
        self._owner = weakref.ref(owner)

    def _touch(self):
        o = self._owner()
        if o is not None:
            o._materialize_s()
        return self

    def _pre_write(self):
        o = self._owner()
        if o is not None:
            o._flush_dependants(self._t)


class _PhysHandle(DeviceArray):
    """Physical buffer of a Variable: the z pass runs on first touch."""

    def __init__(self, tensor, owner):
        super().__init__(tensor)
        self._owner = weakref.ref(owner)

    def _touch(self):
        o = self._owner()
        if o is not None:
            o._ensure_p()
        return self

    def _pre_write(self):
        o = self._owner()
        if o is not None:
            o._p_written()


class Variable:
    """Encapsulates the physical and spectral representations of a variable
    (reference melvin/Variable.py:6-37)."""

    def __init__(self, params, xp, sd=None, st=None, dt=None, array_factory=None,
                 dump_name=None, basis_functions=None):
        self._params = params
        self._xp = _require_device_namespace(xp)
        self._st = st
        self._sd = sd
        self._dt = dt
        self._array_factory = array_factory
        self._dump_name = dump_name
        if basis_functions is None:
            raise Exception("Basis functions must be specified.")
        self._basis_functions = basis_functions
        self._ctx = _backend.context_for(params)
        self._fused = (not self._ctx.fdm_z
                       and basis_functions[0] is BasisFunctions.COMPLEX_EXP
                       and basis_functions[1] is BasisFunctions.COMPLEX_EXP)
        # Fourier-x / FDM-z: the "intermediate" is an x spectrum (nn, nz) itself -- the state,
        # or a private buffer for ux -- and the deferred physical field is its c2r transform
        self._fused_fdm = (self._ctx.fdm_z and basis_functions[0] is BasisFunctions.COMPLEX_EXP
                           and basis_functions[1] is BasisFunctions.FDM)
        self._hat = None                     # FDM-z: private x spectrum (ux = -pddz(psi))
        # ONE handle per Variable for its whole life (gets() / _sdata always return it, as the
        # reference returns the same ndarray); double buffering re-points its tensor
        # (under a process group: this rank's kz-slab / x-slab, melvin/_dist.py)
        ctx = self._ctx
        self._s = _SpecHandle(_backend.zeros(ctx.spec_shape, np.complex128), self)
        self._s_spare = None                 # second state tensor (double buffering)
        self._p = _PhysHandle(_backend.zeros(ctx.phys_shape, np.float64), self)
        if ctx.world > 1:
            self._s._dist = Dist(ctx, "cols", 1, 0, ctx.nm)
            self._p._dist = Dist(ctx, "rows", 0, 0, ctx.nx)
        self._virt = None                    # (op, frozen DeviceArray): deferred spectral definition
        self._i = None                       # torch tensor (nx, ipitch) complex128
        self._i_state = _I_NONE
        self._i_def = None                   # (op, DeviceArray) the pending x pass reads
        self._p_valid = True
        self._red = None                     # (shared {"part", "host"}, max index, sumsq index)
        self._red_part = None                # per-CTA partials of the fused z stage (velocity owner)
        self._ctx_vars().add(self)

    # ------------------------------------------------------------ registry
    def _ctx_vars(self):
        reg = getattr(self._ctx, "_lazy_vars", None)
        if reg is None:
            reg = weakref.WeakSet()
            self._ctx._lazy_vars = reg
        return reg

    # Pending work (deferred definitions, requested transforms, deferred right-hand sides) reads
    # the *tensor* it was defined on.  Whoever overwrites a tensor flushes its dependants first,
    # so every read returns what the reference's eager evaluation would have returned.
    def _depends_on(self, tensor):
        ptr = tensor.data_ptr()
        return ((self._virt is not None and self._virt[1]._t.data_ptr() == ptr)
                or (self._i_state == _I_PENDING and self._i_def[1]._t.data_ptr() == ptr))

    def _flush_dependants(self, tensor):
        """`tensor` is about to change in place."""
        ptr = tensor.data_ptr()
        for v in list(self._ctx_vars()):
            if v._i_state == _I_PENDING and v._i_def[1]._t.data_ptr() == ptr:
                v._ensure_i()
            if v._virt is not None and v._virt[1]._t.data_ptr() == ptr:
                v._materialize_s()
        for e in list(self._ctx._lazy_exprs):
            e._detach_from(tensor)

    def _has_dependants(self, tensor, exclude=None):
        return (any(v._depends_on(tensor) for v in self._ctx_vars())
                or any(e is not exclude and e._reads(tensor) for e in list(self._ctx._lazy_exprs)))

    # ---------------------------------------------------- spectral storage
    @property
    def _sdata(self):
        self._materialize_s()
        return self._s

    @_sdata.setter
    def _sdata(self, value):
        self.sets(value)

    @property
    def _pdata(self):
        return self._p

    def _materialize_s(self):
        if self._virt is None:
            return
        op, src = self._virt
        self._virt = None
        lt = _capi.make_lin_terms([(1.0, op, src._t.data_ptr())])
        self._ctx.call("mlv_spec_lincomb", ctypes.byref(lt), ctypes.c_void_p(self._s._t.data_ptr()))

    def _set_virtual(self, op, src):
        """Spectral data := op(src) without writing it (fully spectral mode)."""
        self._flush_dependants(self._s._t)
        self._virt = (op, _frozen(src))

    def _spec_def(self):
        return self._virt if self._virt is not None else (_capi.OP_IDENT, DeviceArray(self._s._t))

    def __setitem__(self, index, value):
        self._materialize_s()
        self._s[index] = value

    def __getitem__(self, index):
        full = index == slice(None) if isinstance(index, slice) else index is Ellipsis
        if full and self._virt is not None:
            op, src = self._virt
            return SpecExpr(self._ctx, [(1.0 + 0j, op, src)])
        return self._sdata[index]

    def sets(self, data):
        """Setter for spectral data"""
        if isinstance(data, SpecExpr):
            self._flush_dependants(self._s._t)
            self._virt = None
            data.materialize(out=DeviceArray(self._s._t))
            return
        self._flush_dependants(self._s._t)
        self._virt = None
        if self._ctx.world > 1 and getattr(data, "_dist", None) is not None:
            DeviceArray(self._s._t)[:, :] = DeviceArray(data._touch()._t)[:, :]     # slab to slab
        else:
            self._s[:, :] = data[:, :]              # (slabs: a full array, cut by the handle)

    def gets(self):
        return self._sdata

    # ---------------------------------------------------- physical storage
    def setp(self, data):
        """Setter for physical data"""
        ctx = self._ctx
        if ctx.world > 1 and tuple(data.shape) == (ctx.nx, ctx.nz) and getattr(data, "_dist", None) is None:
            data = data[ctx.x_off:ctx.x_off + ctx.nxl]          # a full field: keep this rank's rows
        if isinstance(data, np.ndarray):
            data = DeviceArray(_backend.from_host(data))
        self._p_written()
        DeviceArray(self._p._t)[:, :] = DeviceArray(data._touch()._t)[:, :]

    def getp(self):
        return self._p

    def _p_written(self):
        self._i_state = _I_NONE
        self._i_def = None
        self._p_valid = True
        self._red = None

    def _ensure_i(self):
        """Make the x-transformed intermediate valid (runs the pending x pass)."""
        if self._i_state == _I_PENDING:
            if self._ctx.fdm_z:
                self._fdm_physical()
            else:
                _run_x_inverse(self._ctx, [self])
        return self._i_state == _I_VALID

    def _fdm_physical(self):
        """FDM-z: evaluate the pending definition `pdata = c2r(op(x spectrum))`."""
        op, src = self._i_def
        ctx, vp = self._ctx, ctypes.c_void_p
        self._i_def, self._i_state, self._p_valid = None, _I_VALID, True
        if op == _capi.OP_PSI:           # streamfunction of a vorticity: psi = solve(-w) = -solve(w)
            tmp = ctx.take_i()
            ctx.call("mlv_solve_fdm", vp(src._t.data_ptr()), vp(tmp.data_ptr()))
            ctx.call("mlv_to_physical", vp(tmp.data_ptr()), None, vp(self._p._t.data_ptr()))
            ctx.give_i(tmp)
            DeviceArray(self._p._t)[...] = DeviceArray(self._p._t) * -1.0
        else:
            ctx.call("mlv_to_physical", vp(src._t.data_ptr()), None, vp(self._p._t.data_ptr()))

    def _ensure_p(self):
        if self._p_valid:
            return
        if self._ctx.fdm_z:
            self._ensure_i()
            return
        self._ensure_i()
        self._ctx.call("mlv_z_inverse", ctypes.c_void_p(self._i.data_ptr()),
                       ctypes.c_void_p(self._p._t.data_ptr()))
        self._p_valid = True

    def _request_physical(self):
        """Record `pdata = T(sdata)` for the current spectral definition."""
        self._i_def = self._spec_def()
        self._i_state = _I_PENDING
        self._p_valid = False
        self._red = None

    # ------------------------------------------------------- transforms
    def to_physical(self):
        """Convert spectral data to physical"""
        if self._fused or self._fused_fdm:
            self._request_physical()
        else:
            self._st.to_physical(self._sdata, DeviceArray(self._p._t), self._basis_functions)
            self._p_written()

    def to_spectral(self):
        """Convert physical data to spectral"""
        self._flush_dependants(self._s._t)
        self._virt = None
        self._st.to_spectral(self._p, DeviceArray(self._s._t), self._basis_functions)

    def load(self, data, is_physical=False):
        if isinstance(data, str):
            raise NotImplementedError
        if is_physical:
            self.setp(self._dt.from_host(data))
            self.to_spectral()
        else:
            # the reference compares data.shape with (nn, nm), which never matches, and always
            # takes the (complex-discarding) scaling route (Variable.py:82); here a
            # state of the right shape is loaded as it is and any other resolution is re-sampled
            # in spectral space, on the host for host data (one upload of the final size)
            if tuple(data.shape) != tuple(self._params.spectral_shape):
                # (FDM-z mode has no params.nm: re-sampling fails there as in the reference, App. A-14)
                nn, nm = self._params.nn, self._params.nm
                data = scale_variable(data, (nn, nm), np if isinstance(data, np.ndarray) else self._xp)
            if (isinstance(data, np.ndarray) and self._ctx.world == 1
                    and tuple(data.shape) == tuple(self._s._t.shape)):
                # host state of the right shape: straight into the state buffer (asynchronous
                # on the current stream when `data` lives in pinned memory)
                self._flush_dependants(self._s._t)
                self._virt = None
                _backend.copy_from_host(self._s._t, data)
            else:
                self.sets(self._dt.from_host(data))

    # ------------------------------------------------------ derivatives
    def pddx(self):
        """Calculate spatial derivative of physical data"""
        return self._sd.pddx(self.getp())

    def pddz(self):
        """Calculate spatial derivative of physical data"""
        return self._sd.pddz(self.getp())

    def _sterm(self, op):
        cop, src = self._spec_def()
        if cop == _capi.OP_IDENT:
            return SpecExpr(self._ctx, [(1.0 + 0j, op, src)])
        return SpecExpr(self._ctx, [(1.0 + 0j, op, self._sdata)])

    def sddx(self):
        """Calculate first derivative of spectral data"""
        return self._sd.sddx(self._spec_operand(), self._basis_functions[0])

    def sddz(self):
        """Calculate first derivative of spectral data"""
        return self._sd.sddz(self._spec_operand(), self._basis_functions[1])

    def sd2dx2(self):
        """Calculate second derivative of spectral data"""
        return self._sd.sd2dx2(self._spec_operand(), self._basis_functions[0])

    def sd2dz2(self):
        """Calculate second derivative of spectral data"""
        return self._sd.sd2dz2(self._spec_operand(), self._basis_functions[1])

    def _spec_operand(self):
        return self._sdata

    def snabla2(self):
        """Calculate nabla^2 in spectral form"""
        if self._fused:
            return SpecExpr(self._ctx, [(1.0 + 0j, _capi.OP_LAP, self._sdata)])
        if self._fused_fdm:
            return SpecExpr(self._ctx, [(1.0 + 0j, _capi.OP_FDM_NABLA2, self._sdata)])
        return self.sd2dx2() + self.sd2dz2()

    def lap(self):
        """Returns linear operator representing laplacian"""
        return self._sd.calc_lap(self._basis_functions)

    # --------------------------------------------------- nonlinear term
    def vec_dot_nabla(self, ux, uz, out=None, convert_to_physical=True):
        """d/dx(ux q) + d/dz(uz q) -> spectral (reference melvin/Variable.py:119-128)."""
        uxo = ux._owner() if isinstance(ux, _PhysHandle) else None
        uzo = uz._owner() if isinstance(uz, _PhysHandle) else None
        if (self._fused_fdm and out is None and uxo is not None and uzo is not None
                and uxo._fused_fdm and uzo._fused_fdm
                and uxo._i_state == _I_PENDING and uzo._i_state == _I_PENDING
                and uxo._i_def[0] == _capi.OP_IDENT and uzo._i_def[0] == _capi.OP_IDENT):
            if convert_to_physical:
                self._request_physical()
            if self._i_state == _I_PENDING and self._i_def[0] == _capi.OP_IDENT:
                return self._vec_dot_nabla_fdm(uxo, uzo)
        fused = (self._fused and out is None and uxo is not None and uzo is not None
                 and uxo._fused and uzo._fused
                 and uxo._i_state != _I_NONE and uzo._i_state != _I_NONE)
        if fused:
            if convert_to_physical:
                self._request_physical()
            fused = self._i_state != _I_NONE
        if not fused:
            return self._vec_dot_nabla_eager(ux, uz, out, convert_to_physical)
        ctx = self._ctx
        # scalar first: when it is the vorticity itself the kernel reads the column once,
        # then turns its copy into psi for the two velocity components
        _run_x_inverse(ctx, [v for v in (self, uxo, uzo) if v._i_state == _I_PENDING])
        ia, ib = ctx.take_i(), ctx.take_i()
        # the CFL / kinetic-energy reductions of the velocities ride in this kernel, but the
        # tickers want them only every cfl_cadence / tracker_cadence loops: the kernel leaves
        # its per-CTA partials in a buffer owned by the x velocity, _cached_reduction combines
        # them when (and if) somebody asks
        if uxo._red_part is None:
            uxo._red_part = _backend.empty((ctx.red_doubles,), np.float64)
        reduced = ctx.sync_reduction_mode()
        ctx.call("mlv_set_reduction_partials", ctypes.c_void_p(uxo._red_part.data_ptr()), count=False)
        ctx.call("mlv_advect_z", ctypes.c_void_p(uxo._i.data_ptr()), ctypes.c_void_p(uzo._i.data_ptr()),
                 ctypes.c_void_p(self._i.data_ptr()), ctypes.c_void_p(ia.data_ptr()),
                 ctypes.c_void_p(ib.data_ptr()), None)
        ctx.call("mlv_set_reduction_partials", None, count=False)
        if ctx.world > 1:                           # tile block h of both products to rank h
            sa, sb = ia, ib
            ia, ib = ctx.exchange(sa, True), ctx.exchange(sb, True)
            ctx.give_i(sa)
            ctx.give_i(sb)
        if reduced:
            shared = {"part": uxo._red_part, "host": None}
            uxo._red = (shared, 0, 2)
            uzo._red = (shared, 1, 3)
        else:                       # nobody asked: a reader falls back to an explicit reduction
            uxo._red = uzo._red = None
        return SpecExpr(ctx, [], [(1.0, NLTerm(ctx, ia, ib))])

    def _vec_dot_nabla_fdm(self, uxo, uzo):
        """Fourier-x / FDM-z: fused c2r x 3 -> products -> r2c x 2 on the x spectra of the
        velocities and of the scalar; the two derivatives become row-wise linear terms of the
        right-hand side (NLTerm.lin_terms)."""
        ctx, vp = self._ctx, ctypes.c_void_p
        ia, ib = ctx.take_i(), ctx.take_i()
        if uxo._red_part is None:
            uxo._red_part = _backend.empty((ctx.red_doubles,), np.float64)
        reduced = ctx.sync_reduction_mode()
        ctx.call("mlv_set_reduction_partials", vp(uxo._red_part.data_ptr()), count=False)
        ctx.call("mlv_fdm_advect", vp(uxo._i_def[1]._t.data_ptr()), vp(uzo._i_def[1]._t.data_ptr()),
                 vp(self._i_def[1]._t.data_ptr()), vp(ia.data_ptr()), vp(ib.data_ptr()), None)
        ctx.call("mlv_set_reduction_partials", None, count=False)
        if reduced:
            shared = {"part": uxo._red_part, "host": None}
            uxo._red = (shared, 0, 2)
            uzo._red = (shared, 1, 3)
        else:
            uxo._red = uzo._red = None
        return SpecExpr(ctx, [], [(1.0, NLTerm(ctx, ia, ib))])

    def _vec_dot_nabla_eager(self, ux, uz, out, convert_to_physical):
        if self._ctx.world > 1:
            raise NotImplementedError("slab-decomposed runs advect with the velocities of "
                                      "calc_velocity_from_vorticity (the fused path); a physical-space "
                                      "x stencil on raw arrays would need a halo exchange")
        if convert_to_physical:
            self.to_physical()
        if isinstance(ux, np.ndarray):
            ux = DeviceArray(_backend.from_host(ux))
        if isinstance(uz, np.ndarray):
            uz = DeviceArray(_backend.from_host(uz))
        q = self.getp()
        if out is None:
            out = self._xp.zeros_like(self._p)
        args = []
        for a in (ux, uz, q, out):
            t = a._touch()._t
            if not t.is_contiguous() or t.is_complex():
                raise NotImplementedError("vec_dot_nabla needs contiguous real (nx, nz) operands")
            args.append(ctypes.c_void_p(t.data_ptr()))
        self._ctx.call("mlv_advect_phys", *args)
        return self._st.to_spectral(out, basis_functions=self._basis_functions)

    # ------------------------------------------------------ reductions
    def _cached_reduction(self, which):
        """max (which=1) or sum of squares (which=2) of the physical field if the fused
        z stage already produced it for the current intermediate."""
        if self._red is None or self._i_state == _I_NONE:
            return None
        shared = self._red[0]
        if shared["host"] is None:
            red4 = _backend.empty((4,), np.float64)
            self._ctx.call("mlv_reduce_partials", ctypes.c_void_p(shared["part"].data_ptr()),
                           ctypes.c_void_p(red4.data_ptr()))
            shared["host"] = _backend.to_host(red4)
            if self._ctx.world > 1:                 # slabs: global maxima and sums
                from . import _dist
                h = shared["host"]
                shared["host"] = np.concatenate((_dist.all_reduce_host(h[:2], "max"),
                                                 _dist.all_reduce_host(h[2:], "sum")))
        return float(shared["host"][self._red[which]])

    # ------------------------------------------------------------- I/O
    def save(self, dump_counter):
        fname = self._dump_name + f"{dump_counter:04d}.npy"
        self._xp.save(fname, self._p)

    def on_host(self, out=None):
        """Host copy of the spectral data (reference melvin/Variable.py: on_host).  With `out`
        (a host array of the right shape; pinned memory makes the copy asynchronous on the
        current stream -- synchronise the stream or an event before reading it) no new host
        array is allocated."""
        if out is not None and self._ctx.world == 1:
            return _backend.copy_to_host(out, self.gets()._touch()._t)
        return self._dt.to_host(self.gets())

    def get_name(self):
        return self._dump_name


def _run_x_inverse(ctx, variables):
    """One inverse x pass for every listed Variable with a pending request."""
    variables = [v for i, v in enumerate(variables) if v not in variables[:i]]
    while variables:
        chunk, variables = variables[:4], variables[4:]
        n = len(chunk)
        srcs = (ctypes.c_void_p * n)()
        ops = (ctypes.c_int32 * n)()
        dsts = (ctypes.c_void_p * n)()
        sends = []
        for k, v in enumerate(chunk):
            op, src = v._i_def
            if ctx.world > 1:                       # [peer][block] send buffer, exchanged below
                sends.append(ctx.take_i())
                dsts[k] = sends[-1].data_ptr()
            else:
                if v._i is None:
                    v._i = _backend.empty((ctx.nx, ctx.ipitch), np.complex128)
                dsts[k] = v._i.data_ptr()
            srcs[k] = src._t.data_ptr()
            ops[k] = op
        ctx.call("mlv_x_inverse", n, srcs, ops, dsts)
        for k, v in enumerate(chunk):
            if ctx.world > 1:                       # row block h of every field to rank h
                ctx.give_i(v._i)
                v._i = ctx.exchange(sends[k], False)
                ctx.give_i(sends[k])
            v._i_state = _I_VALID
            v._i_def = None


def scale_variable(var, outsize, xp):
    """Scale an array in spectral space from its size to `outsize` (nn, nm)
    (reference melvin/Variable.py:142-151, but complex-preserving: the reference
    allocates a real buffer here)."""
    insize = var.shape                       # (2 nn_in + 1, nm_in)
    outvar = xp.zeros((2 * outsize[0] + 1, outsize[1]), dtype=np.complex128)
    # the reference takes insize[0] (= 2 nn_in + 1 rows) for the mode count nn_in, which mixes
    # positive and negative modes whenever the resolutions differ; use the real count
    nx_min = min((insize[0] - 1) // 2, outsize[0])
    nz_min = min(insize[1], outsize[1])
    outvar[: nx_min + 1, :nz_min] = var[: nx_min + 1, :nz_min]
    if nx_min > 0:
        outvar[-nx_min:, :nz_min] = var[-nx_min:, :nz_min]
    return outvar


class TimeDerivative:
    """Represents a sequence of derivatives in time
    (reference melvin/TimeDerivative.py:4-65).  Assigning a deferred expression
    to the current level (``dvar[:] = expr``) is recorded and evaluated by
    ``Integrator.integrate`` (or on first read)."""

    def __init__(self, params, xp, dump_name="", array_factory=None):
        self._params = params
        self._xp = _require_device_namespace(xp)
        self._curr_idx = 0
        ctx = _backend.context_for(params)
        self._store = DeviceArray(_backend.zeros(
            (params.integrator_order,) + tuple(ctx.spec_shape), np.complex128))
        if ctx.world > 1:
            self._store._dist = Dist(ctx, "cols", 2, 0, ctx.nm)
        self._dump_name = dump_name
        self._pending = None

    # -- deferred assignment
    def _flush(self):
        if self._pending is not None:
            expr, self._pending = self._pending, None
            expr.materialize(out=self._level(0))

    def _level(self, back):
        """History level curr_idx+back with Python negative wrap (:38-39)."""
        k = (self._curr_idx + back) % self._params.integrator_order
        lv = DeviceArray(self._store._t[k])
        if self._store._dist is not None:
            lv._dist = Dist(self._store._dist.ctx, "cols", 1, 0, self._store._dist.count)
        return lv

    @property
    def _data(self):
        self._flush()
        return self._store

    def __setitem__(self, index, value):
        full = (isinstance(index, slice) and index == slice(None)) or index is Ellipsis
        if full and isinstance(value, SpecExpr):
            self._pending = value
            return
        self._flush()
        self._level(0)[index] = value

    def __getitem__(self, index):
        self._flush()
        return self._level(0)[index]

    def get_curr_idx(self):
        return self._curr_idx

    def set_curr_idx(self, curr_idx):
        self._flush()
        self._curr_idx = int(curr_idx)

    def advance(self):
        self._flush()
        self._curr_idx = (self._curr_idx + 1) % self._params.integrator_order

    def get(self, idx=0):
        self._flush()
        return self._level(idx)

    def get_all(self):
        return self._data

    def set(self, data, idx=0):
        self._flush()
        self._level(idx)[...] = data

    def load(self, data):
        if isinstance(data, str):
            raise NotImplementedError
        self._flush()
        nn, nm = self._params.nn, self._params.nm
        for i in range(self._params.integrator_order):
            level = data[i]
            if tuple(level.shape) != (2 * nn + 1, nm):
                level = scale_variable(level, (nn, nm), np if isinstance(level, np.ndarray) else self._xp)
            self._level(i - self._curr_idx)[...] = level

    def get_name(self):
        return self._dump_name
