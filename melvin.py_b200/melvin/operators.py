"""Numerical operators: ArrayFactory, SpectralTransformer, SpatialDifferentiator,
LaplacianSolver, Integrator.

API mirrors of the reference classes of the same names (constructor signatures,
slot names, argument meaning and error behaviour); every operation is a call into
libmelvin_b200.so.  Reference locations are cited per method.
"""
import ctypes

import numpy as np

from . import _backend, _capi
from .b200 import DeviceArray, Dist, LazyLap, SpecExpr
from .basis import BasisFunctions

_CE = BasisFunctions.COMPLEX_EXP
_FDM = BasisFunctions.FDM
_COS = BasisFunctions.COSINE
_SIN = BasisFunctions.SINE


def _require_device_namespace(xp):
    """The array namespace the package computes with.  There is no CPU path: `numpy` raises --
    unless MELVIN_B200_NUMPY_IS_DEVICE=1, the switch for running scripts that hard-code `xp = np`
    (examples/resistive_tearing_instability.py:19-20) unchanged: the NumPy namespace is then
    replaced by `melvin.b200` (with a warning) and everything still runs on the device."""
    if getattr(xp, "__name__", "") != "numpy":
        return xp
    import os
    if os.environ.get("MELVIN_B200_NUMPY_IS_DEVICE", "") not in ("", "0"):
        import warnings
        from . import b200
        warnings.warn("melvin-b200: xp = numpy requested, computing on the device (melvin.b200); "
                      "there is no CPU path", UserWarning, stacklevel=3)
        return b200
    raise _backend.BackendUnavailable(
        "melvin-b200 has no CPU path: pass the device namespace (`from melvin import b200 as xp`, "
        "or the `cupy` shim) instead of numpy (or set MELVIN_B200_NUMPY_IS_DEVICE=1 to run an "
        "`xp = np` script on the device)")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _dev(x, dtype=None):
    """Host arrays are uploaded; device arrays pass through (materialised)."""
    if isinstance(x, (SpecExpr, LazyLap)):
        return x.materialize()
    if isinstance(x, DeviceArray):
        return x
    return DeviceArray(_backend.from_host(np.asarray(x), dtype))


def _contig(a):
    """Contiguous tensor holding `a` (strided views are packed by mlv_elementwise)."""
    t = a._touch()._t
    if t.is_contiguous():
        return t
    packed = DeviceArray(_backend.empty(tuple(t.shape), a.dtype))
    packed[...] = a
    return packed._t


class ArrayFactory:
    """Creates arrays of the correct size and shape (reference melvin/ArrayFactory.py)."""

    def __init__(self, params, xp):
        xp = _require_device_namespace(xp)
        self._p = params
        self._xp = xp

    def make_mode_number_matrices(self):
        """Integer mode numbers n, m (ArrayFactory.py:8-26).  Host metadata: the
        kernels derive n and m from the element index and never read these."""
        p = self._p
        if p.is_fully_spectral():
            n = np.concatenate((np.arange(0, p.nn + 1), np.arange(-p.nn, 0)))
            m = np.arange(0, p.nm)
        elif p.discretisation[0] == "fdm":
            n, m = np.arange(0, p.nx), np.arange(0, p.nm)
        else:
            n, m = np.arange(0, p.nn), np.arange(0, p.nz)
        return np.meshgrid(n, m, indexing="ij")

    def _ctx(self):
        return _backend.context_for(self._p)

    def make_spectral(self, ni=None, nj=None):
        if ni is None and nj is None and self._p.is_fully_spectral() and self._ctx().world > 1:
            ctx = self._ctx()                         # this rank's kz-slab
            out = self._xp.zeros(ctx.spec_shape, dtype=self._p.complex)
            out._dist = Dist(ctx, "cols", 1, 0, ctx.nm)
            return out
        ni = self._p.spectral_shape[0] if ni is None else ni
        nj = self._p.spectral_shape[1] if nj is None else nj
        return self._xp.zeros((ni, nj), dtype=self._p.complex)

    def make_physical(self, nx=None, nz=None):
        if nx is None and nz is None and self._p.is_fully_spectral() and self._ctx().world > 1:
            ctx = self._ctx()                         # this rank's x-slab
            out = self._xp.zeros(ctx.phys_shape, dtype=self._p.float)
            out._dist = Dist(ctx, "rows", 0, 0, ctx.nx)
            return out
        nx = self._p.nx if nx is None else nx
        nz = self._p.nz if nz is None else nz
        return self._xp.zeros((nx, nz), dtype=self._p.float)


class SpectralTransformer:
    """Physical <-> spectral transforms with 2/3-rule truncation
    (reference melvin/SpectralTransformer.py)."""

    def __init__(self, params, xp, array_factory):
        xp = _require_device_namespace(xp)
        self._p = params
        self._xp = xp
        self._array_factory = array_factory
        self._ctx = _backend.context_for(params)
        if params.is_fully_spectral():
            self.to_physical = self.__to_physical_2d
            self.to_spectral = self.__to_spectral_2d
            self._scale = self._scale_2d
        else:
            self.to_physical = self.__to_physical_1d
            self.to_spectral = self.__to_spectral_1d
            self._scale = self._scale_1d

    # The reference's third slot (SpectralTransformer.py:21-31): copy between the truncated spectrum and
    # the full FFT layout.  The kernels prune / pad inside the transforms and never call it; kept for
    # code that does (plain slice copies on either kind of array).
    def _scale_2d(self, in_arr, out):
        nn, nm = self._p.nn, self._p.nm
        for rows in (slice(None, nn + 1), slice(-nn, None)):
            out[rows, :nm] = in_arr[rows, :nm]

    def _scale_1d(self, in_arr, out, axis):
        if axis == 0:
            out[:self._p.nn] = in_arr[:self._p.nn]
        elif axis == 1:
            out[:, :self._p.nm] = in_arr[:, :self._p.nm]

    _DEFAULT = [_CE, _CE]

    @staticmethod
    def _only_fourier(basis_functions):
        """True for the all-Fourier pair (the register-transform kernels); COSINE / SINE axes go
        through mlv_trig_axis."""
        for b in basis_functions:
            if b not in (_CE, _COS, _SIN):
                raise NotImplementedError("fully spectral transforms take COMPLEX_EXP, COSINE or SINE bases")
        return basis_functions[0] is _CE and basis_functions[1] is _CE

    # -- COSINE / SINE axes (SpectralTransformer.py:108-125,134-146,169-196): the reference mirrors
    #    the field to period 2(n-1) and calls rfft2 / irfft2; mlv_trig_axis evaluates the retained
    #    modes (or the n samples) of that transform directly, one axis per call, in rfft2's order
    def _trig_axes(self, basis_functions):
        p = self._p
        ax = []
        for b, n in ((basis_functions[0], p.nx), (basis_functions[1], p.nz)):
            if b is _CE:
                ax.append((_capi.EXT_PERIODIC, n, complex(n), 1.0))
            elif b is _COS:
                ax.append((_capi.EXT_EVEN, 2 * (n - 1), complex(n - 1), 0.5))
            else:
                ax.append((_capi.EXT_ODD, 2 * (n - 1), -1j * (n - 1), 1.0))
        return ax                     # per axis: (extension, period, factor, forward weight of mode 0)

    def _trig_check(self):
        if self._ctx.world > 1:
            raise NotImplementedError("COSINE / SINE bases are not slab-decomposed")

    def _to_spectral_trig(self, in_arr, out, basis_functions):
        self._trig_check()
        p, ctx = self._p, self._ctx
        if out is None:
            out = self._array_factory.make_spectral()
        src = _contig(_dev(in_arr, np.float64))
        if src.is_complex() or tuple(src.shape) != (p.nx, p.nz):
            raise TypeError("to_spectral expects a real (nx, nz) physical array")
        if not isinstance(out, DeviceArray) or not out._touch()._t.is_contiguous() \
                or tuple(out.shape) != tuple(p.spectral_shape):
            raise TypeError("out must be a contiguous spectral-shaped device array")
        out._pre_write()
        (ex, mx, fx, wx), (ez, mz, fz, wz) = self._trig_axes(basis_functions)
        nm, rows = p.nm, 2 * p.nn + 1
        tmp = _backend.empty((p.nx, nm), np.complex128)
        d = _capi.Trig()                     # z axis: rows of the physical field -> one-sided modes
        d.inverse, d.ext, d.period, d.n_samp, d.n_modes = 0, ez, mz, p.nz, nm
        d.two_sided, d.hermitian, d.samp_complex, d.batch_fastest, d.nbatch = 0, 0, 0, 0, p.nx
        d.samp_stride, d.samp_batch_stride, d.mode_stride, d.mode_batch_stride = 1, p.nz, 1, nm
        d.in_, d.out, d.scale_re, d.scale_im, d.w0 = src.data_ptr(), tmp.data_ptr(), 1.0, 0.0, wz
        ctx.call("mlv_trig_axis", ctypes.byref(d))
        scale = 1.0 / (fx * fz)              # SpectralTransformer.py:191
        d = _capi.Trig()                     # x axis: columns -> modes 0..nn,-nn..-1
        d.inverse, d.ext, d.period, d.n_samp, d.n_modes = 0, ex, mx, p.nx, rows
        d.two_sided, d.hermitian, d.samp_complex, d.batch_fastest, d.nbatch = 1, 0, 1, 1, nm
        d.samp_stride, d.samp_batch_stride, d.mode_stride, d.mode_batch_stride = nm, 1, nm, 1
        d.in_, d.out, d.scale_re, d.scale_im, d.w0 = tmp.data_ptr(), out._t.data_ptr(), scale.real, scale.imag, wx
        ctx.call("mlv_trig_axis", ctypes.byref(d))
        return out

    def _to_physical_trig(self, in_arr, out, basis_functions):
        """The reference doubles the mean mode of a cosine axis inside the caller's array
        (SpectralTransformer.py:123-126, so a second call sees doubled data); here the weight is
        applied on the fly and `in_arr` is left untouched."""
        self._trig_check()
        p, ctx = self._p, self._ctx
        if out is None:
            out = self._array_factory.make_physical()
        src = _contig(_dev(in_arr, np.complex128))
        if tuple(src.shape) != tuple(p.spectral_shape):
            raise TypeError("to_physical expects a spectral-shaped array")
        dst = out._t if isinstance(out, DeviceArray) else None
        if dst is None or not dst.is_contiguous() or dst.is_complex() or tuple(dst.shape) != (p.nx, p.nz):
            raise TypeError("out must be a contiguous real (nx, nz) device array")
        out._pre_write()
        (ex, mx, fx, wx), (ez, mz, fz, wz) = self._trig_axes(basis_functions)
        nm, rows = p.nm, 2 * p.nn + 1
        tmp = _backend.empty((p.nx, nm), np.complex128)
        scale = fx * fz / (mx * mz)          # :131 and the 1/(Mx Mz) of irfft2
        d = _capi.Trig()                     # x axis: modes -> the first nx samples of the period
        d.inverse, d.ext, d.period, d.n_samp, d.n_modes = 1, ex, mx, p.nx, rows
        d.two_sided, d.hermitian, d.samp_complex, d.batch_fastest, d.nbatch = 1, 0, 1, 1, nm
        d.samp_stride, d.samp_batch_stride, d.mode_stride, d.mode_batch_stride = nm, 1, nm, 1
        d.in_, d.out, d.scale_re, d.scale_im, d.w0 = src.data_ptr(), tmp.data_ptr(), scale.real, scale.imag, 1.0 / wx
        ctx.call("mlv_trig_axis", ctypes.byref(d))
        d = _capi.Trig()                     # z axis: one-sided modes -> real samples (irfft)
        d.inverse, d.ext, d.period, d.n_samp, d.n_modes = 1, ez, mz, p.nz, nm
        d.two_sided, d.hermitian, d.samp_complex, d.batch_fastest, d.nbatch = 0, 1, 0, 0, p.nx
        d.samp_stride, d.samp_batch_stride, d.mode_stride, d.mode_batch_stride = 1, p.nz, 1, nm
        d.in_, d.out, d.scale_re, d.scale_im, d.w0 = tmp.data_ptr(), dst.data_ptr(), 1.0, 0.0, 1.0 / wz
        ctx.call("mlv_trig_axis", ctypes.byref(d))
        return out

    @staticmethod
    def _fdm_axis(basis_functions):
        if basis_functions[0] is _FDM:
            raise NotImplementedError("Finite difference not implemented in x direction")
        if basis_functions[1] is not _FDM:
            raise Exception("One basis function must be FDM")

    def __to_physical_2d(self, in_arr, out=None, basis_functions=_DEFAULT):
        """SpectralTransformer.py:90-150"""
        if not self._only_fourier(basis_functions):
            return self._to_physical_trig(in_arr, out, basis_functions)
        if out is None:
            out = self._array_factory.make_physical()
        ctx = self._ctx
        if ctx.world > 1:
            in_arr = self._local(in_arr, ctx, spectral=True)
        src = _contig(_dev(in_arr, np.complex128))
        dst = out._t if isinstance(out, DeviceArray) else None
        if dst is None or not dst.is_contiguous():
            raise TypeError("out must be a contiguous device array")
        out._pre_write()
        if ctx.world > 1:
            # x pass on the local columns -> row block h to rank h -> z pass on the local rows
            send = ctx.take_i()
            ctx.call("mlv_x_inverse", 1, (ctypes.c_void_p * 1)(src.data_ptr()),
                     (ctypes.c_int32 * 1)(_capi.OP_IDENT), (ctypes.c_void_p * 1)(send.data_ptr()))
            recv = ctx.exchange(send, False)
            ctx.call("mlv_z_inverse", _ptr(recv), _ptr(dst))
            ctx.give_i(send)
            ctx.give_i(recv)
            return out
        ctx.call("mlv_to_physical", _ptr(src), _ptr(ctx.scratch_i()), _ptr(dst))
        return out

    @staticmethod
    def _local(arr, ctx, spectral):
        """Slab-decomposed runs: a full (global-shaped) host or device array is cut to this
        rank's slab; arrays that are slabs already pass through."""
        if getattr(arr, "_dist", None) is not None or isinstance(arr, (SpecExpr, LazyLap)):
            return arr
        if spectral and tuple(arr.shape) == tuple(ctx.global_spec_shape):
            out = DeviceArray(_backend.zeros(ctx.spec_shape, np.complex128))
            out._dist = Dist(ctx, "cols", 1, 0, ctx.nm)
            out[:, :] = arr
            return out
        if not spectral and tuple(arr.shape) == (ctx.nx, ctx.nz):
            return arr[ctx.x_off:ctx.x_off + ctx.nxl]
        return arr

    def __to_spectral_2d(self, in_arr, out=None, basis_functions=_DEFAULT):
        """SpectralTransformer.py:152-199"""
        if not self._only_fourier(basis_functions):
            return self._to_spectral_trig(in_arr, out, basis_functions)
        if out is None:
            out = self._array_factory.make_spectral()
        ctx = self._ctx
        if ctx.world > 1:
            in_arr = self._local(in_arr, ctx, spectral=False)
        src = _contig(_dev(in_arr, np.float64))
        if src.is_complex():
            raise TypeError("to_spectral expects a real physical array")
        out._touch()
        out._pre_write()
        if ctx.world > 1:
            # z pass on the local rows -> tile block h to rank h -> x pass on the local columns
            send = ctx.take_i()
            ctx.call("mlv_z_forward", _ptr(src), _ptr(send))
            recv = ctx.exchange(send, True)
            d = _capi.XFwd()
            d.nf, d.mode = 1, 0
            d.src[0], d.sym[0], d.coef[0], d.dst = recv.data_ptr(), _capi.SYM_ONE, 1.0, out._t.data_ptr()
            ctx.call("mlv_x_forward", ctypes.byref(d))
            ctx.give_i(send)
            ctx.give_i(recv)
            return out
        ctx.call("mlv_to_spectral", _ptr(src), _ptr(ctx.scratch_i()), _ptr(out._t))
        return out

    def __to_physical_1d(self, in_arr, out=None, basis_functions=_DEFAULT):
        """SpectralTransformer.py:33-61"""
        self._fdm_axis(basis_functions)
        if out is None:
            out = self._array_factory.make_physical()
        src = _contig(_dev(in_arr, np.complex128))
        out._pre_write()
        self._ctx.call("mlv_to_physical", _ptr(src), None, _ptr(out._t))
        return out

    def __to_spectral_1d(self, in_arr, out=None, basis_functions=_DEFAULT):
        """SpectralTransformer.py:63-88"""
        self._fdm_axis(basis_functions)
        if out is None:
            out = self._array_factory.make_spectral()
        src = _contig(_dev(in_arr, np.float64))
        out._touch()
        out._pre_write()
        self._ctx.call("mlv_to_spectral", _ptr(src), None, _ptr(out._t))
        return out


class SpatialDifferentiator:
    """Spectral derivative symbols and physical-space central differences
    (reference melvin/SpatialDifferentiator.py)."""

    def __init__(self, params, xp, array_factory=None):
        xp = _require_device_namespace(xp)
        self._xp = xp
        self._params = params
        self._ctx = _backend.context_for(params)
        self._x_periodic = params.discretisation[0] == "spectral"
        self._z_periodic = params.discretisation[1] == "spectral"
        if params.spatial_derivative_order not in (2, 4):
            raise NotImplementedError("spatial_derivative_order must be 2 or 4")
        self._order = params.spatial_derivative_order
        self._n, self._m = array_factory.make_mode_number_matrices()
        if params.discretisation[0] == "fdm":
            raise NotImplementedError("Finite difference not implemented in x direction")
        self.sddx = self.__s_ddx
        self.sd2dx2 = self.__s_d2dx2
        if params.discretisation[1] == "fdm":
            self.sddz = lambda var, bs: self.pddz(var)
            self.sd2dz2 = self.__s_d2dz2_fdm
        else:
            self.sddz = self.__s_ddz
            self.sd2dz2 = self.__s_d2dz2

    # -- spectral symbols (SpatialDifferentiator.py:50-74): deferred terms
    def _term(self, var, basis_fn, op):
        if basis_fn is _FDM:
            return 0.0 * _dev(var)          # diff factor of FDM is 0 (BasisFunctions.py:26-36)
        if basis_fn is not _CE:
            # diff factors of the trigonometric bases (BasisFunctions.py:26-48): -pi/L (COSINE),
            # +pi/L (SINE), second derivative -pi^2/L^2 -- exact multiples (powers of two, times
            # +-i) of the Fourier symbols i 2pi/L and -(2pi/L)^2 the kernels evaluate
            first = op in (_capi.OP_DDX, _capi.OP_DDZ)
            c = (0.5j if basis_fn is _COS else -0.5j) if first else 0.25
            return c * self._term(var, _CE, op)
        if isinstance(var, SpecExpr) and not var.nls and len(var.terms) == 1 \
                and var.terms[0][1] == _capi.OP_IDENT:
            c, _, a = var.terms[0]
            return SpecExpr(self._ctx, [(c, op, a)])
        a = _dev(var, np.complex128)
        lifted = SpecExpr._lift(self._ctx, a)
        if lifted is None:
            raise TypeError("spectral derivative expects a contiguous complex spectral-shaped array")
        return SpecExpr(self._ctx, [(1.0 + 0j, op, a)])

    def __s_ddx(self, var, basis_fn):
        return self._term(var, basis_fn, _capi.OP_DDX)

    def __s_ddz(self, var, basis_fn):
        return self._term(var, basis_fn, _capi.OP_DDZ)

    def __s_d2dx2(self, var, basis_fn):
        return self._term(var, basis_fn, _capi.OP_D2DX2)

    def __s_d2dz2(self, var, basis_fn):
        return self._term(var, basis_fn, _capi.OP_D2DZ2)

    def __s_d2dz2_fdm(self, var, basis_fn):
        """FDM-z: the second z difference of a spectral array (SpatialDifferentiator.py:36-40,
        106-128) as a deferred row-stencil term; other operands go through the eager stencil."""
        a = var if isinstance(var, DeviceArray) else None
        if a is not None and SpecExpr._lift(self._ctx, a) is not None:
            return SpecExpr(self._ctx, [(1.0 + 0j, _capi.OP_FDM_D2DZ2, a)])
        return self.pd2dz2(var)

    def calc_lap(self, basis_fns):
        """SpatialDifferentiator.py:70-74"""
        if basis_fns[0] is _CE and basis_fns[1] is _CE and self._z_periodic:
            return LazyLap(self._ctx, 1.0)
        from .basis import gen_diff2_factors
        fx = gen_diff2_factors(self._params.lx)[basis_fns[0]]
        fz = gen_diff2_factors(self._params.lz)[basis_fns[1]]
        return DeviceArray(_backend.from_host(np.asarray(fx * self._n ** 2 + fz * self._m ** 2,
                                                         dtype=np.float64)))

    # -- physical stencils (SpatialDifferentiator.py:76-185)
    def _stencil(self, var, out, axis, periodic, second, h):
        a = _dev(var)
        src = _contig(a)
        if src.dim() != 2:
            raise ValueError("stencils operate on 2-D arrays")
        ncomp = 2 if src.is_complex() else 1
        if out is None:
            if not second:
                # the reference allocates a real (nx, nz) buffer here; a complex or
                # differently shaped operand fails on assignment
                if ncomp == 2 or tuple(src.shape) != (self._params.nx, self._params.nz):
                    raise ValueError("could not broadcast input array into a real (nx, nz) output")
                out = self._xp.zeros((self._params.nx, self._params.nz), dtype=np.float64)
            else:
                out = self._xp.zeros_like(a)
        dst = out._touch()._t
        out._pre_write()
        if not dst.is_contiguous() or tuple(dst.shape) != tuple(src.shape) \
                or dst.is_complex() != src.is_complex():
            raise ValueError("out must be contiguous and match the operand")
        self._ctx.call("mlv_stencil", _ptr(src), _ptr(dst), int(src.shape[0]), int(src.shape[1]),
                       ncomp, axis, self._order, int(periodic), int(second), ctypes.c_double(h))
        return out

    def pddx(self, var, out=None):
        return self._stencil(var, out, 0, self._x_periodic, 0, self._params.dx)

    def pddz(self, var, out=None):
        return self._stencil(var, out, 1, self._z_periodic, 0, self._params.dz)

    def pd2dz2(self, var, out=None):
        return self._stencil(var, out, 1, False, 1, self._params.dz)


class LaplacianSolver:
    """Solves lap(psi) = rhs (reference melvin/LaplacianSolver.py)."""

    def __init__(self, params, xp, basis_fns, spatial_diff=None, array_factory=None):
        xp = _require_device_namespace(xp)
        self._params = params
        self._xp = xp
        self._array_factory = array_factory
        self._ctx = _backend.context_for(params)
        if params.is_fully_spectral():
            self._lap = spatial_diff.calc_lap(basis_fns)
            self.solve = self._solve_fully_spectral
        else:
            print("WARNING: Laplacian solver only implemented for boundary conditions "
                  "where soln matches value of rhs")
            self.solve = self._solve_fdm
            # EXTENSION (no reference code path): `laplacian_order: 4` in the parameters selects
            # the pentadiagonal 4th-order operator; the default is the reference's tridiagonal one
            self._order = int(getattr(params, "laplacian_order", 2))
            if self._order not in (2, 4):
                raise NotImplementedError("laplacian_order must be 2 or 4")

    @property
    def lap(self):
        return self._lap.materialize() if isinstance(self._lap, LazyLap) else self._lap

    @property
    def laps(self):
        """Host copies of the nn tridiagonal matrices (LaplacianSolver.py:25-49);
        the device solve uses their Thomas factors held by the context."""
        import scipy.sparse as sp
        p = self._params
        kx0 = abs(1j * 2 * np.pi / p.lx)
        mats = []
        if getattr(self, "_order", 2) == 4:
            h2 = p.dz ** 2
            for n in range(p.nn):
                k2 = (n * kx0) ** 2
                m = sp.lil_matrix((p.nz, p.nz), dtype=np.complex128)
                for i in range(p.nz):
                    if i in (0, p.nz - 1):
                        m[i, i] = 1.0
                    elif i in (1, p.nz - 2):
                        m[i, i - 1], m[i, i], m[i, i + 1] = 1.0 / h2, -2.0 / h2 - k2, 1.0 / h2
                    else:
                        m[i, i - 2] = m[i, i + 2] = -1.0 / (12.0 * h2)
                        m[i, i - 1] = m[i, i + 1] = 4.0 / (3.0 * h2)
                        m[i, i] = -5.0 / (2.0 * h2) - k2
                mats.append(m.tocsr())
            return mats
        for n in range(p.nn):
            diag = np.full(p.nz, -((n * kx0) ** 2 + 2.0 / p.dz ** 2))
            off = np.full(p.nz, 1.0 / p.dz ** 2)
            m = sp.dia_matrix((np.array([off, diag, off]), np.array([-1, 0, 1])),
                              shape=(p.nz, p.nz), dtype=np.complex128).tolil()
            m[0, 0], m[0, 1], m[-1, -1], m[-1, -2] = 1.0, 0.0, 1.0, 0.0
            mats.append(m.tocsr())
        return mats

    def _solve_fully_spectral(self, rhs, out=None):
        """rhs / lap with lap[0,0] := 1 (LaplacianSolver.py:58-68)"""
        if out is None:
            out = self._array_factory.make_spectral()
        src = _contig(_dev(rhs, np.complex128))
        out._touch()
        out._pre_write()
        lt = _capi.make_lin_terms([(1.0, _capi.OP_INVLAP, src.data_ptr())])
        self._ctx.call("mlv_spec_lincomb", ctypes.byref(lt), _ptr(out._t))
        return out

    def _solve_fdm(self, rhs, out=None):
        """nn tridiagonal Dirichlet systems (LaplacianSolver.py:70-79)"""
        if out is None:
            out = self._array_factory.make_spectral()
        src = _contig(_dev(rhs, np.complex128))
        out._touch()
        out._pre_write()
        self._ctx.call("mlv_solve_fdm_o4" if self._order == 4 else "mlv_solve_fdm", _ptr(src), _ptr(out._t))
        return out


class Integrator:
    """Adams-Bashforth predictor with explicit or semi-implicit (theta-scheme)
    treatment of the linear term (reference melvin/Integrator.py)."""

    def __init__(self, params, xp):
        xp = _require_device_namespace(xp)
        self._dt = params.initial_dt
        self._dx = params.dx
        self._dz = params.dz
        self._cfl_cutoff = params.cfl_cutoff
        self._xp = xp
        self._order = params.integrator_order
        self._ctx = _backend.context_for(params)
        if params.integrator_order == 2:
            self.predictor = self._adams_bashforth_2
            self.corrector = self._adams_moulton_2
        elif params.integrator_order == 4:
            self.predictor = self._adams_bashforth_4
            self.corrector = self._adams_moulton_4
        if params.integrator == "semi-implicit":
            if params.is_fully_spectral():
                self.integrate = self._semi_implicit_spectral
                self._alpha = params.alpha
        elif params.integrator == "explicit":
            self.integrate = self._explicit
            self._alpha = params.alpha

    # -- stand-alone predictor / corrector formulas (Integrator.py:5-33); the fused
    #    kernels evaluate the predictor themselves, these exist for API parity
    def _adams_bashforth_2(self, dvar):
        return self._dt / 2 * (3 * dvar.get() - dvar.get(-1))

    def _adams_bashforth_4(self, dvar):
        return self._dt / 24 * (55 * dvar.get() - 59 * dvar.get(-1)
                                + 37 * dvar.get(-2) - 9 * dvar.get(-3))

    def _adams_moulton_2(self, dvar):
        return self._dt / 2 * (dvar.get() + dvar.get(-1))

    def _adams_moulton_4(self, dvar):
        return self._dt / 24 * (9 * dvar.get() + 19 * dvar.get(-1)
                                - 5 * dvar.get(-2) + dvar.get(-3))

    def predictor_corrector(self, var, dvar, rhs, diffusion_term=None):
        """One PECE step of the Adams-Bashforth / Adams-Moulton pair the reference advertises
        (README.md:55) but never wires up (`corrector` has no call site):

            predict   y* = y_n + AB(f_n, f_n-1, ...)            (Integrator.py:5-18)
            evaluate  f* = rhs(y*)                               (the script's right-hand side)
            correct   y_n+1 = y_n + AM(f*, f_n, f_n-1, ...)      (Integrator.py:20-33)

        `dvar` holds f_n in its current level on entry (as for `integrate`) and f* on exit,
        `rhs` is a callable returning the right-hand side for the state now in `var`
        (explicit treatment of every term; `diffusion_term`, if given, is added to f_n first as
        `_explicit` does).  An extension without a reference implementation: parity is
        unpinned, its order of accuracy is verified by convergence tests."""
        dvar._flush()
        if diffusion_term is not None:
            dvar[:] = dvar[:] + _dev(diffusion_term, np.complex128)
        y_n = _dev(var[:], np.complex128).copy()
        var[:] = y_n + self.predictor(dvar)                      # P
        dvar.advance()
        dvar[:] = rhs()                                          # E (at the predicted state)
        dvar._flush()
        var[:] = y_n + self.corrector(dvar)                      # C

    def set_dt(self, ux, uz):
        """Sets dt based on CFL limit (Integrator.py:35-44; signed max)."""
        mx = ux._cached_reduction(1) if hasattr(ux, "_cached_reduction") else None
        mz = uz._cached_reduction(1) if hasattr(uz, "_cached_reduction") else None
        if mx is None:
            mx = self._xp.max(ux.getp())
        if mz is None:
            mz = self._xp.max(uz.getp())
        with np.errstate(divide="ignore", invalid="ignore"):
            cfl_dt = min(np.float64(self._dx) / np.float64(mx), np.float64(self._dz) / np.float64(mz))
        if self._dt > cfl_dt or np.isnan(cfl_dt):
            raise Exception("CFL condition breached")
        while self._dt > self._cfl_cutoff * cfl_dt:
            self._dt = self._dt * 0.9

    def override_dt(self, dt):
        """Manually set dt"""
        self._dt = dt

    def get_dt(self):
        return self._dt

    # -- the update
    def _launch(self, var, dvar, scheme, lcoef, larr, extra):
        """f0 = pending RHS (+ extra terms); var <- update; dvar.advance()."""
        from .fields import Variable
        ctx = self._ctx
        order = self._order
        pending, dvar._pending = dvar._pending, None
        f0 = dvar._level(0)._t
        older = [dvar._level(-k)._t for k in range(1, order)]
        is_var = isinstance(var, Variable)
        if is_var:
            var._materialize_s()
            q_in = var._s._t
            # somebody (psi/ux/uz, a requested transform, another deferred right-hand side)
            # still reads the old state: write the new one into the second buffer
            double = var._has_dependants(q_in, exclude=pending)
            # a row stencil of the state itself reads neighbours other threads are updating
            stencil = any(op >= _capi.OP_FDM_D2DZ2 and op != _capi.OP_FDX_SYM and a._t.data_ptr() == q_in.data_ptr()
                          for _, op, a in (list(pending.terms) if pending is not None else []) + list(extra))
            double = double or stencil
            if double:
                if var._s_spare is None:
                    var._s_spare = _backend.empty(tuple(q_in.shape), np.complex128)
                q_out = var._s_spare
                var._flush_dependants(q_out)
            else:
                q_out = q_in
        else:
            double = False
            var._pre_write()
            q_in = q_out = var._touch()._t
        g = _capi.Integ()
        g.ab_order, g.scheme = order, scheme
        g.dt, g.alpha, g.lcoef = float(self._dt), float(getattr(self, "_alpha", 0.0)), float(lcoef)
        g.larr = larr._t.data_ptr() if larr is not None else None
        g.q_in, g.q_out = q_in.data_ptr(), q_out.data_ptr()
        g.f0, g.fm1 = f0.data_ptr(), older[0].data_ptr()
        if order == 4:
            g.fm2, g.fm3 = older[1].data_ptr(), older[2].data_ptr()
        keep = [larr]
        lin = list(extra)
        fused = (not ctx.fdm_z and pending is not None and 1 <= len(pending.nls) <= 2
                 and len(pending.terms) + len(lin) <= 4)
        fdm_terms = None
        if ctx.fdm_z and pending is not None:
            fdm_terms = [t for coef, nl in pending.nls for t in nl.lin_terms(coef)] + list(pending.terms) + lin
            if len(fdm_terms) > _capi.MAXLIN:
                fdm_terms = None
        if fdm_terms is not None:
            # K3: the whole right-hand side (row stencils included), the history write and the
            # update in one row-wise kernel
            g.f0_set = 1
            lt = _capi.make_lin_terms([(c, op, a._touch()._t.data_ptr()) for c, op, a in fdm_terms])
            ctx.call("mlv_integrate", ctypes.byref(lt), ctypes.byref(g))
        elif fused:
            d = _capi.XFwd()
            d.nf, d.mode = 2 * len(pending.nls), 1
            for i, (coef, nl) in enumerate(pending.nls):
                d.src[2 * i], d.src[2 * i + 1] = nl.ia.data_ptr(), nl.ib.data_ptr()
                d.sym[2 * i], d.sym[2 * i + 1] = _capi.SYM_FDX, _capi.SYM_FDZ
                d.coef[2 * i] = d.coef[2 * i + 1] = coef
            lin = list(pending.terms) + lin
            d.lin = _capi.make_lin_terms([(c, op, a._touch()._t.data_ptr()) for c, op, a in lin])
            d.integ = g
            ctx.call("mlv_x_forward", ctypes.byref(d))
        else:
            if pending is not None:
                pending.materialize(out=dvar._level(0))
            lt = _capi.make_lin_terms([(c, op, a._touch()._t.data_ptr()) for c, op, a in lin])
            ctx.call("mlv_integrate", ctypes.byref(lt) if lin else None, ctypes.byref(g))
        del keep
        if double:
            var._s._t, var._s_spare = q_out, q_in       # the handle object never changes
        dvar.advance()

    def _explicit(self, var, dvar, diffusion_term):
        """dvar += diffusion; var += AB(dvar); advance (Integrator.py:53-56)"""
        if isinstance(diffusion_term, SpecExpr) and not diffusion_term.nls \
                and len(diffusion_term.terms) <= 4:
            extra = list(diffusion_term.terms)
        else:
            extra = [(1.0 + 0j, _capi.OP_IDENT, _dev(diffusion_term, np.complex128))]
        pend = dvar._pending
        if pend is not None and not self._ctx.fdm_z and (len(pend.terms) + len(extra) > 4 or not pend.nls):
            dvar._flush()
        self._launch(var, dvar, _capi.SCHEME_EXPLICIT, 0.0, None, extra)

    def _semi_implicit_spectral(self, var, dvar, lin_op):
        """theta-scheme with the AB predictor (Integrator.py:58-63)"""
        if isinstance(lin_op, LazyLap):
            self._launch(var, dvar, _capi.SCHEME_SI_LAP, lin_op.coef, None, [])
        else:
            larr = _dev(lin_op, np.float64)
            if larr.is_complex or not larr._t.is_contiguous():
                raise TypeError("lin_op must be a real contiguous spectral-shaped array")
            self._launch(var, dvar, _capi.SCHEME_SI_ARR, 0.0, larr, [])
