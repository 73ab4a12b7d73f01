"""Helper functions used by the example scripts (reference melvin/utility.py)."""
import numpy as np
from numpy.random import default_rng

from . import _capi
from .basis import BasisFunctions


def sech(x):
    return 1.0 / np.cosh(x)


def load_scipy_sparse(xp):
    """Kept for import compatibility (melvin/utility.py:13-19); the device
    Laplacian solve does not go through scipy/cupyx."""
    import scipy.sparse
    return scipy.sparse


def load_scipy_sparse_linalg(xp):
    import scipy.sparse.linalg
    return scipy.sparse.linalg


def init_var_with_noise(var, epsilon, seed=0):
    """Uniform noise in physical space (melvin/utility.py:31-39)."""
    rng = default_rng(seed)
    shape = tuple(var._params.physical_shape)      # the global field (slab runs keep their rows in load)
    data_p = np.zeros(shape)
    data_p += epsilon * (2 * rng.random(shape) - 1.0)
    var.load(data_p, is_physical=True)


def calc_kinetic_energy(ux, uz, xp, params):
    """0.5 * sum(uz^2 + ux^2) / (nx nz)   (melvin/utility.py:42-59).

    The fused physical-space stage already reduced ux^2 and uz^2 while the
    velocities were on chip; otherwise the fields are materialised and reduced."""
    nx, nz = params.nx, params.nz
    sx = ux._cached_reduction(2) if hasattr(ux, "_cached_reduction") else None
    sz = uz._cached_reduction(2) if hasattr(uz, "_cached_reduction") else None
    if sx is None or sz is None:
        from .operators import _require_device_namespace
        xp = _require_device_namespace(xp)
    if sx is None:
        sx = xp.sum_of_squares(ux.getp())
    if sz is None:
        sz = xp.sum_of_squares(uz.getp())
    return 0.5 * (sz + sx) / (nx * nz)


def calc_velocity_from_vorticity(vorticity, streamfunction, ux, uz, laplacian_solver):
    """psi = solve(-w); ux = -d(psi)/dz; uz = d(psi)/dx, each taken to physical space
    (melvin/utility.py:62-79).

    Fully spectral: psi, ux, uz become *deferred* diagonal operators of the
    vorticity buffer (MLV_OP_PSI / UX / UZ); they are evaluated inside the inverse
    x pass that the next vec_dot_nabla issues, and written out only if somebody
    reads them."""
    psi = streamfunction
    fused = (getattr(vorticity, "_fused", False) and getattr(psi, "_fused", False)
             and getattr(ux, "_fused", False) and getattr(uz, "_fused", False))
    if fused:
        src = vorticity.gets()
        psi._set_virtual(_capi.OP_PSI, src)
        ux._set_virtual(_capi.OP_UX, src)
        ux._request_physical()
        uz._set_virtual(_capi.OP_UZ, src)
        uz._request_physical()
        return

    if (all(getattr(v, "_fused_fdm", False) for v in (vorticity, psi, ux, uz))
            and getattr(laplacian_solver, "_order", 2) == 2):     # (the 4th-order solve is not fused)
        _velocity_fdm(vorticity, psi, ux, uz)
        return

    laplacian_solver.solve(-vorticity.gets(), out=psi._sdata)
    if psi._basis_functions[1] is BasisFunctions.FDM:
        psi.to_physical()
        ux.setp(-psi.pddz())
    else:
        ux[:] = -psi.sddz()
        ux.to_physical()
    if psi._basis_functions[0] is BasisFunctions.FDM:
        psi.to_physical()
        uz.setp(psi.pddx())
    else:
        uz[:] = psi.sddx()
        uz.to_physical()


def _velocity_fdm(vorticity, psi, ux, uz):
    """Fourier-x / FDM-z branch in ONE row-wise kernel (mlv_fdm_velocity): the batched
    tridiagonal solve psi = solve(-w), uz^ = (i kx n) psi written to uz's spectral data as the
    reference does, and ux^ = -pddz(psi) kept as a private x spectrum (the reference never
    writes ux._sdata in this mode; its z stencil acts on psi in physical
    space, which commutes with the x transform).  The three physical fields are deferred."""
    import ctypes
    from . import _backend
    from .b200 import DeviceArray
    from .fields import _I_NONE, _I_PENDING
    ctx, vp = vorticity._ctx, ctypes.c_void_p
    w_t = vorticity.gets()._touch()._t
    for v in (psi, ux, uz):                   # their pending physical requests are being replaced
        v._i_state, v._i_def, v._red = _I_NONE, None, None
    psi._virt = uz._virt = None
    psi._flush_dependants(psi._s._t)
    uz._flush_dependants(uz._s._t)
    if ux._hat is None:
        ux._hat = _backend.empty(ctx.spec_shape, np.complex128)
    else:
        ux._flush_dependants(ux._hat)
    ctx.call("mlv_fdm_velocity", vp(w_t.data_ptr()), vp(psi._s._t.data_ptr()),
             vp(ux._hat.data_ptr()), vp(uz._s._t.data_ptr()))
    # psi.to_physical(); ux.setp(-psi.pddz()); uz.to_physical()  (utility.py:67-79), all deferred.
    # psi's physical field is re-derived from the vorticity buffer if somebody asks: the scripts'
    # boundary fix-ups write psi's spectral data every step and must not force a transform
    psi._i_def, psi._i_state, psi._p_valid = (_capi.OP_PSI, DeviceArray(w_t)), _I_PENDING, False
    ux._i_def, ux._i_state, ux._p_valid = (_capi.OP_IDENT, DeviceArray(ux._hat)), _I_PENDING, False
    uz._i_def, uz._i_state, uz._p_valid = (_capi.OP_IDENT, DeviceArray(uz._s._t)), _I_PENDING, False
