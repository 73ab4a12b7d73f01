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
    shape = var.getp().shape
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
