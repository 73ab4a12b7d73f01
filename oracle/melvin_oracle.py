"""CPU oracle for the Melvin.py per-timestep pseudo-spectral hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` are allowed to use it, and only as
the checker / the timed CPU baseline -- never as a compute fallback.

This is a NumPy float64 *restatement* (stateless functions over plain arrays,
one ``Grid`` description object) of the algorithm the reference implements in
its classes.  Every function cites the reference ``file:line`` it follows
(paths are relative to the reference checkout, e.g. ``melvin/Variable.py``).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference (NumPy backend, float64), runs its operator sequence and stores the
results in ``tests/golden/*.npz``; ``tests/test_oracle_vs_golden.py`` checks
every function below against those vectors and against the analytic
known-answer tests of the reference's own suite (``test/*_test.py``).

Unpinned extensions (no reference code path exists): none are
implemented here.

Third-party arithmetic at the boundary: ``numpy.fft`` (pocketfft) -- the same
library the reference calls (``melvin/SpectralTransformer.py:58,85,132,191``).
The finite-difference Laplacian solve is restated as a Thomas recurrence
instead of SuperLU (``melvin/LaplacianSolver.py:53-55``); SURVEY F8 documents
that the two differ by rounding only (<=1e-12 at the grid sizes the tests use).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

TWO_PI = 2.0 * np.pi


# --------------------------------------------------------------------------
# grid description  (melvin/Parameters.py:63-92, melvin/ArrayFactory.py:8-62)
# --------------------------------------------------------------------------
@dataclass
class Grid:
    nx: int
    nz: int
    lx: float
    lz: float
    fdm_z: bool = False          # discretisation == ["spectral", "fdm"]
    fd_order: int = 2            # spatial_derivative_order
    int_order: int = 2           # integrator_order
    integrator: str = "semi-implicit"
    alpha: float = 0.51          # melvin/Parameters.py:13
    cfl_cutoff: float = 0.5      # melvin/Parameters.py:14
    nn: int = field(init=False)
    nm: int = field(init=False, default=-1)

    def __post_init__(self):
        # melvin/Parameters.py:67-70
        self.nn = (self.nx - 1) // 3
        if not self.fdm_z:
            self.nm = (self.nz - 1) // 3
        self.dx = self.lx / self.nx
        self.dz = self.lz / self.nz

    @property
    def spectral_shape(self):
        # melvin/Parameters.py:72-78
        if self.fdm_z:
            return (self.nn, self.nz)
        return (2 * self.nn + 1, self.nm)

    @property
    def physical_shape(self):
        return (self.nx, self.nz)

    def default_dt(self):
        # melvin/Parameters.py:91-92
        return 0.2 * min(self.dx, self.dz)


def mode_numbers(g: Grid):
    """Integer mode numbers as broadcastable column / row vectors.

    melvin/ArrayFactory.py:8-26 builds full meshgrids; a broadcastable pair
    gives identical arithmetic.
    """
    if g.fdm_z:
        n = np.arange(0, g.nn)
        m = np.arange(0, g.nz)
    else:
        n = np.concatenate((np.arange(0, g.nn + 1), np.arange(-g.nn, 0)))
        m = np.arange(0, g.nm)
    return n[:, None], m[None, :]


# --------------------------------------------------------------------------
# transforms  (melvin/SpectralTransformer.py)
# --------------------------------------------------------------------------
def to_physical(g: Grid, spec: np.ndarray) -> np.ndarray:
    """Spectral -> physical, complex-exponential basis.

    Fully spectral: melvin/SpectralTransformer.py:90-150 with _scale_2d :21-24.
    FDM in z:       melvin/SpectralTransformer.py:33-61 with _scale_1d :26-31.
    """
    if g.fdm_z:
        padded = np.zeros((g.nx // 2 + 1, g.nz), dtype=np.complex128)
        padded[: g.nn] = spec[: g.nn]
        return np.fft.irfft(padded * g.nx, axis=0)
    padded = np.zeros((g.nx, g.nz // 2 + 1), dtype=np.complex128)
    padded[: g.nn + 1, : g.nm] = spec[: g.nn + 1, : g.nm]
    padded[-g.nn:, : g.nm] = spec[-g.nn:, : g.nm]
    padded *= g.nx * g.nz
    return np.fft.irfft2(padded)


def to_spectral(g: Grid, phys: np.ndarray) -> np.ndarray:
    """Physical -> truncated spectral.

    Fully spectral: melvin/SpectralTransformer.py:152-199.
    FDM in z:       melvin/SpectralTransformer.py:63-88.
    """
    out = np.zeros(g.spectral_shape, dtype=np.complex128)
    if g.fdm_z:
        full = np.fft.rfft(phys, axis=0) / g.nx
        out[: g.nn] = full[: g.nn]
        return out
    full = np.fft.rfft2(phys) / (g.nx * g.nz)
    out[: g.nn + 1, : g.nm] = full[: g.nn + 1, : g.nm]
    out[-g.nn:, : g.nm] = full[-g.nn:, : g.nm]
    return out


# COSINE / SINE bases.  Basis codes as in melvin/BasisFunctions.py:5-9.
COMPLEX_EXP, COSINE, SINE = 0, 1, 2


def to_spectral_basis(g: Grid, phys: np.ndarray, bx: int, bz: int) -> np.ndarray:
    """Physical -> truncated spectral for any pair of spectral bases
    (melvin/SpectralTransformer.py:152-199): even / odd mirror image, rfft2, scaling,
    halved mean mode of a cosine axis, 2/3-rule truncation (_scale_2d :21-24)."""
    arr = np.asarray(phys)
    xf, zf = g.nx, g.nz
    if bx == COSINE:
        arr = np.concatenate((arr[:-1], arr[:0:-1]))                   # :169-172
        xf = g.nx - 1
    elif bx == SINE:
        arr = np.concatenate((arr[:-1], -arr[:0:-1]))                  # :174-178
        xf = -1j * (g.nx - 1)
    if bz == COSINE:
        arr = np.concatenate((arr[:, :-1], arr[:, :0:-1]), axis=1)     # :180-184
        zf = g.nz - 1
    elif bz == SINE:
        arr = np.concatenate((arr[:, :-1], -arr[:, :0:-1]), axis=1)    # :185-189
        zf = -1j * (g.nz - 1)
    full = np.fft.rfft2(arr) / (xf * zf)                               # :191
    if bx == COSINE:
        full[0] /= 2                                                   # :193-194
    if bz == COSINE:
        full[:, 0] /= 2                                                # :195-196
    out = np.zeros(g.spectral_shape, dtype=np.complex128)
    out[: g.nn + 1, : g.nm] = full[: g.nn + 1, : g.nm]
    out[-g.nn:, : g.nm] = full[-g.nn:, : g.nm]
    return out


def to_physical_basis(g: Grid, spec: np.ndarray, bx: int, bz: int) -> np.ndarray:
    """Spectral -> physical for any pair of spectral bases (melvin/SpectralTransformer.py:90-150).
    The reference doubles the mean mode of a cosine axis IN the caller's array (:123-126); this
    restatement works on a copy (same output, no side effect)."""
    spec = np.array(spec, dtype=np.complex128, copy=True)
    xf, zf = g.nx, g.nz
    size = [g.nx, g.nz // 2 + 1]
    if bx == COSINE:
        size[0], xf = 2 * (g.nx - 1), g.nx - 1                         # :108-110
    elif bx == SINE:
        size[0], xf = 2 * (g.nx - 1), -1j * (g.nx - 1)                 # :111-113
    if bz == COSINE:
        size[1], zf = 2 * (g.nz - 1) // 2 + 1, g.nz - 1                # :115-117
    elif bz == SINE:
        size[1], zf = 2 * (g.nz - 1) // 2 + 1, -1j * (g.nz - 1)        # :118-120
    if bx == COSINE:
        spec[0] *= 2
    if bz == COSINE:
        spec[:, 0] *= 2
    padded = np.zeros(size, dtype=np.complex128)
    padded[: g.nn + 1, : g.nm] = spec[: g.nn + 1, : g.nm]
    padded[-g.nn:, : g.nm] = spec[-g.nn:, : g.nm]
    padded *= xf * zf                                                  # :131
    out = np.fft.irfft2(padded)                                        # :132
    if bx in (COSINE, SINE):
        out = out[: g.nx]                                              # :134-138
    if bz in (COSINE, SINE):
        out = out[:, : g.nz]                                           # :140-146
    return out


def diff_factor(basis: int, length: float):
    """melvin/BasisFunctions.py:26-48"""
    return {COMPLEX_EXP: 1j * 2 * np.pi, SINE: np.pi, COSINE: -np.pi}[basis] / length


def diff2_factor(basis: int, length: float):
    """melvin/BasisFunctions.py:39-42,57-61"""
    return -np.abs({COMPLEX_EXP: 1j * 2 * np.pi, SINE: np.pi, COSINE: -np.pi}[basis]) ** 2 / length ** 2


# --------------------------------------------------------------------------
# spectral derivative symbols  (melvin/SpatialDifferentiator.py:50-74,
#                               melvin/BasisFunctions.py:26-59)
# --------------------------------------------------------------------------
def _d1_factor(length):      # COMPLEX_EXP first-derivative factor
    return 1j * 2 * np.pi / length


def _d2_factor(length):      # COMPLEX_EXP second-derivative factor
    return -np.abs(1j * 2 * np.pi) ** 2 / length ** 2


def sddx(g: Grid, spec):
    n, _ = mode_numbers(g)
    return _d1_factor(g.lx) * n * spec          # SpatialDifferentiator.py:50-53


def sddz(g: Grid, spec):
    if g.fdm_z:
        raise NotImplementedError("reference raises here (SURVEY App. A-14)")
    _, m = mode_numbers(g)
    return _d1_factor(g.lz) * m * spec          # SpatialDifferentiator.py:55-58


def sd2dx2(g: Grid, spec):
    n, _ = mode_numbers(g)
    return _d2_factor(g.lx) * n ** 2 * spec     # SpatialDifferentiator.py:60-63


def sd2dz2(g: Grid, spec):
    if g.fdm_z:
        return pd2dz2(g, spec)                  # SpatialDifferentiator.py:36-40
    _, m = mode_numbers(g)
    return _d2_factor(g.lz) * m ** 2 * spec     # SpatialDifferentiator.py:65-68


def snabla2(g: Grid, spec):
    return sd2dx2(g, spec) + sd2dz2(g, spec)    # Variable.py:111-113


def lap_symbol(g: Grid):
    """Real array -(kx^2) - (kz^2) (SpatialDifferentiator.py:70-74)."""
    n, m = mode_numbers(g)
    return _d2_factor(g.lx) * n ** 2 + _d2_factor(g.lz) * m ** 2


# --------------------------------------------------------------------------
# physical-space stencils  (melvin/SpatialDifferentiator.py:76-185)
# --------------------------------------------------------------------------
def _central1(f, h, order, axis, periodic):
    """First derivative, central, order 2 or 4, along ``axis``."""
    f = np.moveaxis(f, axis, 0)
    out = np.zeros_like(f)
    N = f.shape[0]
    if order == 2:
        out[1:-1] = (f[2:] - f[:-2]) / (2 * h)             # :83, :98
        if periodic:
            out[0] = (f[1] - f[-1]) / (2 * h)               # :85-87
            out[-1] = (f[0] - f[-2]) / (2 * h)
    elif order == 4:
        out[2:-2] = (-0.25 * f[4:] + 2 * f[3:-1] - 2 * f[1:-3]
                     + 0.25 * f[:-4]) / (3 * h)             # :138-140
        if periodic:
            for i in (0, 1, N - 2, N - 1):                  # :141-153
                out[i] = (-0.25 * f[(i + 2) % N] + 2 * f[(i + 1) % N]
                          - 2 * f[(i - 1) % N] + 0.25 * f[(i - 2) % N]) / (3 * h)
    else:
        raise NotImplementedError
    return np.moveaxis(out, 0, axis)


def pddx(g: Grid, f):
    return _central1(f, g.dx, g.fd_order, 0, True)


def pddz(g: Grid, f):
    return _central1(f, g.dz, g.fd_order, 1, not g.fdm_z)


def pd2dz2(g: Grid, f):
    """Second z-derivative, interior only (:106-128); works on complex."""
    out = np.zeros_like(f)
    dz = g.dz
    if g.fd_order == 2:
        out[:, 1:-1] = (f[:, 2:] - 2 * f[:, 1:-1] + f[:, :-2]) / dz ** 2
    elif g.fd_order == 4:
        out[:, 2:-2] = (-1.0 / 12 * f[:, 4:] + 4.0 / 3 * f[:, 3:-1]
                        - 5.0 / 2 * f[:, 2:-2] + 4.0 / 3 * f[:, 1:-3]
                        - 1.0 / 12 * f[:, :-4]) / dz ** 2
    else:
        raise NotImplementedError
    return out


# --------------------------------------------------------------------------
# nonlinear term  (melvin/Variable.py:119-128)
# --------------------------------------------------------------------------
def vec_dot_nabla(g: Grid, q_spec, ux_p, uz_p):
    """Conservative advection  d/dx(ux q) + d/dz(uz q)  -> spectral.

    Returns (spectral result, q in physical space).
    """
    q_p = to_physical(g, q_spec)
    out = pddx(g, ux_p * q_p) + pddz(g, uz_p * q_p)
    return to_spectral(g, out), q_p


# --------------------------------------------------------------------------
# inverse Laplacian  (melvin/LaplacianSolver.py)
# --------------------------------------------------------------------------
def solve_spectral(g: Grid, rhs):
    """rhs / lap with lap[0,0] patched to 1 (LaplacianSolver.py:58-68)."""
    lap = lap_symbol(g).copy()
    lap[0, 0] = 1
    return rhs / lap


def fdm_tridiagonal(g: Grid):
    """(lower, diag, upper) of the nn Dirichlet-z systems, shape (nn, nz).

    LaplacianSolver.py:22-49: off-diagonals 1/dz^2, diagonal
    -((n |2 pi i/lx|)^2 + 2/dz^2), first/last rows replaced by identity.
    """
    nn, nz = g.nn, g.nz
    kx = np.arange(nn) * np.abs(_d1_factor(g.lx))
    lower = np.full((nn, nz), 1.0 / g.dz ** 2)
    upper = np.full((nn, nz), 1.0 / g.dz ** 2)
    diag = np.repeat(-(kx ** 2 + 2.0 / g.dz ** 2)[:, None], nz, axis=1)
    diag[:, 0] = 1.0
    upper[:, 0] = 0.0
    diag[:, -1] = 1.0
    lower[:, -1] = 0.0
    lower[:, 0] = 0.0
    upper[:, -1] = 0.0
    return lower, diag, upper


def solve_fdm(g: Grid, rhs, dtype=np.complex128):
    """Thomas recurrence, vectorised over the nn systems (F8)."""
    lower, diag, upper = fdm_tridiagonal(g)
    nz = g.nz
    cp = np.zeros((g.nn, nz))
    dp = np.zeros((g.nn, nz), dtype=dtype)
    cp[:, 0] = upper[:, 0] / diag[:, 0]
    dp[:, 0] = rhs[:, 0] / diag[:, 0]
    for i in range(1, nz):
        den = diag[:, i] - lower[:, i] * cp[:, i - 1]
        cp[:, i] = upper[:, i] / den
        dp[:, i] = (rhs[:, i] - lower[:, i] * dp[:, i - 1]) / den
    out = np.zeros((g.nn, nz), dtype=dtype)
    out[:, -1] = dp[:, -1]
    for i in range(nz - 2, -1, -1):
        out[:, i] = dp[:, i] - cp[:, i] * out[:, i + 1]
    return out


def solve(g: Grid, rhs):
    return solve_fdm(g, rhs) if g.fdm_z else solve_spectral(g, rhs)


# --------------------------------------------------------------------------
# velocity from vorticity  (melvin/utility.py:62-79)
# --------------------------------------------------------------------------
def velocity_from_vorticity(g: Grid, w_spec):
    """Returns dict(psi_s, ux_s, uz_s, ux_p, uz_p, psi_p)."""
    psi = solve(g, -w_spec)
    res = {"psi_s": psi}
    if g.fdm_z:
        psi_p = to_physical(g, psi)
        res["psi_p"] = psi_p
        res["ux_s"] = None                      # never written (App. A-11)
        res["ux_p"] = -pddz(g, psi_p)
    else:
        res["ux_s"] = -sddz(g, psi)
        res["ux_p"] = to_physical(g, res["ux_s"])
    res["uz_s"] = sddx(g, psi)
    res["uz_p"] = to_physical(g, res["uz_s"])
    return res


# --------------------------------------------------------------------------
# time integration  (melvin/Integrator.py, melvin/TimeDerivative.py)
# --------------------------------------------------------------------------
class History:
    """Ring buffer of RHS levels, zero initialised (TimeDerivative.py:9-45)."""

    def __init__(self, g: Grid):
        self.order = g.int_order
        self.data = np.zeros((self.order,) + g.spectral_shape, dtype=np.complex128)
        self.curr = 0

    def set_current(self, value):
        self.data[self.curr] = value            # TimeDerivative.py:23-24

    def get(self, back=0):
        return self.data[self.curr + back]      # negative wrap, :38-39

    def advance(self):
        self.curr = (self.curr + 1) % self.order


def ab_increment(hist: History, dt):
    """Adams-Bashforth predictor increment (Integrator.py:5-18)."""
    if hist.order == 2:
        return dt / 2 * (3 * hist.get() - hist.get(-1))
    if hist.order == 4:
        return dt / 24 * (55 * hist.get() - 59 * hist.get(-1)
                          + 37 * hist.get(-2) - 9 * hist.get(-3))
    raise NotImplementedError


def integrate_semi_implicit(g: Grid, q, hist: History, lin_op, dt):
    """theta-scheme + AB predictor (Integrator.py:58-63); returns new q."""
    a = g.alpha
    rhs = (1 + (1 - a) * dt * lin_op) * q + ab_increment(hist, dt)
    out = rhs / (1 - a * dt * lin_op)
    hist.advance()
    return out


def integrate_explicit(g: Grid, q, hist: History, diffusion, dt):
    """f0 += diffusion; q += AB(f) (Integrator.py:53-56); returns new q."""
    hist.data[hist.curr] += diffusion
    out = q + ab_increment(hist, dt)
    hist.advance()
    return out


def cfl_dt(g: Grid, dt, ux_p, uz_p):
    """CFL limiter (Integrator.py:35-44): signed max, 0.9 back-off."""
    with np.errstate(divide="ignore"):
        lim = min(g.dx / np.max(ux_p), g.dz / np.max(uz_p))
    if dt > lim or np.isnan(lim):
        raise Exception("CFL condition breached")
    while dt > g.cfl_cutoff * lim:
        dt = dt * 0.9
    return dt


def kinetic_energy(g: Grid, ux_p, uz_p):
    """0.5 * sum(uz^2 + ux^2) / (nx nz)  (melvin/utility.py:42-59)."""
    return 0.5 * np.sum(uz_p ** 2 + ux_p ** 2) / (g.nx * g.nz)


# --------------------------------------------------------------------------
# initial conditions of the example scripts (host side, NumPy in the reference
# as well; restated so tests / bench can build identical inputs on any box)
# --------------------------------------------------------------------------
def _mesh(g: Grid):
    x = np.linspace(0, g.lx, g.nx, endpoint=False)
    z = np.linspace(0, g.lz, g.nz, endpoint=False)
    return np.meshgrid(x, z, indexing="ij")


def sech(x):
    return 1.0 / np.cosh(x)


def ic_taylor_green(g: Grid):
    X, Z = _mesh(g)                              # examples/taylor_green_vortex.py:22-29
    return -2 * np.cos(X) * np.cos(Z)


def ic_kelvin_helmholtz(g: Grid):
    X, Z = _mesh(g)                              # examples/kelvin_helmholtz_instability.py:22-49
    R = np.sqrt((X - (g.lx / 2)) ** 2 + (Z - 0.5) ** 2)
    rng = np.random.default_rng(0)
    w0 = np.power(sech((R - 0.25) / 0.1), 2) / 0.1
    w0 += 0.01 * (2 * rng.random((g.nx, g.nz)) - 1.0)
    return w0


def ic_noise(g: Grid, epsilon=0.01, seed=0):
    rng = np.random.default_rng(seed)            # melvin/utility.py:31-39
    data = np.zeros(g.physical_shape)
    data += epsilon * (2 * rng.random(g.physical_shape) - 1.0)
    return data


def ic_tearing_current(g: Grid):
    X, Z = _mesh(g)                              # examples/resistive_tearing_instability.py:36-60
    rng = np.random.default_rng(0)
    j0 = -np.power(sech((Z - 0.5) / 0.01), 2) / 0.01
    j0 += 0.01 * (2 * rng.random((g.nx, g.nz)) - 1.0)
    return j0


def ic_rbc_temperature(g: Grid):
    X, Z = _mesh(g)                              # examples/rayleigh_benard_convection.py:20-31
    return 1 - Z + 0.01 * (np.sin(np.pi * X / 2.44))


# --------------------------------------------------------------------------
# whole-loop drivers: the per-iteration operator sequences of the example
# scripts, with the Simulation.end_loop ticker semantics
# (melvin/Simulation.py:205-209, melvin/Ticker.py:14-25).
# --------------------------------------------------------------------------
class _Ticker:
    def __init__(self, cadence):
        self.cadence = cadence
        self.counter = 0

    def due(self, loop_counter):
        if self.counter < loop_counter:
            self.counter += self.cadence
            return True
        return False


class Run:
    """Book-keeping shared by the loop drivers."""

    def __init__(self, g: Grid, dt, cfl_cadence=10, tracker_cadence=100):
        self.g = g
        self.dt = dt
        self.t = 0.0
        self.loop = 0
        self._cfl = _Ticker(cfl_cadence)
        self._trk = _Ticker(tracker_cadence)
        self.times = []
        self.ke = []
        self.extra = []

    def end_loop(self, ux_p, uz_p, extra_fn=None):
        self.loop += 1
        self.t += self.dt
        if self._cfl.due(self.loop):
            self.dt = cfl_dt(self.g, self.dt, ux_p, uz_p)
        if self._trk.due(self.loop):
            self.times.append(self.t)
            self.ke.append(kinetic_energy(self.g, ux_p, uz_p))
            if extra_fn is not None:
                self.extra.append(extra_fn())


def step_single_scalar(g: Grid, run: Run, w, dw: History, coef):
    """TG / KH iteration (examples/kelvin_helmholtz_instability.py:115-131)."""
    vel = velocity_from_vorticity(g, w)
    lin_op = coef * lap_symbol(g)
    nl, _ = vec_dot_nabla(g, w, vel["ux_p"], vel["uz_p"])
    dw.set_current(-nl)
    w = integrate_semi_implicit(g, w, dw, lin_op, run.dt)
    run.end_loop(vel["ux_p"], vel["uz_p"])
    return w


def step_double_diffusive(g: Grid, run: Run, state, hists, Pr, R0, tau):
    """DDC iteration (examples/double_diffusive_convection.py:100-126)."""
    w, tmp, xi = state
    dw, dtmp, dxi = hists
    vel = velocity_from_vorticity(g, w)
    ux_p, uz_p = vel["ux_p"], vel["uz_p"]
    lap = lap_symbol(g)

    nl, _ = vec_dot_nabla(g, w, ux_p, uz_p)
    dw.set_current(-nl + Pr * sddx(g, xi) - Pr * sddx(g, tmp))
    w = integrate_semi_implicit(g, w, dw, Pr * lap, run.dt)

    nl, tmp_p = vec_dot_nabla(g, tmp, ux_p, uz_p)
    dtmp.set_current(-nl - vel["uz_s"])
    tmp = integrate_semi_implicit(g, tmp, dtmp, lap, run.dt)

    nl, _ = vec_dot_nabla(g, xi, ux_p, uz_p)
    dxi.set_current(-nl - vel["uz_s"] / R0)
    xi = integrate_semi_implicit(g, xi, dxi, tau * lap, run.dt)

    tmp[:, 0] = 0.0
    xi[:, 0] = 0.0
    # Nusselt number uses the physical temperature of *this* iteration's
    # vec_dot_nabla (examples/double_diffusive_convection.py:20-23).
    run.end_loop(ux_p, uz_p, lambda: 1.0 - np.mean(tmp_p * uz_p))
    return (w, tmp, xi)


def step_tearing(g: Grid, run: Run, state, hists, Re, S):
    """MHD iteration (examples/resistive_tearing_instability.py:125-148)."""
    w, j = state
    dw, dj = hists
    vel = velocity_from_vorticity(g, w)
    mag = velocity_from_vorticity(g, j)
    ux_p, uz_p = vel["ux_p"], vel["uz_p"]
    bx_p, bz_p = mag["ux_p"], mag["uz_p"]
    lap = lap_symbol(g)

    a, _ = vec_dot_nabla(g, w, ux_p, uz_p)
    b, _ = vec_dot_nabla(g, j, bx_p, bz_p)
    dw.set_current(-a + b)
    w = integrate_semi_implicit(g, w, dw, 1.0 / Re * lap, run.dt)

    a, _ = vec_dot_nabla(g, j, ux_p, uz_p)
    b, _ = vec_dot_nabla(g, w, bx_p, bz_p)      # uses the UPDATED w
    dj.set_current(-a + b)
    j = integrate_semi_implicit(g, j, dj, 1.0 / S * lap, run.dt)

    run.end_loop(ux_p, uz_p)
    return (w, j)


def step_rayleigh_benard(g: Grid, run: Run, state, hists, Pr, Ra):
    """RBC iteration, Fourier-x / FD-z
    (examples/rayleigh_benard_convection.py:95-145)."""
    w, tmp, psi = state
    dw, dtmp = hists
    vel = velocity_from_vorticity(g, w)
    psi = vel["psi_s"]
    ux_p, uz_p = vel["ux_p"], vel["uz_p"]

    diffusion = Pr * snabla2(g, w)
    nl, _ = vec_dot_nabla(g, w, ux_p, uz_p)
    dw.set_current(-nl - Pr * Ra * sddx(g, tmp))
    w = integrate_explicit(g, w, dw, diffusion, run.dt)

    diffusion = snabla2(g, tmp)
    nl, _ = vec_dot_nabla(g, tmp, ux_p, uz_p)
    dtmp.set_current(-nl)
    tmp = integrate_explicit(g, tmp, dtmp, diffusion, run.dt)

    k = 1 if g.fd_order == 2 else 2
    w[1:, :k] = 0.0
    w[1:, -k:] = 0.0
    psi[1:, :k] = 0.0
    psi[1:, -k:] = 0.0
    tmp[0, :k] = 1.0
    tmp[0, -k:] = 0.0
    tmp[1:, :k] = 0.0
    tmp[1:, -k:] = 0.0
    psi[0, :] = 0.0
    w[0, :] = 0.0

    run.end_loop(ux_p, uz_p)
    return (w, tmp, psi)


# --------------------------------------------------------------------------
# convenience: complete runs used by tests and by the CPU-baseline leg
# --------------------------------------------------------------------------
def run_single_scalar(g: Grid, w0_phys, coef, dt, nsteps, tracker_cadence=100,
                      cfl_cadence=10, snapshots=()):
    run = Run(g, dt, cfl_cadence, tracker_cadence)
    w = to_spectral(g, w0_phys)
    dw = History(g)
    snaps = {}
    for _ in range(nsteps):
        w = step_single_scalar(g, run, w, dw, coef)
        if run.loop in snapshots:
            snaps[run.loop] = w.copy()
    return w, run, snaps


def relative_l2(a, b):
    den = np.linalg.norm(np.asarray(b).ravel())
    num = np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel())
    return num / den if den > 0 else num


__all__ = [name for name in globals() if not name.startswith("_")]
