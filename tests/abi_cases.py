"""Kernel-level cases driven straight through the C ABI (include/melvin_b200.h).

Each case takes a harness ``H`` providing ``H.Ctx(nx, nz, lx, lz, fdm_z=, fd_order=)``
(with ``.call(name, *args)``, ``.ibuf()``, ``.close()``) and ``H.ptr(numpy_array)``:
  * tests/gpu_harness.py  -- the real library on a CUDA device (``-m gpu`` tests);
  * tests/emu/emu_harness.py -- the host emulation build (development, CPU box).
Expected values always come from the oracle.
"""
import ctypes

import numpy as np

from melvin import _capi
from oracle import melvin_oracle as mo

SIZES_2D = [(16, 16), (32, 64), (64, 32), (128, 16), (256, 128), (512, 32), (16, 1024),
            (2048, 16), (16, 2048), (4096, 16), (16, 4096), (8192, 16), (16, 8192)]
SIZES_1D = [(16, 16), (64, 13), (64, 32), (256, 24), (1024, 6)]
SIZES_FUSED = [(32, 64), (64, 32), (256, 64), (32, 512), (16, 1024), (18 * 0 + 16, 2048)]


# long-line kernels (16384-point lines: x passes split in two half-length transforms, z stage
# on single real rows) forced onto small grids through MLV_FORCE_SPLIT (bit 0: x, bit 1: z)
SIZES_SPLIT = [(128, 128, 3), (256, 128, 1), (128, 256, 2), (256, 256, 3), (128, 32, 1), (32, 128, 2)]


class forced_split:
    """Context manager: contexts created inside take the long-line code paths."""

    def __init__(self, bits):
        self.bits = bits

    def __enter__(self):
        import os
        self.old = os.environ.get("MLV_FORCE_SPLIT")
        os.environ["MLV_FORCE_SPLIT"] = str(self.bits)

    def __exit__(self, *exc):
        import os
        if self.old is None:
            os.environ.pop("MLV_FORCE_SPLIT", None)
        else:
            os.environ["MLV_FORCE_SPLIT"] = self.old


def case_split_lines(H, nx, nz, bits, order):
    with forced_split(bits):
        ctx = H.Ctx(nx, nz, 1.5, 1.0)
        assert ctx.lib.mlv_long_lines(ctx.h) == bits      # the long-line kernels really run
        ctx.close()
        case_transforms_2d(H, nx, nz)
        case_fused_advection_step(H, nx, nz, order)


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


SIZES_2D = [(16, 16), (32, 64), (64, 32), (128, 16), (256, 128), (512, 32), (16, 1024),
            (2048, 16), (16, 2048), (4096, 16), (16, 4096), (8192, 16), (16, 8192)]


def case_transforms_2d(H, nx, nz, tol=1e-14):
    g = mo.Grid(nx, nz, 1.5, 1.0)
    ctx = H.Ctx(nx, nz, 1.5, 1.0)
    rng = np.random.default_rng(nx * 1000 + nz)
    phys = rng.standard_normal((nx, nz))
    spec = np.zeros(g.spectral_shape, complex)
    I = ctx.ibuf()
    ctx.call("mlv_to_spectral", H.ptr(phys), H.ptr(I), H.ptr(spec))
    assert rel(spec, mo.to_spectral(g, phys)) < tol
    srand = rng.standard_normal(g.spectral_shape) + 1j * rng.standard_normal(g.spectral_shape)
    out = np.zeros((nx, nz))
    ctx.call("mlv_to_physical", H.ptr(srand), H.ptr(I), H.ptr(out))
    assert rel(out, mo.to_physical(g, srand)) < tol
    ctx.close()


def case_transforms_1d_fdm(H, nx, nz, tol=1e-14):
    g = mo.Grid(nx, nz, 2.44, 1.0, fdm_z=True)
    ctx = H.Ctx(nx, nz, 2.44, 1.0, fdm_z=True)
    rng = np.random.default_rng(7)
    phys = rng.standard_normal((nx, nz))
    spec = np.zeros(g.spectral_shape, complex)
    ctx.call("mlv_to_spectral", H.ptr(phys), None, H.ptr(spec))
    assert rel(spec, mo.to_spectral(g, phys)) < tol
    srand = rng.standard_normal(g.spectral_shape) + 1j * rng.standard_normal(g.spectral_shape)
    out = np.zeros((nx, nz))
    ctx.call("mlv_to_physical", H.ptr(srand), None, H.ptr(out))
    assert rel(out, mo.to_physical(g, srand)) < tol
    ctx.close()


def case_fused_advection_step(H, nx, nz, order, tol=1e-13):
    """x-inverse (velocity prologue) -> fused z stage -> x-forward with stencil
    symbols == oracle vec_dot_nabla on oracle velocities."""
    g = mo.Grid(nx, nz, 1.5, 1.0, fd_order=order)
    ctx = H.Ctx(nx, nz, 1.5, 1.0, fd_order=order)
    rng = np.random.default_rng(3)
    w = mo.to_spectral(g, rng.standard_normal((nx, nz)))
    q = mo.to_spectral(g, rng.standard_normal((nx, nz)))
    vel = mo.velocity_from_vorticity(g, w)
    want, _ = mo.vec_dot_nabla(g, q, vel["ux_p"], vel["uz_p"])

    Iux, Iuz, Iq, IA, IB = (ctx.ibuf() for _ in range(5))
    srcs = (ctypes.c_void_p * 3)(H.ptr(w), H.ptr(w), H.ptr(q))
    ops = (ctypes.c_int32 * 3)(_capi.OP_UX, _capi.OP_UZ, _capi.OP_IDENT)
    dsts = (ctypes.c_void_p * 3)(H.ptr(Iux), H.ptr(Iuz), H.ptr(Iq))
    ctx.call("mlv_x_inverse", 3, srcs, ops, dsts)
    red = np.zeros(4)
    ctx.call("mlv_advect_z", H.ptr(Iux), H.ptr(Iuz), H.ptr(Iq), H.ptr(IA), H.ptr(IB), H.ptr(red))
    np.testing.assert_allclose(red[0], vel["ux_p"].max(), rtol=1e-13)
    np.testing.assert_allclose(red[1], vel["uz_p"].max(), rtol=1e-13)
    np.testing.assert_allclose(red[2], (vel["ux_p"] ** 2).sum(), rtol=1e-13)
    np.testing.assert_allclose(red[3], (vel["uz_p"] ** 2).sum(), rtol=1e-13)

    got = np.zeros(g.spectral_shape, complex)
    d = _capi.XFwd()
    d.nf, d.mode = 2, 0
    d.src[0], d.src[1] = H.ptr(IA), H.ptr(IB)
    d.sym[0], d.sym[1] = _capi.SYM_FDX, _capi.SYM_FDZ
    d.coef[0] = d.coef[1] = 1.0
    d.dst = H.ptr(got)
    ctx.call("mlv_x_forward", ctypes.byref(d))
    assert rel(got, want) < tol

    # fused epilogue: f0 = -N + 0.3*ddx(q); theta-scheme AB2 update of w
    hist = mo.History(g)
    hist.data[1] = rng.standard_normal(g.spectral_shape) + 0j
    hist.curr = 0
    f0_want = -want + 0.3 * mo.sddx(g, q)
    hist.set_current(f0_want)
    w_want = mo.integrate_semi_implicit(g, w.copy(), hist, 0.02 * mo.lap_symbol(g), 1e-3)
    f0 = np.zeros(g.spectral_shape, complex)
    fm1 = hist.data[1].copy()
    w_new = np.zeros(g.spectral_shape, complex)
    d.mode = 1
    d.coef[0] = d.coef[1] = -1.0
    d.lin = _capi.make_lin_terms([(0.3, _capi.OP_DDX, H.ptr(q))])
    d.integ.ab_order, d.integ.scheme = 2, _capi.SCHEME_SI_LAP
    d.integ.dt, d.integ.alpha, d.integ.lcoef = 1e-3, g.alpha, 0.02
    d.integ.q_in, d.integ.q_out = H.ptr(w), H.ptr(w_new)
    d.integ.f0, d.integ.fm1 = H.ptr(f0), H.ptr(fm1)
    ctx.call("mlv_x_forward", ctypes.byref(d))
    assert rel(f0, f0_want) < tol
    assert rel(w_new, w_want) < tol
    # a NaN anywhere in a velocity makes its CFL maximum NaN, as numpy.max does
    # (Integrator.py:41 tests np.isnan(cfl_dt)); the other component is untouched
    Iux[nx // 2 + 1, 1] = np.nan
    ctx.call("mlv_advect_z", H.ptr(Iux), H.ptr(Iuz), H.ptr(Iq), H.ptr(IA), H.ptr(IB), H.ptr(red))
    assert np.isnan(red[0]) and np.isnan(red[2])
    np.testing.assert_allclose(red[1], vel["uz_p"].max(), rtol=1e-13)
    ctx.close()


def case_pointwise_and_stencils(H):
    nx, nz = 32, 64
    g = mo.Grid(nx, nz, 1.5, 1.0)
    ctx = H.Ctx(nx, nz, 1.5, 1.0)
    rng = np.random.default_rng(5)
    s = rng.standard_normal(g.spectral_shape) + 1j * rng.standard_normal(g.spectral_shape)
    out = np.zeros_like(s)
    for op, fn in ((_capi.OP_DDX, mo.sddx), (_capi.OP_DDZ, mo.sddz), (_capi.OP_D2DX2, mo.sd2dx2),
                   (_capi.OP_D2DZ2, mo.sd2dz2), (_capi.OP_LAP, mo.snabla2),
                   (_capi.OP_INVLAP, mo.solve_spectral)):
        lt = _capi.make_lin_terms([(1.0, op, H.ptr(s))])
        ctx.call("mlv_spec_lincomb", ctypes.byref(lt), H.ptr(out))
        assert rel(out, fn(g, s)) < 1e-15, op
    vel = mo.velocity_from_vorticity(g, s)
    for op, key in ((_capi.OP_PSI, "psi_s"), (_capi.OP_UX, "ux_s"), (_capi.OP_UZ, "uz_s")):
        lt = _capi.make_lin_terms([(1.0, op, H.ptr(s))])
        ctx.call("mlv_spec_lincomb", ctypes.byref(lt), H.ptr(out))
        assert rel(out, vel[key]) < 1e-15, key
    lap = np.zeros(g.spectral_shape)
    ctx.call("mlv_lap_array", 0.5, H.ptr(lap))
    assert rel(lap, 0.5 * mo.lap_symbol(g)) < 1e-15
    f = rng.standard_normal((nx, nz))
    o = np.zeros_like(f)
    for order in (2, 4):
        go = mo.Grid(nx, nz, 1.5, 1.0, fd_order=order)
        ctx.call("mlv_stencil", H.ptr(f), H.ptr(o), nx, nz, 1, 0, order, 1, 0, go.dx)
        assert rel(o, mo.pddx(go, f)) < 1e-14
        ctx.call("mlv_stencil", H.ptr(f), H.ptr(o), nx, nz, 1, 1, order, 1, 0, go.dz)
        assert rel(o, mo.pddz(go, f)) < 1e-14
    ctx.close()


def case_fdm_solver_and_stencils(H):
    nx, nz = 64, 40
    for order in (2, 4):
        g = mo.Grid(nx, nz, 2.44, 1.0, fdm_z=True, fd_order=order)
        ctx = H.Ctx(nx, nz, 2.44, 1.0, fdm_z=True, fd_order=order)
        rng = np.random.default_rng(11)
        s = rng.standard_normal(g.spectral_shape) + 1j * rng.standard_normal(g.spectral_shape)
        out = np.zeros_like(s)
        ctx.call("mlv_solve_fdm", H.ptr(s), H.ptr(out))
        assert rel(out, mo.solve_fdm(g, s)) < 1e-13
        ctx.call("mlv_stencil", H.ptr(s), H.ptr(out), g.nn, nz, 2, 1, order, 0, 1, g.dz)
        assert rel(out, mo.pd2dz2(g, s)) < 1e-14
        f = rng.standard_normal((nx, nz))
        o = np.zeros_like(f)
        ctx.call("mlv_stencil", H.ptr(f), H.ptr(o), nx, nz, 1, 1, order, 0, 0, g.dz)
        assert rel(o, mo.pddz(g, f)) < 1e-14
        ux, uz = rng.standard_normal((nx, nz)), rng.standard_normal((nx, nz))
        ctx.call("mlv_advect_phys", H.ptr(ux), H.ptr(uz), H.ptr(f), H.ptr(o))
        assert rel(o, mo.pddx(g, ux * f) + mo.pddz(g, uz * f)) < 1e-14
        ctx.close()


def case_fdm_fused_step(H, nx, nz, order, truth_rows=4):
    """K1 (batched solve + velocities), K2 (fused 1-D advection) and K3 (row-wise right-hand side +
    AB + explicit update) against the oracle's eager statements of the same operations
    (utility.py:62-79, Variable.py:119-128, Integrator.py:53-56)."""
    g = mo.Grid(nx, nz, 2.44, 1.0, fdm_z=True, fd_order=order, integrator="explicit", int_order=4)
    ctx = H.Ctx(nx, nz, 2.44, 1.0, fdm_z=True, fd_order=order)
    rng = np.random.default_rng(nx + nz + order)
    cplx = lambda: rng.standard_normal(g.spectral_shape) + 1j * rng.standard_normal(g.spectral_shape)  # noqa: E731
    w = mo.to_spectral(g, rng.standard_normal((nx, nz)))
    q = mo.to_spectral(g, rng.standard_normal((nx, nz)))
    # ---- K1: psi = solve(-w), uxh = -pddz(psi), uzh = i kx psi
    psi, uxh, uzh = (np.zeros(g.spectral_shape, complex) for _ in range(3))
    ctx.call("mlv_fdm_velocity", H.ptr(w), H.ptr(psi), H.ptr(uxh), H.ptr(uzh))
    vel = mo.velocity_from_vorticity(g, w)
    # two fp64 Thomas variants (c' = a/den there, a * (1/den) here) differ by an ulp per pivot, which
    # the conditioning of the n = 0 system (~nz^2) amplifies; identical when 1/dz^2 is a power of two
    assert rel(psi, vel["psi_s"]) < 1e-11
    assert rel(uzh, vel["uz_s"]) < 1e-11
    # the z stencil commutes with the x transform: c2r(uxh) == -pddz(c2r(psi))
    assert rel(mo.to_physical(g, uxh), vel["ux_p"]) < 1e-11
    # F8 gate: at least as close to the extended-precision solution as the oracle's fp64 Thomas
    rows = sorted(set(np.linspace(0, g.nn - 1, truth_rows).astype(int)))
    truth = np.stack([thomas_longdouble_row(g, -w[n], n) for n in rows])
    err_dev = max(float(np.linalg.norm(psi[n] - t) / np.linalg.norm(t)) for n, t in zip(rows, truth))
    err_ora = max(float(np.linalg.norm(vel["psi_s"][n] - t) / np.linalg.norm(t)) for n, t in zip(rows, truth))
    # ... and than the reference's own SuperLU solve of the same systems (scipy.sparse.linalg.factorized,
    # LaplacianSolver.py:53-55), which is the accuracy the 1e-12 parity gate cannot exceed on this path
    err_slu = max(float(np.linalg.norm(superlu_row(g, -w[n], n) - t) / np.linalg.norm(t)) for n, t in zip(rows, truth))
    assert err_dev < 1e-9 and err_dev < 2 * err_ora + 1e-15 and err_dev < 2 * err_slu + 1e-15, (err_dev, err_ora, err_slu)
    out = np.zeros_like(w)
    ctx.call("mlv_solve_fdm", H.ptr(q), H.ptr(out))
    assert rel(out, mo.solve_fdm(g, q)) < 1e-11
    # ---- K2: x spectra of ux q and uz q, reductions
    A, B = np.zeros_like(w), np.zeros_like(w)
    red = np.zeros(4)
    ctx.call("mlv_fdm_advect", H.ptr(uxh), H.ptr(uzh), H.ptr(q), H.ptr(A), H.ptr(B), H.ptr(red))
    ux_p, uz_p, q_p = mo.to_physical(g, uxh), mo.to_physical(g, uzh), mo.to_physical(g, q)
    assert rel(A, mo.to_spectral(g, ux_p * q_p)) < 1e-12
    assert rel(B, mo.to_spectral(g, uz_p * q_p)) < 1e-12
    np.testing.assert_allclose(red, [ux_p.max(), uz_p.max(), (ux_p ** 2).sum(), (uz_p ** 2).sum()], rtol=1e-12)
    # ---- K3: f0 = -(d/dx(ux q) + d/dz(uz q)) - 3 ddx(w) + 0.5 snabla2(q); AB4 explicit update of q
    nl, _ = mo.vec_dot_nabla(g, q, ux_p, uz_p)
    f0_want = -nl - 3.0 * mo.sddx(g, w) + 0.5 * mo.snabla2(g, q)
    hist = mo.History(g)
    for k in range(1, 4):
        hist.data[k] = cplx()
    hist.curr = 0
    hist.set_current(f0_want)
    q_want = mo.integrate_explicit(g, q.copy(), hist, 0.0 * q, 1e-3)
    f0, q_new = np.zeros_like(w), np.zeros_like(w)
    fm = [hist.data[k].copy() for k in (3, 2, 1)]            # curr-1, curr-2, curr-3 with negative wrap
    lt = _capi.make_lin_terms([(-1.0, _capi.OP_FDX_SYM, H.ptr(A)), (-1.0, _capi.OP_FDM_DDZ, H.ptr(B)),
                               (-3.0, _capi.OP_DDX, H.ptr(w)), (0.5, _capi.OP_FDM_NABLA2, H.ptr(q))])
    gi = _capi.Integ()
    gi.ab_order, gi.scheme, gi.dt, gi.f0_set = 4, _capi.SCHEME_EXPLICIT, 1e-3, 1
    gi.q_in, gi.q_out, gi.f0 = H.ptr(q), H.ptr(q_new), H.ptr(f0)
    gi.fm1, gi.fm2, gi.fm3 = H.ptr(fm[0]), H.ptr(fm[1]), H.ptr(fm[2])
    ctx.call("mlv_integrate", ctypes.byref(lt), ctypes.byref(gi))
    assert rel(f0, f0_want) < 1e-11
    assert rel(q_new, q_want) < 1e-12
    # the stencil operators alone, through the expression kernel
    for op, want in ((_capi.OP_FDM_D2DZ2, mo.pd2dz2(g, q)), (_capi.OP_FDM_DDZ, mo.pddz(g, q)),
                     (_capi.OP_FDM_NABLA2, mo.snabla2(g, q))):
        lt1 = _capi.make_lin_terms([(1.0, op, H.ptr(q))])
        ctx.call("mlv_spec_lincomb", ctypes.byref(lt1), H.ptr(out))
        assert rel(out, want) < 1e-13, op
    ctx.close()


def superlu_row(g, r, n):
    """The reference's solver for system n: the matrix of LaplacianSolver.py:25-49 factorised by SuperLU."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    nz = r.shape[0]
    kx = n * 2 * np.pi / g.lx
    diag = np.full(nz, -(kx ** 2 + 2.0 / g.dz ** 2))
    off = np.full(nz, 1.0 / g.dz ** 2)
    m = sp.dia_matrix((np.array([off, diag, off]), np.array([-1, 0, 1])), shape=(nz, nz), dtype=np.complex128).tolil()
    m[0, 0], m[0, 1], m[-1, -1], m[-1, -2] = 1.0, 0.0, 1.0, 0.0
    return spl.factorized(m.tocsc())(r.astype(np.complex128))


def thomas_longdouble_row(g, r, n):
    """One system of LaplacianSolver.py:22-56 in extended precision."""
    ld, cld = np.longdouble, np.clongdouble
    nz = r.shape[0]
    off = ld(1) / (ld(g.dz) * ld(g.dz))
    kx = ld(n) * ld(2) * ld(np.pi) / ld(g.lx)
    b = -(kx * kx + ld(2) * off)
    rr = r.astype(cld)
    cp = np.zeros(nz, dtype=ld)
    dp = np.zeros(nz, dtype=cld)
    dp[0] = rr[0]
    for i in range(1, nz - 1):
        den = b - off * cp[i - 1]
        cp[i] = off / den
        dp[i] = (rr[i] - off * dp[i - 1]) / den
    dp[nz - 1] = rr[nz - 1]
    x = np.zeros(nz, dtype=cld)
    x[nz - 1] = dp[nz - 1]
    for i in range(nz - 2, -1, -1):
        x[i] = dp[i] - cp[i] * x[i + 1]
    return x


def case_integrate_and_array_ops(H):
    nx, nz = 32, 32
    ctx = H.Ctx(nx, nz, 1.0, 1.0)
    rng = np.random.default_rng(2)
    for order in (2, 4):
        g = mo.Grid(nx, nz, 1.0, 1.0, int_order=order)
        shape = g.spectral_shape
        q = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        hist = mo.History(g)
        hist.data[:] = rng.standard_normal(hist.data.shape) + 1j * rng.standard_normal(hist.data.shape)
        hist.curr = 1
        diff = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        levels = [hist.data[(hist.curr - k) % order].copy() for k in range(order)]
        h2 = mo.History(g)
        h2.data[:] = hist.data
        h2.curr = hist.curr
        want = mo.integrate_explicit(g, q.copy(), h2, diff, 1e-2)
        out = np.zeros_like(q)
        ig = _capi.Integ()
        ig.ab_order, ig.scheme, ig.dt, ig.alpha = order, _capi.SCHEME_EXPLICIT, 1e-2, 0.51
        ig.q_in, ig.q_out, ig.f0 = H.ptr(q), H.ptr(out), H.ptr(levels[0])
        ig.fm1 = H.ptr(levels[1])
        if order == 4:
            ig.fm2, ig.fm3 = H.ptr(levels[2]), H.ptr(levels[3])
        lt = _capi.make_lin_terms([(1.0, _capi.OP_IDENT, H.ptr(diff))])
        ctx.call("mlv_integrate", ctypes.byref(lt), ctypes.byref(ig))
        assert rel(out, want) < 1e-15
        assert rel(levels[0], h2.data[hist.curr]) < 1e-15
        L = rng.standard_normal(shape)
        h3 = mo.History(g)
        h3.data[:] = hist.data
        h3.curr = hist.curr
        want = mo.integrate_semi_implicit(g, q.copy(), h3, L, 1e-2)
        levels[0] = hist.data[hist.curr].copy()
        ig.scheme, ig.larr, ig.f0 = _capi.SCHEME_SI_ARR, H.ptr(L), H.ptr(levels[0])
        ctx.call("mlv_integrate", None, ctypes.byref(ig))
        assert rel(out, want) < 1e-15
    # elementwise on strided views + reductions
    a = rng.standard_normal((20, 30))
    b = rng.standard_normal((20, 30))
    o = np.zeros((10, 15))
    e = _capi.Ew()
    e.op, e.rows, e.cols = _capi.EW_MUL, 10, 15
    e.out_kind, e.a_kind, e.b_kind = _capi.KIND_REAL, _capi.KIND_REAL, _capi.KIND_REAL
    e.out = _capi.View(H.ptr(o), 15, 1)
    e.a = _capi.View(H.ptr(a), 60, 2)
    e.b = _capi.View(H.ptr(b[10:]), 30, 1)
    ctx.call("mlv_elementwise", ctypes.byref(e))
    np.testing.assert_allclose(o, a[::2, ::2] * b[10:, :15], rtol=1e-15)
    r = np.zeros(1)
    va = _capi.View(H.ptr(a), 30, 1)
    vb = _capi.View(H.ptr(b), 30, 1)
    for op, want in ((_capi.RED_SUM, a.sum()), (_capi.RED_MAX, a.max()), (_capi.RED_MIN, a.min()),
                     (_capi.RED_SUMSQ, (a * a).sum()), (_capi.RED_SUMPROD, (a * b).sum())):
        ctx.call("mlv_reduce", op, 20, 30, ctypes.byref(va), ctypes.byref(vb), H.ptr(r))
        np.testing.assert_allclose(r[0], want, rtol=1e-13)
    ctx.close()
