"""Host-logic cases shared by the emulation (CPU) and the GPU test modules: restart from the
reference's dump format (same and changed resolution) and the Adams-Moulton
corrector formulas (reference melvin/Integrator.py:20-33; dead code there)."""
from functools import partial

import numpy as np

import parity_cases as pc
from melvin import b200 as xp
from melvin.utility import calc_kinetic_energy, calc_velocity_from_vorticity
from oracle import melvin_oracle as mo


def _build(n, lx, lz, order=2):
    d = pc.base_params(n, n, lx, lz, initial_dt=1e-3, nu=0.25, integrator_order=order)
    p, sim, (w,), (dw,), psi, ux, uz = pc.make_sim(d, ["w"], ["dw"], [pc.CE, pc.CE])
    sim.config_dump([w], [dw])
    sim.config_scalar_trackers({"ke": partial(calc_kinetic_energy, ux, uz, xp, p)})
    return p, sim, w, dw, psi, ux, uz


def _step(p, sim, w, dw, psi, ux, uz):
    calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
    dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
    sim._integrator.integrate(w, dw, p.nu * w.lap())
    sim.end_loop()


def restart_roundtrip():
    """dump after 5 steps, run 5 more; a fresh simulation loaded from the dump and stepped 5
    times must hold bit-identical state."""
    g = mo.Grid(64, 64, 2 * np.pi, 2 * np.pi)
    with pc.scratch_cwd():
        a = _build(64, g.lx, g.lz)
        a[2].load(mo.ic_taylor_green(g) + 0.1 * mo.ic_noise(g, 1.0, 3), is_physical=True)
        for _ in range(5):
            _step(*a)
        a[1].dump(a[1]._dump_ticker)
        idx = a[1]._dump_ticker.times_fired
        for _ in range(5):
            _step(*a)
        b = _build(64, g.lx, g.lz)
        b[1].load(idx)
        for _ in range(5):
            _step(*b)
        return (b[1]._loop_counter, a[1]._loop_counter), pc.host(b[2][:]), pc.host(a[2][:])


def restart_resolution_change(n_from=64, n_to=128):
    """A dump written at n_from^2 is loaded into a simulation at n_to^2: state and history are
    re-sampled in spectral space (modes common to both grids are kept, signed-mode aware)."""
    g = mo.Grid(n_from, n_from, 2 * np.pi, 2 * np.pi)
    rng = np.random.default_rng(5)
    with pc.scratch_cwd():
        a = _build(n_from, g.lx, g.lz)
        a[2].load(mo.ic_taylor_green(g) + rng.standard_normal(g.physical_shape), is_physical=True)
        for _ in range(3):
            _step(*a)
        a[1].dump(a[1]._dump_ticker)
        idx = a[1]._dump_ticker.times_fired
        w_old, h_old = pc.host(a[2][:]), pc.host(a[3].get_all())
        b = _build(n_to, g.lx, g.lz)
        b[1].load(idx)
        w_new, h_new = pc.host(b[2][:]), pc.host(b[3].get_all())
        meta = (b[1]._loop_counter, float(b[1]._t), b[3].get_curr_idx(), a[3].get_curr_idx())
        _step(*b)                                   # and it keeps running
        ok = bool(np.all(np.isfinite(pc.host(b[2][:]))))
    return w_old, h_old, w_new, h_new, meta, ok


def expected_rescale(old, nn_new, nm_new):
    """Signed-mode re-sampling of a (2nn+1, nm) spectrum (NumPy statement of the intent of
    reference melvin/Variable.py:142-151)."""
    nn_old, nm_old = (old.shape[0] - 1) // 2, old.shape[1]
    out = np.zeros((2 * nn_new + 1, nm_new), dtype=np.complex128)
    k, m = min(nn_old, nn_new), min(nm_old, nm_new)
    for n in range(-k, k + 1):
        out[n if n >= 0 else n + 2 * nn_new + 1, :m] = old[n if n >= 0 else n + 2 * nn_old + 1, :m]
    return out


def corrector_formulas(order):
    """Integrator.corrector(dvar) for AB/AM order 2 and 4 against the formulas of
    Integrator.py:20-33 evaluated with NumPy on the same history."""
    rng = np.random.default_rng(order)
    with pc.scratch_cwd():
        p, sim, w, dw, psi, ux, uz = _build(32, 2 * np.pi, 2 * np.pi, order=order)
        levels = []
        for k in range(order):
            lv = rng.standard_normal(p.spectral_shape) + 1j * rng.standard_normal(p.spectral_shape)
            dw[:] = lv
            levels.append(lv)
            if k < order - 1:
                dw.advance()
        dt = sim._integrator._dt
        got_c = pc.host(sim._integrator.corrector(dw))
        got_p = pc.host(sim._integrator.predictor(dw))
    f = levels[::-1]                                  # f[0] newest
    if order == 2:
        want_c = dt / 2 * (f[0] + f[1])
        want_p = dt / 2 * (3 * f[0] - f[1])
    else:
        want_c = dt / 24 * (9 * f[0] + 19 * f[1] - 5 * f[2] + f[3])
        want_p = dt / 24 * (55 * f[0] - 59 * f[1] + 37 * f[2] - 9 * f[3])
    return got_c, want_c, got_p, want_p


def predictor_corrector_convergence(order):
    """PECE (AB/AM pair) on dq/dt = nu lap(q) - c ddx(q) for a single Fourier mode,
    whose exact solution is known: the error after a fixed time must fall with the scheme's
    order when dt is halved.  Returns the list of (dt, error)."""
    nx = nz = 32
    nu, c, T = 0.05, 0.7, 0.2
    g = mo.Grid(nx, nz, 2 * np.pi, 2 * np.pi)
    x = np.linspace(0, 2 * np.pi, nx, endpoint=False)
    X = np.meshgrid(x, x, indexing="ij")
    q0 = np.cos(2 * X[0] + X[1])
    lam = -nu * (4 + 1) - 1j * c * 2                    # symbol of nu lap - c ddx on mode (2, 1)
    q0_s = mo.to_spectral(g, q0)
    res = []
    for nsteps in (20, 40, 80):
        dt = T / nsteps
        with pc.scratch_cwd():
            d = pc.base_params(nx, nz, g.lx, g.lz, initial_dt=dt, integrator_order=order, integrator="explicit")
            p, sim, (q,), (dq,), psi, ux, uz = pc.make_sim(d, ["q"], ["dq"], [pc.CE, pc.CE])
            q.load(q0, is_physical=True)
            rhs = lambda: nu * q.snabla2() - c * q.sddx()            # noqa: E731
            # start the multistep history from the exact solution (the reference zero-initialises it,
            # F5, which would cap the observed order at 1)
            for k in range(order - 1, 0, -1):
                dq.set(pc.host(q[:]) * 0 + lam * q0_s * np.exp(-lam * k * dt), idx=0)
                dq.advance()
            dq[:] = rhs()
            for _ in range(nsteps):
                sim._integrator.predictor_corrector(q, dq, rhs)
            got = pc.host(q[:])
        want = q0_s * np.exp(lam * T)
        res.append((dt, float(np.linalg.norm(got - want) / np.linalg.norm(want))))
    return res


def trig_bases(tol=1e-12):
    """COSINE / SINE bases: every basis pair against vectors of the unmodified
    reference (tests/golden/ops_trig_64x32.npz) and the oracle restatement, the spectral
    derivative factors, and the reference's analytic known-answer tests
    (test/SpectralTransformer_test.py:64-310, restated for this grid)."""
    import os
    from melvin import ArrayFactory, BasisFunctions, Parameters, SpatialDifferentiator, SpectralTransformer
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ops_trig_64x32.npz"))
    nx, nz = g["phys_in"].shape
    p = Parameters({"nx": nx, "nz": nz, "lx": float(g["lx"]), "lz": float(g["lz"]), "final_time": 1.0},
                   validate=False)
    G = mo.Grid(nx, nz, p.lx, p.lz)
    af = ArrayFactory(p, xp)
    st = SpectralTransformer(p, xp, af)
    sd = SpatialDifferentiator(p, xp, af)
    B = BasisFunctions
    worst = 0.0
    for bx in (B.COMPLEX_EXP, B.COSINE, B.SINE):
        for bz in (B.COMPLEX_EXP, B.COSINE, B.SINE):
            if bx is B.COMPLEX_EXP and bz is B.COMPLEX_EXP:
                continue
            tag = f"b{int(bx)}{int(bz)}"
            spec = st.to_spectral(xp.array(g["phys_in"]), basis_functions=[bx, bz])
            e = [mo.relative_l2(spec.get(), g[f"{tag}_to_spectral"]),
                 mo.relative_l2(spec.get(), mo.to_spectral_basis(G, g["phys_in"], int(bx), int(bz)))]
            held = xp.array(g["spec_in"])
            phys = st.to_physical(held, af.make_physical(), basis_functions=[bx, bz])
            e.append(mo.relative_l2(phys.get(), g[f"{tag}_to_physical"]))
            assert np.array_equal(held.get(), g["spec_in"])          # the argument is left untouched
            back = st.to_physical(spec, basis_functions=[bx, bz])
            e.append(mo.relative_l2(back.get(), g[f"{tag}_roundtrip"]))
            e.append(mo.relative_l2(np.asarray(sd.calc_lap([bx, bz])), g[f"{tag}_lap"]))
            assert max(e) < tol, (tag, e)
            worst = max(worst, max(e))
    sin_ = xp.array(g["spec_in"])
    for b in (B.COSINE, B.SINE):
        for name, fn in (("sddx", sd.sddx), ("sddz", sd.sddz), ("sd2dx2", sd.sd2dx2), ("sd2dz2", sd.sd2dz2)):
            got = fn(sin_, b)
            got = got.materialize().get() if hasattr(got, "materialize") else np.asarray(got)
            assert mo.relative_l2(got, g[f"{name}_b{int(b)}"]) < 1e-14, (name, int(b))
    # analytic known answers (reference test/SpectralTransformer_test.py)
    xe, ze = np.linspace(0, 1.0, nx), np.linspace(0, 1.0, nz)                    # end points included
    xo, zo = np.linspace(0, 1.0, nx, endpoint=False), np.linspace(0, 1.0, nz, endpoint=False)

    def check(bases, X, Z, fn, entries):
        Xg, Zg = np.meshgrid(X, Z, indexing="ij")
        true_physical = fn(Xg, Zg)
        spectral, physical = af.make_spectral(), af.make_physical()
        st.to_spectral(xp.array(true_physical), spectral, basis_functions=bases)
        true_spectral = np.zeros(spectral.shape, complex)
        for idx, val in entries.items():
            true_spectral[idx] = val
        np.testing.assert_array_almost_equal(spectral.get(), true_spectral)
        st.to_physical(spectral, physical, basis_functions=bases)
        np.testing.assert_array_almost_equal(physical.get(), true_physical)

    pi = np.pi
    check([B.COSINE, B.COMPLEX_EXP], xe, zo, lambda X, Z: np.cos(pi * X) + 2.0 * np.cos(2 * pi * X),
          {(1, 0): 1.0, (2, 0): 2.0, (-1, 0): 1.0, (-2, 0): 2.0})                         # :64-92
    check([B.COMPLEX_EXP, B.COSINE], xo, ze, lambda X, Z: np.cos(pi * Z) + 2.0 * np.cos(2 * pi * Z),
          {(0, 1): 1.0, (0, 2): 2.0})                                                     # :95-121
    check([B.COSINE, B.COSINE], xe, ze,
          lambda X, Z: np.cos(pi * Z) + 2.0 * np.cos(2 * pi * Z) + np.cos(pi * X)
          + 2.0 * np.cos(2 * pi * X) * np.cos(pi * Z),
          {(0, 1): 1.0, (0, 2): 2.0, (1, 0): 1.0, (2, 1): 2.0, (-1, 0): 1.0, (-2, 1): 2.0})   # :124-160
    check([B.SINE, B.SINE], xe, ze,
          lambda X, Z: np.sin(pi * Z) * np.sin(pi * X) + 2.0 * np.sin(2 * pi * Z) * np.sin(3 * pi * X),
          {(1, 1): 1.0, (-1, 1): -1.0, (3, 2): 2.0, (-3, 2): -2.0})                       # :162-186
    check([B.SINE, B.COMPLEX_EXP], xe, zo, lambda X, Z: np.sin(pi * X) + 2.0 * np.sin(2 * pi * X),
          {(1, 0): 1.0, (-1, 0): -1.0, (2, 0): 2.0, (-2, 0): -2.0})                       # :188-210
    check([B.SINE, B.COSINE], xe, ze, lambda X, Z: np.sin(pi * X) + 2.0 * np.sin(2 * pi * X),
          {(1, 0): 1.0, (-1, 0): -1.0, (2, 0): 2.0, (-2, 0): -2.0})                       # :212-234
    check([B.COSINE, B.SINE], xe, ze,
          lambda X, Z: np.cos(pi * X) * np.sin(pi * Z) + 2.0 * np.cos(3 * pi * X) * np.sin(2 * pi * Z)
          + 3.0 * np.sin(2 * pi * Z),
          {(1, 1): 1.0, (-1, 1): 1.0, (0, 2): 3.0, (3, 2): 2.0, (-3, 2): 2.0})            # :236-262
    return worst


def pentadiagonal_solve(nx=32, nz=96):
    """4th-order (pentadiagonal) Laplacian solve, SURVEY 8f-3 (extension, parity unpinned):
    (1) the device solution of A x = b against a long-double banded solve of the same matrices
        (LaplacianSolver.laps), random complex right-hand sides;
    (2) A @ solve(b) == b with the host matrices;
    (3) convergence: psi'' - k^2 psi = f with a smooth exact solution, the error falls ~16x per halving of dz
        where the tridiagonal solve falls ~4x."""
    import contextlib
    import io
    from melvin import ArrayFactory, BasisFunctions, LaplacianSolver, Parameters
    CE, FDM = BasisFunctions.COMPLEX_EXP, BasisFunctions.FDM

    def make(nz_, order):
        p = Parameters({"nx": nx, "nz": nz_, "lx": 2.0, "lz": 1.0, "final_time": 1.0,
                        "discretisation": ["spectral", "fdm"], "laplacian_order": order}, validate=False)
        af = ArrayFactory(p, xp)
        with contextlib.redirect_stdout(io.StringIO()):
            return p, LaplacianSolver(p, xp, [CE, FDM], array_factory=af)

    p, solver = make(nz, 4)
    rng = np.random.default_rng(7)
    rhs = rng.standard_normal(p.spectral_shape) + 1j * rng.standard_normal(p.spectral_shape)
    got = solver.solve(xp.array(rhs)).get()
    laps = solver.laps
    worst = 0.0
    for n, A in enumerate(laps):
        Ad = np.asarray(A.todense()).real.astype(np.longdouble)
        # Gaussian elimination without pivoting in long double (the operator is negative definite)
        M = Ad.copy()
        b = rhs[n].astype(np.clongdouble)
        for i in range(nz):
            for r in range(i + 1, min(i + 3, nz)):
                if M[r, i] != 0:
                    f = M[r, i] / M[i, i]
                    M[r, i:i + 3] -= f * M[i, i:i + 3]
                    b[r] -= f * b[i]
        x = np.zeros(nz, dtype=np.clongdouble)
        for i in range(nz - 1, -1, -1):
            x[i] = (b[i] - sum(M[i, j] * x[j] for j in range(i + 1, min(i + 3, nz)))) / M[i, i]
        err = np.linalg.norm((got[n] - x).astype(complex)) / np.linalg.norm(x.astype(complex))
        res = np.linalg.norm(A @ got[n] - rhs[n]) / np.linalg.norm(rhs[n])
        worst = max(worst, float(err))
        assert err < 1e-12 and res < 1e-11, (n, float(err), float(res))
    # convergence on psi = sin(3 pi z) + z^2 (1-z) + 1/2 with Dirichlet values given on the boundary rows
    errs = {2: [], 4: []}
    for order in (2, 4):
        for nzc in (65, 129, 257):
            pc_, sc = make(nzc, order)
            z = np.arange(nzc) * pc_.dz                # the grid the stencils assume (dz = lz / nz)
            kx0 = abs(1j * 2 * np.pi / pc_.lx)
            psi = np.sin(3 * np.pi * z) + z ** 2 * (1 - z) + 0.5
            d2 = -(3 * np.pi) ** 2 * np.sin(3 * np.pi * z) + 2 - 6 * z
            b = np.zeros(pc_.spectral_shape, complex)
            for n in range(pc_.nn):
                b[n] = d2 - (n * kx0) ** 2 * psi
                b[n, 0], b[n, -1] = psi[0], psi[-1]
            sol = sc.solve(xp.array(b)).get()
            errs[order].append(np.abs(sol - psi[None, :]).max())
    rate2 = np.log2(errs[2][0] / errs[2][2]) / 2
    rate4 = np.log2(errs[4][0] / errs[4][2]) / 2
    assert 1.8 < rate2 < 2.3, (rate2, errs[2])
    assert rate4 > 2.9 and errs[4][2] < errs[2][2] / 20, (rate4, errs)
    return worst, rate2, rate4


def pentadiagonal_residual(nx, nz):
    """Large grids: residual of the pentadiagonal solve on a sample of x modes against a banded
    product built with numpy (long double), random right-hand sides."""
    import contextlib
    import io
    from melvin import ArrayFactory, BasisFunctions, LaplacianSolver, Parameters
    CE, FDM = BasisFunctions.COMPLEX_EXP, BasisFunctions.FDM
    p = Parameters({"nx": nx, "nz": nz, "lx": 2.44, "lz": 1.0, "final_time": 1.0,
                    "discretisation": ["spectral", "fdm"], "laplacian_order": 4}, validate=False)
    af = ArrayFactory(p, xp)
    with contextlib.redirect_stdout(io.StringIO()):
        solver = LaplacianSolver(p, xp, [CE, FDM], array_factory=af)
    rng = np.random.default_rng(11)
    rhs = rng.standard_normal(p.spectral_shape) + 1j * rng.standard_normal(p.spectral_shape)
    x = solver.solve(xp.array(rhs)).get().astype(np.clongdouble)
    h2 = np.longdouble(p.dz) ** 2
    kx0 = np.longdouble(abs(1j * 2 * np.pi / p.lx))
    for n in sorted(set([0, 1, p.nn // 3, p.nn - 1])):
        k2 = (n * kx0) ** 2
        v = x[n]
        Ax = np.empty(nz, dtype=np.clongdouble)
        Ax[0], Ax[-1] = v[0], v[-1]
        Ax[1] = (v[0] - 2 * v[1] + v[2]) / h2 - k2 * v[1]
        Ax[-2] = (v[-3] - 2 * v[-2] + v[-1]) / h2 - k2 * v[-2]
        Ax[2:-2] = (-v[4:] / 12 + 4 * v[3:-1] / 3 - 5 * v[2:-2] / 2 + 4 * v[1:-3] / 3 - v[:-4] / 12) / h2 - k2 * v[2:-2]
        res = np.linalg.norm((Ax - rhs[n]).astype(complex)) / np.linalg.norm(rhs[n])
        # the solution is rounded to double: A amplifies that by ~ 1/dz^2
        assert res < 20 * 1.1e-16 * nz ** 2, (n, float(res))


def specialised_kernels_match_generic(nx=64, nz=64, nsteps=6):
    """The kernels specialised for the single-GPU single-scalar step (k_xfwd_scalar, the unsharded
    z stage, reductions compiled out on non-ticker steps) do the same arithmetic in the same order
    as the generic ones.  Identical bits on the host build; on the device the compiler contracts
    multiply-adds per kernel, so the states agree to rounding (1e-14 after a few KH steps)."""
    import os
    import bench
    keys = ("MLV_XFWD_GENERIC", "MLV_ZADV_GENERIC")
    saved = {k: os.environ.pop(k, None) for k in keys}
    cwd = os.getcwd()
    try:
        def run(generic):
            for k in keys:
                if generic:
                    os.environ[k] = "1"
                else:
                    os.environ.pop(k, None)
            step, o = bench.build_public_loop("kh", nx, nz)
            o["sim"].reductions = "always" if generic else "auto"
            for _ in range(nsteps):
                step()
            return o["w"].on_host()
        a, b = run(False), run(True)
        assert mo.relative_l2(a, b) < 1e-14
        return np.array_equal(a, b)
    finally:
        os.chdir(cwd)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
