"""The drop-in Python package driven on the CPU *emulation* build of the kernels
(development harness; see emu_harness.py).  Exercises the host logic -- deferred
expressions, double-buffered state, ticker semantics -- against the committed
goldens at tiny grid sizes.  The real parity tests are the ``-m gpu`` ones."""
import ctypes
import os
import shutil
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "melvin.py_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

if shutil.which("g++") is None:  # pragma: no cover
    pytest.skip("g++ not available for the emulation build", allow_module_level=True)

import emu_harness as eh  # noqa: E402
from melvin import _backend  # noqa: E402
from oracle import melvin_oracle as mo  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def emu_backend():
    saved = dict(_backend._state)
    saved_ctx = dict(_backend._contexts)
    _backend._contexts.clear()
    _backend._install(eh.lib(), "cpu")
    import melvin.b200 as b
    saved_util = b._util_ctx
    b._util_ctx = None
    yield
    _backend._contexts.clear()
    _backend._contexts.update(saved_ctx)
    _backend._state.update(saved)
    b._util_ctx = saved_util


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300)


def test_array_namespace():
    import melvin.b200 as xp
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((6, 5)), rng.standard_normal((6, 5))
    c = rng.standard_normal((6, 5)) + 1j * rng.standard_normal((6, 5))
    A, B, Cc = xp.array(a), xp.array(b), xp.array(c)
    np.testing.assert_allclose((A + B * 2.0 - 1.0).get(), a + b * 2.0 - 1.0)
    np.testing.assert_allclose((-A / B).get(), -a / b)
    np.testing.assert_allclose((A * Cc).get(), a * c)
    np.testing.assert_allclose((Cc / Cc[:, 1:2]).get(), c / c[:, 1:2])
    np.testing.assert_allclose((2.0 / A).get(), 2.0 / a)
    np.testing.assert_allclose((1j * A).get(), 1j * a)
    np.testing.assert_allclose((A ** 2 + B ** 2).get(), a ** 2 + b ** 2)
    np.testing.assert_allclose(xp.max(A), a.max())
    np.testing.assert_allclose(xp.sum(A[1:, :3]), a[1:, :3].sum())
    np.testing.assert_allclose(xp.mean(A * B), (a * b).mean())
    A[1:, -2:] = 0.0
    a[1:, -2:] = 0.0
    A[0, :] = B[0, :]
    a[0, :] = b[0, :]
    np.testing.assert_allclose(A.get(), a)
    z = xp.zeros((3, 4), dtype=np.complex128)
    z[:, 1] = 1.0 + 2.0j
    assert z.get()[2, 1] == 1.0 + 2.0j and z.dtype == np.complex128
    assert float(np.float64(3.0) * A[2, 2]) == 3.0 * a[2, 2]
    # set-up helpers the reference's own classes call on the namespace (ArrayFactory.py:8-26)
    n = xp.concatenate((xp.arange(0, 4), xp.arange(-3, 0)))
    N, M = xp.meshgrid(n, xp.arange(0, 3), indexing="ij")
    Nn, Mn = np.meshgrid(np.concatenate((np.arange(0, 4), np.arange(-3, 0))), np.arange(0, 3), indexing="ij")
    assert np.array_equal(N.get(), Nn) and np.array_equal(M.get(), Mn) and N.get().dtype == Nn.dtype
    np.testing.assert_array_equal(xp.concatenate((A, B), axis=1).get(), np.concatenate((a, b), axis=1))
    np.testing.assert_array_equal(A.take((-1, 0), axis=0).get(), a.take((-1, 0), axis=0))


def test_reference_unit_tests_run_on_the_dropin():
    """The reference's own known-answer tests (test/*_test.py) against the drop-in."""
    import melvin.b200 as xp
    from melvin import (ArrayFactory, BasisFunctions, LaplacianSolver, Parameters,
                        SpatialDifferentiator, SpectralTransformer, Variable)
    CE = BasisFunctions.COMPLEX_EXP
    p = Parameters({"nx": 64, "nz": 32, "lx": 1.0, "lz": 1.0, "final_time": 1.0}, validate=False)
    af = ArrayFactory(p, xp)
    st = SpectralTransformer(p, xp, af)
    sd = SpatialDifferentiator(p, xp, af)
    spectral, physical = af.make_spectral(), af.make_physical()
    assert spectral.shape == (2 * p.nn + 1, p.nm) and physical.shape == (p.nx, p.nz)
    x = np.linspace(0, 1.0, p.nx, endpoint=False)
    z = np.linspace(0, 1.0, p.nz, endpoint=False)
    X, Z = np.meshgrid(x, z, indexing="ij")
    true_physical = np.cos(2 * np.pi * X) + 2.0 * np.sin(2 * 2 * np.pi * Z)
    st.to_spectral(xp.array(true_physical), spectral, basis_functions=[CE, CE])
    true_spectral = np.zeros(spectral.shape, complex)
    true_spectral[1, 0] = true_spectral[-1, 0] = 0.5
    true_spectral[0, 2] = 2.0j / -2.0
    np.testing.assert_array_almost_equal(spectral.get(), true_spectral)
    st.to_physical(spectral, physical, basis_functions=[CE, CE])
    np.testing.assert_array_almost_equal(physical.get(), true_physical)
    # the third slot of the reference's transformer (SpectralTransformer.py:21-24): truncated <-> full layout
    rng = np.random.default_rng(1)
    full = rng.standard_normal((p.nx, p.nz // 2 + 1)) + 1j * rng.standard_normal((p.nx, p.nz // 2 + 1))
    trunc = xp.zeros(spectral.shape, dtype=np.complex128)
    st._scale(xp.array(full), trunc)
    want = np.zeros(spectral.shape, complex)
    want[: p.nn + 1, : p.nm] = full[: p.nn + 1, : p.nm]
    want[-p.nn:, : p.nm] = full[-p.nn:, : p.nm]
    assert np.array_equal(trunc.get(), want)
    # Variable_test.py:16-46, :76-100
    var = Variable(p, xp, st=st, sd=sd, array_factory=af, basis_functions=[CE, CE])
    var.setp(np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Z))
    assert var.pddx().get()[:, 0] == pytest.approx(-2 * np.pi * np.sin(2 * np.pi * x), rel=2e-2, abs=1e-2)
    var[:] = 1
    n, m = af.make_mode_number_matrices()
    np.testing.assert_array_almost_equal(
        var.snabla2().get(), -((2 * np.pi / p.lx * n) ** 2) - (2 * np.pi / p.lz * m) ** 2)
    # FDM: Variable_test.py:49-73, LaplacianSolver_test.py:37-59, SpectralTransformer_test.py:313-348
    pf = Parameters({"nx": 64, "nz": 32, "lx": 1.0, "lz": 1.0, "final_time": 1.0,
                     "discretisation": ["spectral", "fdm"]}, validate=False)
    aff = ArrayFactory(pf, xp)
    sdf = SpatialDifferentiator(pf, xp, array_factory=aff)
    stf = SpectralTransformer(pf, xp, aff)
    assert aff.make_spectral().shape == (pf.nn, pf.nz)
    v = Variable(pf, xp, sd=sdf, array_factory=aff, basis_functions=[CE, CE])
    v[1] = 0.5 * z ** 2
    nab = v.snabla2().get()
    np.testing.assert_array_almost_equal(nab[1, 1:-1], (-((2 * np.pi) ** 2) * 0.5 * z ** 2 + 1)[1:-1])
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        solver = LaplacianSolver(pf, xp, [CE, BasisFunctions.FDM], array_factory=aff)
    true_soln = np.ones(pf.spectral_shape, complex)
    rhs = np.array([lap @ true_soln[k] for k, lap in enumerate(solver.laps)])
    np.testing.assert_array_almost_equal(solver.solve(xp.array(rhs)).get(), true_soln)
    phys = 3.0 + np.cos(2 * np.pi * X) + 2.0 * np.cos(2 * 2 * np.pi * X)
    spec = stf.to_spectral(xp.array(phys), basis_functions=[CE, BasisFunctions.FDM]).get()
    assert np.allclose(spec[0], 3.0) and np.allclose(spec[1], 0.5) and np.allclose(spec[2], 1.0)
    with pytest.raises(Exception):
        stf.to_spectral(xp.array(phys), basis_functions=[CE, CE])


def test_taylor_green_loop_matches_golden():
    import parity_cases as pc
    gl = golden("loop_tg_64x64.npz")
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    before = _backend.call_counts()
    with pc.scratch_cwd():
        out = pc.run_single_scalar(64, 64, float(gl["lx"]), float(gl["lz"]), float(gl["coef"]),
                                   float(gl["dt"]), 20, mo.ic_taylor_green(g), snaps=(1, 2, 10, 20),
                                   strict_reads=True)
    after = _backend.call_counts()
    used = {k: after.get(k, 0) - before.get(k, 0) for k in after}
    # the fused path ran: one inverse x pass, one fused z stage and one forward x pass
    # (with the integrator in its epilogue) per step, nothing through the eager kernels
    assert used["mlv_x_inverse"] == 20 and used["mlv_advect_z"] == 20 and used["mlv_x_forward"] == 20
    assert used.get("mlv_advect_phys", 0) == 0 and used.get("mlv_integrate", 0) == 0
    for k in (1, 2, 10, 20):
        assert rel(out[f"w_step{k}"], gl[f"w_step{k}"]) < 1e-12
    np.testing.assert_allclose(out["ke"], gl["ke"], rtol=1e-11)
    np.testing.assert_allclose(out["ke_t"], gl["ke_t"], rtol=1e-14)
    # deferred psi / ux read *after* the first update still show the pre-update fields
    w0 = mo.to_spectral(g, mo.ic_taylor_green(g))
    vel = mo.velocity_from_vorticity(g, w0)
    assert rel(out["psi_after_step1"], vel["psi_s"]) < 1e-13
    assert rel(out["ux_p_after_step1"], vel["ux_p"]) < 1e-12


def test_order4_ab4_loop_matches_golden():
    import parity_cases as pc
    gl = golden("loop_tg_64x64_o4_ab4.npz")
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    with pc.scratch_cwd():
        out = pc.run_single_scalar(64, 64, float(gl["lx"]), float(gl["lz"]), float(gl["coef"]),
                                   float(gl["dt"]), 12, mo.ic_taylor_green(g), snaps=(1, 5, 12),
                                   order=4, int_order=4)
    for k in (1, 5, 12):
        assert rel(out[f"w_step{k}"], gl[f"w_step{k}"]) < 1e-12
    np.testing.assert_allclose(out["ke"], gl["ke"], rtol=1e-11)


def test_ddc_loop_matches_golden():
    import parity_cases as pc
    gl = golden("loop_ddc_64x64.npz")
    with pc.scratch_cwd():
        out = pc.run_ddc(64, 64, float(gl["lx"]), float(gl["lz"]), float(gl["dt"]), 10,
                         float(gl["Pr"]), float(gl["R0"]), float(gl["tau"]), snaps=(1, 10))
    for k in (1, 10):
        for nm in ("w", "tmp", "xi"):
            assert rel(out[f"{nm}_step{k}"], gl[f"{nm}_step{k}"]) < 1e-12, (nm, k)
    np.testing.assert_allclose(out["ke"], gl["ke"][:10], rtol=1e-10)
    np.testing.assert_allclose(out["nu"] - 1, gl["nu"][:10] - 1, rtol=1e-9, atol=1e-18)


def test_tearing_loop_matches_golden():
    import parity_cases as pc
    gl = golden("loop_tearing_64x64.npz")
    with pc.scratch_cwd():
        out = pc.run_tearing(64, 64, float(gl["lx"]), float(gl["lz"]), float(gl["dt"]), 10,
                             float(gl["Re"]), float(gl["S"]), gl["j0_phys"], snaps=(1, 10))
    for k in (1, 10):
        assert rel(out[f"w_step{k}"], gl[f"w_step{k}"]) < 1e-10
        assert rel(out[f"j_step{k}"], gl[f"j_step{k}"]) < 1e-12
    np.testing.assert_allclose(out["ke"], gl["ke"][:10], rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("order,ab", [(2, 2), (4, 4)])
def test_rbc_fdm_loop_matches_golden(order, ab):
    import parity_cases as pc
    gl = golden(f"loop_rbc_64x32_o{order}_ab{ab}.npz")
    before = _backend.call_counts()
    with pc.scratch_cwd():
        out = pc.run_rbc(64, 32, order, ab, float(gl["dt"]), 10, float(gl["Pr"]), float(gl["Ra"]),
                         snaps=(1, 10))
    after = _backend.call_counts()
    used = {k: after.get(k, 0) - before.get(k, 0) for k in after}
    # the fused three-kernel step ran: one solve+velocities, one fused advection and one row-wise
    # right-hand side + update per scalar and step; nothing through the eager stencil kernels
    assert used["mlv_fdm_velocity"] == 10 and used["mlv_fdm_advect"] == 20 and used["mlv_integrate"] == 20
    assert not any(used.get(k, 0) for k in ("mlv_stencil", "mlv_advect_phys", "mlv_solve_fdm", "mlv_to_physical"))
    for k in (1, 10):
        for nm in ("w", "tmp", "psi"):
            assert rel(out[f"{nm}_step{k}"], gl[f"{nm}_step{k}"]) < 1e-10, (nm, k)
    np.testing.assert_allclose(out["ke"], gl["ke"][:10], rtol=1e-9)


# ---- orderings that only work if deferred work tracks the buffers it reads
@pytest.mark.parametrize("inplace_write", [False, True])
def test_rhs_assigned_before_its_source_changes(inplace_write):
    import aliasing_cases as ac
    got, want = ac.swapped_integrate_order(inplace_write=inplace_write)
    for k in want:
        assert rel(got[k], want[k]) < 1e-12, k


def test_augmented_assignment_and_view_writes_flush_dependants():
    import aliasing_cases as ac
    got, want = ac.augmented_assignment_after_velocity()
    for k in want:
        assert rel(got[k], want[k]) < 1e-12, k


def test_held_handle_stays_current():
    import aliasing_cases as ac
    same, rows_ok, col0 = ac.held_handle_stays_current()
    assert all(same) and all(rows_ok)
    assert np.all(col0 == 0.0)


# ---- restart and the Adams-Moulton formulas (a14)
def test_restart_roundtrip_is_bit_exact():
    import host_cases as hc
    (lb, la), wb, wa = hc.restart_roundtrip()
    assert lb == la and np.array_equal(wb, wa)


@pytest.mark.parametrize("n_from,n_to", [(64, 128), (64, 32)])
def test_restart_at_a_different_resolution(n_from, n_to):
    import host_cases as hc
    w_old, h_old, w_new, h_new, meta, ok = hc.restart_resolution_change(n_from, n_to)
    nn, nm = (n_to - 1) // 3, (n_to - 1) // 3
    assert np.array_equal(w_new, hc.expected_rescale(w_old, nn, nm))
    for k in range(h_old.shape[0]):
        assert np.array_equal(h_new[k], hc.expected_rescale(h_old[k], nn, nm))
    assert meta[0] == 3 and meta[2] == meta[3] and ok
    assert np.linalg.norm(w_new) > 0


@pytest.mark.parametrize("order", [2, 4])
def test_adams_moulton_corrector_formulas(order):
    import host_cases as hc
    got_c, want_c, got_p, want_p = hc.corrector_formulas(order)
    assert rel(got_c, want_c) < 1e-14 and rel(got_p, want_p) < 1e-14


@pytest.mark.parametrize("order", [2, 4])
def test_predictor_corrector_converges_at_its_order(order):
    """SURVEY 8f-3 (parity unpinned: no reference implementation): AB/AM PECE step"""
    import host_cases as hc
    res = hc.predictor_corrector_convergence(order)
    rates = [np.log2(res[i][1] / res[i + 1][1]) for i in range(len(res) - 1)]
    assert all(r > order - 0.35 for r in rates), (res, rates)
    assert res[-1][1] < (1e-4 if order == 2 else 1e-8)


def test_cosine_and_sine_bases():
    import host_cases as hc
    hc.trig_bases()


def test_reductions_follow_the_tickers():
    """The fused z stage computes the CFL / energy partials only for steps whose end_loop fires a
    ticker (Simulation.end_loop); a reader on any other step falls back to an explicit reduction
    and gets the same numbers."""
    import bench
    from melvin.utility import calc_kinetic_energy
    import melvin.b200 as xp
    cwd = os.getcwd()
    try:
        step, o = bench.build_public_loop("kh", 32, 32)
        sim, ctx, ux, uz = o["sim"], o["sim"]._ctx, o["ux"], o["uz"]
        on = []
        for _ in range(23):
            on.append(ctx.want_reductions)
            step()
        # loop 1 fires every ticker; then the CFL ticker every 10 loops (11, 21), tracker every 100
        assert [i for i, f in enumerate(on) if f] == [0, 10, 20]
        assert ux._red is None                              # the last step skipped them
        ke_fallback = float(calc_kinetic_energy(ux, uz, xp, o["params"]))
        sim.reductions = "always"
        sim.end_loop.__self__._ctx.want_reductions = True
        step2, o2 = bench.build_public_loop("kh", 32, 32)
        o2["sim"].reductions = "always"
        for _ in range(23):
            step2()
        assert o2["ux"]._red is not None
        ke_fused = float(calc_kinetic_energy(o2["ux"], o2["uz"], xp, o2["params"]))
        assert abs(ke_fallback / ke_fused - 1) < 1e-13
        assert np.array_equal(o["w"].on_host(), o2["w"].on_host())       # same trajectory, same dt history
    finally:
        os.chdir(cwd)


def test_pentadiagonal_fourth_order_solve():
    import host_cases as hc
    worst, r2, r4 = hc.pentadiagonal_solve()
    assert worst < 1e-12


def test_specialised_kernels_match_generic():
    import host_cases as hc
    assert hc.specialised_kernels_match_generic()          # bit-identical on the host build
