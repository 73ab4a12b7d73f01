"""Pin the CPU oracle: every oracle function against vectors produced by the
unmodified reference (tests/golden/make_golden.py) and against the analytic
known-answer tests of the reference's own suite (reference test/*_test.py).
CPU only."""
import numpy as np
import pytest
from numpy.testing import assert_array_almost_equal

from conftest import golden, rel_l2
from oracle import melvin_oracle as mo

TOL = 1e-13   # same library FFT, same operation order: rounding only


# ---------------------------------------------------------------- operators
@pytest.fixture(scope="module")
def gs():
    return golden("ops_spectral_64x32.npz")


def grid_s(gs, order=2):
    return mo.Grid(64, 32, float(gs["lx"]), float(gs["lz"]), fd_order=order)


def test_transforms_2d(gs):
    g = grid_s(gs)
    assert rel_l2(mo.to_spectral(g, gs["phys_in"]), gs["to_spectral"]) < TOL
    assert rel_l2(mo.to_physical(g, gs["to_spectral"]), gs["to_physical"]) < TOL
    # random non-Hermitian spectrum: Im of the m=0 column is dropped (F4)
    assert rel_l2(mo.to_physical(g, gs["spec_rand_in"]), gs["to_physical_rand"]) < TOL


def test_spectral_derivatives(gs):
    g = grid_s(gs)
    s = gs["to_spectral"]
    for name in ("sddx", "sddz", "sd2dx2", "sd2dz2"):
        assert rel_l2(getattr(mo, name)(g, s), gs[name]) < TOL, name
    assert rel_l2(mo.lap_symbol(g), gs["lap"]) < TOL
    assert rel_l2(mo.solve_spectral(g, s), gs["solve"]) < TOL


@pytest.mark.parametrize("order", [2, 4])
def test_stencils_and_advection(gs, order):
    g = grid_s(gs, order)
    assert rel_l2(mo.pddx(g, gs["phys_in"]), gs[f"pddx_o{order}"]) < TOL
    assert rel_l2(mo.pddz(g, gs["phys_in"]), gs[f"pddz_o{order}"]) < TOL
    nl, qp = mo.vec_dot_nabla(g, gs["to_spectral"], gs["adv_ux_p"], gs["adv_uz_p"])
    assert rel_l2(nl, gs[f"vec_dot_nabla_o{order}"]) < TOL
    assert rel_l2(qp, gs[f"vec_dot_nabla_qp_o{order}"]) < TOL


def test_velocity_from_vorticity(gs):
    g = grid_s(gs)
    v = mo.velocity_from_vorticity(g, gs["to_spectral"])
    for k in ("psi_s", "ux_s", "uz_s", "ux_p", "uz_p"):
        assert rel_l2(v[k], gs["vel_" + k]) < TOL, k


@pytest.fixture(scope="module")
def gf():
    return golden("ops_fdm_64x32.npz")


@pytest.mark.parametrize("order", [2, 4])
def test_fdm_operators(gf, order):
    g = mo.Grid(64, 32, float(gf["lx"]), float(gf["lz"]), fdm_z=True, fd_order=order)
    assert rel_l2(mo.to_spectral(g, gf["phys_in"]), gf["to_spectral"]) < TOL
    assert rel_l2(mo.to_physical(g, gf["spec_rand_in"]), gf["to_physical_rand"]) < TOL
    assert rel_l2(mo.sddx(g, gf["to_spectral"]), gf["sddx"]) < TOL
    assert rel_l2(mo.sd2dx2(g, gf["to_spectral"]), gf["sd2dx2"]) < TOL
    assert rel_l2(mo.pddx(g, gf["phys_in"]), gf[f"pddx_o{order}"]) < TOL
    assert rel_l2(mo.pddz(g, gf["phys_in"]), gf[f"pddz_o{order}"]) < TOL
    assert rel_l2(mo.sd2dz2(g, gf["spec_rand_in"]), gf[f"sd2dz2_o{order}"]) < TOL
    assert rel_l2(mo.snabla2(g, gf["to_spectral"]), gf[f"snabla2_o{order}"]) < TOL
    nl, _ = mo.vec_dot_nabla(g, gf["to_spectral"], gf["adv_ux_p"], gf["adv_uz_p"])
    assert rel_l2(nl, gf[f"vec_dot_nabla_o{order}"]) < TOL
    # SuperLU (reference) vs Thomas (oracle): rounding only at nz=32 (F8)
    assert rel_l2(mo.solve_fdm(g, gf["spec_rand_in"]), gf["solve"]) < 1e-12
    v = mo.velocity_from_vorticity(g, gf["spec_rand_in"])
    assert rel_l2(v["psi_s"], gf[f"vel_psi_s_o{order}"]) < 1e-12
    assert rel_l2(v["uz_s"], gf[f"vel_uz_s_o{order}"]) < 1e-12
    assert rel_l2(v["ux_p"], gf[f"vel_ux_p_o{order}"]) < 1e-12
    assert rel_l2(v["uz_p"], gf[f"vel_uz_p_o{order}"]) < 1e-12


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("kind", ["si", "ex"])
def test_integrators(order, kind):
    gi = golden("integrator_32x32.npz")
    tag = f"o{order}_{kind}"
    g = mo.Grid(32, 32, 1.0, 1.0, int_order=order)
    q = gi[f"q0_{tag}"].copy()
    hist = mo.History(g)
    dt = 1e-2
    for k in range(gi[f"rhs_{tag}"].shape[0]):
        hist.set_current(gi[f"rhs_{tag}"][k])
        if kind == "si":
            q = mo.integrate_semi_implicit(g, q, hist, gi[f"third_{tag}"], dt)
        else:
            q = mo.integrate_explicit(g, q, hist, gi[f"third_{tag}"], dt)
        if k == 2:
            dt = 0.9e-2
        assert rel_l2(q, gi[f"states_{tag}"][k]) < TOL, (tag, k)


# ------------------------------------------------- reference known answers
def test_kat_transform_periodic():
    """reference test/SpectralTransformer_test.py:27-61"""
    g = mo.Grid(256, 256, 1.0, 1.0)
    x = np.linspace(0, 1.0, g.nx, endpoint=False)
    z = np.linspace(0, 1.0, g.nz, endpoint=False)
    X, Z = np.meshgrid(x, z, indexing="ij")
    phys = np.cos(2 * np.pi * X) + 2.0 * np.sin(2 * 2 * np.pi * Z)
    spec = mo.to_spectral(g, phys)
    true = np.zeros_like(spec)
    true[1, 0] = 0.5
    true[-1, 0] = 0.5
    true[0, 2] = 2.0j / -2.0
    assert_array_almost_equal(spec, true)
    assert_array_almost_equal(mo.to_physical(g, spec), phys)


def test_kat_transform_1d_fdm():
    """reference test/SpectralTransformer_test.py:313-348"""
    g = mo.Grid(256, 256, 1.0, 1.0, fdm_z=True)
    x = np.linspace(0, 1.0, g.nx, endpoint=False)
    z = np.linspace(0, 1.0, g.nz, endpoint=False)
    X, _ = np.meshgrid(x, z, indexing="ij")
    phys = 3.0 + np.cos(2 * np.pi * X) + 2.0 * np.cos(2 * 2 * np.pi * X)
    spec = mo.to_spectral(g, phys)
    true = np.zeros_like(spec)
    true[0, :] = 3.0
    true[1, :] = 0.5
    true[2, :] = 1.0
    assert_array_almost_equal(spec, true)
    assert_array_almost_equal(mo.to_physical(g, spec), phys)


def test_kat_stencils_and_nabla2():
    """reference test/Variable_test.py:16-100"""
    g = mo.Grid(256, 256, 1.0, 1.0)
    x = np.linspace(0, g.lx, g.nx, endpoint=False)
    z = np.linspace(0, g.lz, g.nz, endpoint=False)
    X, Z = np.meshgrid(x, z, indexing="ij")
    f = np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Z)
    assert mo.pddx(g, f)[:, 0] == pytest.approx(-2 * np.pi * np.sin(2 * np.pi * x), rel=1e-3)
    assert mo.pddz(g, f)[0, :] == pytest.approx(-2 * np.pi * np.sin(2 * np.pi * z), rel=1e-3)
    n, m = mo.mode_numbers(g)
    ones = np.ones(g.spectral_shape, dtype=complex)
    assert_array_almost_equal(mo.snabla2(g, ones),
                              -((2 * np.pi / g.lx * n) ** 2) - (2 * np.pi / g.lz * m) ** 2 + 0 * ones)
    gf = mo.Grid(256, 256, 1.0, 1.0, fdm_z=True)
    var = np.zeros(gf.spectral_shape, dtype=complex)
    var[1] = 0.5 * z ** 2
    true = -((2 * np.pi) ** 2) * var[1] + 1
    assert_array_almost_equal(mo.snabla2(gf, var)[1, 1:-1], true[1:-1])


def test_kat_fdm_solver_inverts_its_matrix():
    """reference test/LaplacianSolver_test.py:37-59 (with the true x factor)"""
    g = mo.Grid(256, 256, 1.0, 1.0, fdm_z=True)
    lo, di, up = mo.fdm_tridiagonal(g)
    true = np.ones(g.spectral_shape, dtype=complex)
    rhs = di * true
    rhs[:, 1:] += lo[:, 1:] * true[:, :-1]
    rhs[:, :-1] += up[:, :-1] * true[:, 1:]
    assert_array_almost_equal(mo.solve_fdm(g, rhs), true)


# ------------------------------------------------------------- whole loops
def _run_single(gl, ic):
    g = mo.Grid(int(gl["nx"]), int(gl["nz"]), float(gl["lx"]), float(gl["lz"]),
                fd_order=int(gl["order"]), int_order=int(gl["int_order"]))
    nsteps = int(gl["nsteps"])
    snaps = [int(k[6:]) for k in gl.files if k.startswith("w_step")]
    w, run, got = mo.run_single_scalar(g, ic(g), float(gl["coef"]), float(gl["dt"]),
                                       nsteps, tracker_cadence=1, snapshots=snaps)
    return g, w, run, got


@pytest.mark.parametrize("name,ic", [
    ("loop_tg_64x64.npz", mo.ic_taylor_green),
    ("loop_tg_64x64_o4_ab4.npz", mo.ic_taylor_green),
    ("loop_kh_128x64.npz", mo.ic_kelvin_helmholtz),
])
def test_loop_single_scalar(name, ic):
    gl = golden(name)
    g, w, run, got = _run_single(gl, ic)
    for k, v in got.items():
        assert rel_l2(v, gl[f"w_step{k}"]) < 1e-12, (name, k)
    assert rel_l2(w, gl["w_final"]) < 1e-12
    np.testing.assert_allclose(run.ke, gl["ke"], rtol=1e-12)
    np.testing.assert_allclose(run.times, gl["ke_t"], rtol=1e-14)


def test_loop_config1_1000_steps():
    """BASELINE config 1: Taylor-Green 256^2, AB2 semi-implicit, 1000 steps."""
    gl = golden("loop_tg_256x256_1000.npz")
    g, w, run, got = _run_single(gl, mo.ic_taylor_green)
    for k, v in got.items():
        assert rel_l2(v, gl[f"w_step{k}"]) < 1e-12, k
    np.testing.assert_allclose(run.ke, gl["ke"], rtol=1e-12)
    ratio = run.ke[-1] / run.ke[0]
    assert abs(ratio - np.exp(-(run.times[-1] - run.times[0]))) < 2e-3


def test_loop_ddc():
    gl = golden("loop_ddc_64x64.npz")
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    run = mo.Run(g, float(gl["dt"]), tracker_cadence=1)
    noise = mo.ic_noise(g)
    state = tuple(mo.to_spectral(g, noise) for _ in range(3))
    hists = tuple(mo.History(g) for _ in range(3))
    for _ in range(int(gl["nsteps"])):
        state = mo.step_double_diffusive(g, run, state, hists, float(gl["Pr"]),
                                         float(gl["R0"]), float(gl["tau"]))
        if f"w_step{run.loop}" in gl.files:
            for nm, arr in zip(("w", "tmp", "xi"), state):
                assert rel_l2(arr, gl[f"{nm}_step{run.loop}"]) < 1e-12, (nm, run.loop)
    np.testing.assert_allclose(run.ke, gl["ke"], rtol=1e-11)
    np.testing.assert_allclose(np.array(run.extra) - 1, gl["nu"] - 1, rtol=1e-9, atol=1e-18)


def test_loop_tearing():
    gl = golden("loop_tearing_64x64.npz")
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    run = mo.Run(g, float(gl["dt"]), tracker_cadence=1)
    state = (np.zeros(g.spectral_shape, complex), mo.to_spectral(g, gl["j0_phys"]))
    hists = (mo.History(g), mo.History(g))
    for _ in range(int(gl["nsteps"])):
        state = mo.step_tearing(g, run, state, hists, float(gl["Re"]), float(gl["S"]))
        if f"w_step{run.loop}" in gl.files:
            assert rel_l2(state[0], gl[f"w_step{run.loop}"]) < 1e-11
            assert rel_l2(state[1], gl[f"j_step{run.loop}"]) < 1e-12
    np.testing.assert_allclose(run.ke, gl["ke"], rtol=1e-10, atol=1e-300)


@pytest.mark.parametrize("order,ab", [(2, 2), (4, 4)])
def test_loop_rbc(order, ab):
    gl = golden(f"loop_rbc_64x32_o{order}_ab{ab}.npz")
    g = mo.Grid(64, 32, 2.44, 1.0, fdm_z=True, fd_order=order, int_order=ab,
                integrator="explicit")
    run = mo.Run(g, float(gl["dt"]), tracker_cadence=1)
    w = mo.to_spectral(g, mo.ic_noise(g))
    tmp = mo.to_spectral(g, mo.ic_rbc_temperature(g))
    state = (w, tmp, np.zeros(g.spectral_shape, complex))
    hists = (mo.History(g), mo.History(g))
    for _ in range(int(gl["nsteps"])):
        state = mo.step_rayleigh_benard(g, run, state, hists, float(gl["Pr"]), float(gl["Ra"]))
        if f"w_step{run.loop}" in gl.files:
            for nm, arr in zip(("w", "tmp", "psi"), state):
                assert rel_l2(arr, gl[f"{nm}_step{run.loop}"]) < 1e-10, (nm, run.loop)
    np.testing.assert_allclose(run.ke, gl["ke"], rtol=1e-9)


def test_cosine_sine_bases_vs_reference_vectors():
    """oracle restatement of the mirrored rfft2 / irfft2 (SpectralTransformer.py:90-199) and of the
    basis diff factors (BasisFunctions.py:26-61) against vectors of the unmodified reference"""
    g = golden("ops_trig_64x32.npz")
    G = mo.Grid(64, 32, float(g["lx"]), float(g["lz"]))
    n, m = mo.mode_numbers(G) if hasattr(mo, "mode_numbers") else (None, None)
    for bx in range(3):
        for bz in range(3):
            if bx == 0 and bz == 0:
                continue
            tag = f"b{bx}{bz}"
            spec = mo.to_spectral_basis(G, g["phys_in"], bx, bz)
            assert mo.relative_l2(spec, g[f"{tag}_to_spectral"]) < 1e-15
            assert mo.relative_l2(mo.to_physical_basis(G, g["spec_in"], bx, bz), g[f"{tag}_to_physical"]) < 1e-15
            assert mo.relative_l2(mo.to_physical_basis(G, spec, bx, bz), g[f"{tag}_roundtrip"]) < 1e-15
    nn_ = np.concatenate((np.arange(0, G.nn + 1), np.arange(-G.nn, 0)))[:, None]
    mm_ = np.arange(0, G.nm)[None, :]
    for b in (1, 2):
        assert mo.relative_l2(mo.diff_factor(b, G.lx) * nn_ * g["spec_in"], g[f"sddx_b{b}"]) < 1e-15
        assert mo.relative_l2(mo.diff_factor(b, G.lz) * mm_ * g["spec_in"], g[f"sddz_b{b}"]) < 1e-15
        assert mo.relative_l2(mo.diff2_factor(b, G.lx) * nn_ ** 2 * g["spec_in"], g[f"sd2dx2_b{b}"]) < 1e-15
        assert mo.relative_l2(mo.diff2_factor(b, G.lz) * mm_ ** 2 * g["spec_in"], g[f"sd2dz2_b{b}"]) < 1e-15
        assert mo.relative_l2(mo.diff2_factor(b, G.lx) * nn_ ** 2 + mo.diff2_factor(3 - b, G.lz) * mm_ ** 2,
                              g[f"b{b}{3 - b}_lap"]) < 1e-15
