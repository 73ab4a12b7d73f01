#!/usr/bin/env python3
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the development container only (the reference checkout does not exist
on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

It imports the reference package (NumPy backend, ``precision="double"``), drives
the reference's own classes with the loop bodies of the example scripts
(cited per function) and stores inputs + outputs as small ``.npz`` files.
Nothing from the reference is copied into the repository: only numbers.
"""
import os
import sys
import tempfile
from functools import partial

import numpy as np

REF = next((a for a in sys.argv[1:] if not a.startswith("--")), "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import scipy  # noqa: E402
from melvin import (  # noqa: E402  (the reference package)
    ArrayFactory,
    BasisFunctions,
    Integrator,
    LaplacianSolver,
    Parameters,
    Simulation,
    SpatialDifferentiator,
    SpectralTransformer,
    TimeDerivative,
    Variable,
)
from melvin.utility import (  # noqa: E402
    calc_kinetic_energy,
    calc_velocity_from_vorticity,
    init_var_with_noise,
    sech,
)

CE = BasisFunctions.COMPLEX_EXP
FDM = BasisFunctions.FDM
VERSIONS = np.array(
    [f"numpy {np.__version__}", f"scipy {scipy.__version__}",
     f"python {sys.version.split()[0]}"]
)


def base_params(nx, nz, lx, lz, **kw):
    d = {"nx": nx, "nz": nz, "lx": lx, "lz": lz, "final_time": 1.0,
         "precision": "double"}
    d.update(kw)
    return d


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, versions=VERSIONS, **arrays)
    print(f"wrote {name}: {os.path.getsize(path)/1024:.1f} KiB")


# ---------------------------------------------------------------------------
# operator-level vectors
# ---------------------------------------------------------------------------
def ops_fully_spectral(nx, nz, lx, lz):
    out = {}
    rng = np.random.default_rng(1234)
    for order in (2, 4):
        p = Parameters(base_params(nx, nz, lx, lz,
                                   spatial_derivative_order=order),
                       validate=False)
        af = ArrayFactory(p, np)
        st = SpectralTransformer(p, np, af)
        sd = SpatialDifferentiator(p, np, af)
        basis = [CE, CE]

        def mkvar():
            return Variable(p, np, sd=sd, st=st, array_factory=af,
                            basis_functions=basis, dump_name="v")

        if order == 2:
            phys = rng.standard_normal((nx, nz))
            spec_rand = (rng.standard_normal(p.spectral_shape)
                         + 1j * rng.standard_normal(p.spectral_shape))
            out["phys_in"] = phys
            out["spec_rand_in"] = spec_rand
            out["to_spectral"] = st.to_spectral(phys, basis_functions=basis).copy()
            # random (non-Hermitian) spectrum: pins the m=0 projection (F4)
            out["to_physical_rand"] = st.to_physical(spec_rand.copy(),
                                                      basis_functions=basis).copy()
            spec = out["to_spectral"]
            out["to_physical"] = st.to_physical(spec.copy(),
                                                basis_functions=basis).copy()
            out["sddx"] = sd.sddx(spec, CE)
            out["sddz"] = sd.sddz(spec, CE)
            out["sd2dx2"] = sd.sd2dx2(spec, CE)
            out["sd2dz2"] = sd.sd2dz2(spec, CE)
            out["lap"] = sd.calc_lap(basis)
            solver = LaplacianSolver(p, np, basis, spatial_diff=sd,
                                     array_factory=af)
            out["solve"] = solver.solve(spec.copy()).copy()
            w, psi, ux, uz = mkvar(), mkvar(), mkvar(), mkvar()
            w.sets(spec)
            calc_velocity_from_vorticity(w, psi, ux, uz, solver)
            out["vel_psi_s"] = psi.gets().copy()
            out["vel_ux_s"] = ux.gets().copy()
            out["vel_uz_s"] = uz.gets().copy()
            out["vel_ux_p"] = ux.getp().copy()
            out["vel_uz_p"] = uz.getp().copy()
            ux_p = rng.standard_normal((nx, nz))
            uz_p = rng.standard_normal((nx, nz))
            out["adv_ux_p"] = ux_p
            out["adv_uz_p"] = uz_p
        phys = out["phys_in"]
        out[f"pddx_o{order}"] = sd.pddx(phys).copy()
        out[f"pddz_o{order}"] = sd.pddz(phys).copy()
        q = mkvar()
        q.sets(out["to_spectral"])
        out[f"vec_dot_nabla_o{order}"] = q.vec_dot_nabla(
            out["adv_ux_p"], out["adv_uz_p"]).copy()
        out[f"vec_dot_nabla_qp_o{order}"] = q.getp().copy()
    save(f"ops_spectral_{nx}x{nz}.npz", lx=lx, lz=lz, **out)


def ops_fdm(nx, nz, lx, lz):
    out = {}
    rng = np.random.default_rng(4321)
    for order in (2, 4):
        p = Parameters(base_params(nx, nz, lx, lz,
                                   spatial_derivative_order=order,
                                   discretisation=["spectral", "fdm"],
                                   integrator="explicit"),
                       validate=False)
        af = ArrayFactory(p, np)
        st = SpectralTransformer(p, np, af)
        sd = SpatialDifferentiator(p, np, af)
        basis = [CE, FDM]

        def mkvar():
            return Variable(p, np, sd=sd, st=st, array_factory=af,
                            basis_functions=basis, dump_name="v")

        if order == 2:
            phys = rng.standard_normal((nx, nz))
            out["phys_in"] = phys
            out["to_spectral"] = st.to_spectral(phys, basis_functions=basis).copy()
            spec_rand = (rng.standard_normal(p.spectral_shape)
                         + 1j * rng.standard_normal(p.spectral_shape))
            out["spec_rand_in"] = spec_rand
            out["to_physical_rand"] = st.to_physical(
                spec_rand.copy(), basis_functions=basis).copy()
            out["sddx"] = sd.sddx(out["to_spectral"], CE)
            out["sd2dx2"] = sd.sd2dx2(out["to_spectral"], CE)
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                solver = LaplacianSolver(p, np, basis, spatial_diff=sd,
                                         array_factory=af)
            out["solve"] = solver.solve(spec_rand.copy()).copy()
            out["adv_ux_p"] = rng.standard_normal((nx, nz))
            out["adv_uz_p"] = rng.standard_normal((nx, nz))
        spec = out["to_spectral"]
        out[f"pddx_o{order}"] = sd.pddx(out["phys_in"]).copy()
        out[f"pddz_o{order}"] = sd.pddz(out["phys_in"]).copy()
        out[f"sd2dz2_o{order}"] = sd.sd2dz2(out["spec_rand_in"], FDM).copy()
        q = mkvar()
        q.sets(spec)
        out[f"snabla2_o{order}"] = q.snabla2().copy()
        out[f"vec_dot_nabla_o{order}"] = q.vec_dot_nabla(
            out["adv_ux_p"], out["adv_uz_p"]).copy()
        w, psi, ux, uz = mkvar(), mkvar(), mkvar(), mkvar()
        w.sets(out["spec_rand_in"])
        calc_velocity_from_vorticity(w, psi, ux, uz, solver)
        out[f"vel_psi_s_o{order}"] = psi.gets().copy()
        out[f"vel_uz_s_o{order}"] = uz.gets().copy()
        out[f"vel_ux_p_o{order}"] = ux.getp().copy()
        out[f"vel_uz_p_o{order}"] = uz.getp().copy()
    save(f"ops_fdm_{nx}x{nz}.npz", lx=lx, lz=lz, **out)


def integrator_vectors(nx, nz):
    """Integrator.py:5-18,53-63 driven with random RHS levels."""
    out = {}
    rng = np.random.default_rng(99)
    for order in (2, 4):
        for kind in ("semi-implicit", "explicit"):
            p = Parameters(base_params(nx, nz, 1.0, 1.0,
                                       integrator_order=order, integrator=kind,
                                       initial_dt=1e-2), validate=False)
            af = ArrayFactory(p, np)
            sd = SpatialDifferentiator(p, np, af)
            integ = Integrator(p, np)
            var = Variable(p, np, sd=sd, array_factory=af,
                           basis_functions=[CE, CE], dump_name="v")
            dvar = TimeDerivative(p, np)
            shape = p.spectral_shape
            q0 = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
            var.sets(q0)
            nsteps = 6
            rhs = (rng.standard_normal((nsteps,) + shape)
                   + 1j * rng.standard_normal((nsteps,) + shape))
            third = (0.03 * var.lap() if kind == "semi-implicit"
                     else rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
            states = []
            for k in range(nsteps):
                dvar[:] = rhs[k]
                integ.integrate(var, dvar, third)
                if k == 2:
                    integ.override_dt(0.9e-2)   # variable dt, no re-derivation (F5)
                states.append(var.gets().copy())
            tag = f"o{order}_{'si' if kind == 'semi-implicit' else 'ex'}"
            out[f"q0_{tag}"] = q0
            out[f"rhs_{tag}"] = rhs
            out[f"third_{tag}"] = np.asarray(third)
            out[f"states_{tag}"] = np.array(states)
    save(f"integrator_{nx}x{nz}.npz", **out)


# ---------------------------------------------------------------------------
# whole-loop vectors (loop bodies cited from examples/)
# ---------------------------------------------------------------------------
def _sim(params_dict):
    p = Parameters(params_dict)
    sim = Simulation(p, np)
    return p, sim


def _common_setup(sim, p, basis, names, dnames):
    vs = [sim.make_variable(n, basis) for n in names]
    ds = [sim.make_derivative(n) for n in dnames]
    psi = sim.make_variable("psi", basis)
    ux = sim.make_variable("ux", basis)
    uz = sim.make_variable("uz", basis)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        sim.init_laplacian_solver(basis)
    sim.config_cfl(ux, uz)
    return vs, ds, psi, ux, uz


def loop_single_scalar(tag, nx, nz, lx, lz, coef_name, coef, dt, nsteps, ic,
                       snaps, order=2, int_order=2, store_final=True):
    """examples/taylor_green_vortex.py:85-95 /
    examples/kelvin_helmholtz_instability.py:115-131"""
    d = base_params(nx, nz, lx, lz, initial_dt=dt, tracker_cadence=1,
                    spatial_derivative_order=order, integrator_order=int_order,
                    integrator="semi-implicit", cfl_cutoff=0.5,
                    save_cadence=1e9, dump_cadence=1e9)
    d[coef_name] = coef
    d["final_time"] = 1e9
    p, sim = _sim(d)
    (w,), (dw,), psi, ux, uz = _common_setup(sim, p, [CE, CE], ["w"], ["dw"])
    sim.config_scalar_trackers(
        {"ke": partial(calc_kinetic_energy, ux, uz, np, p)})
    w0 = ic(p)
    w.load(w0, is_physical=True)
    out = {"w0_phys": w0} if w0.size <= 64 * 64 else {}
    snapshots = {}
    dts = []
    while sim._loop_counter < nsteps:
        calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
        lin_op = coef * w.lap()
        dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
        sim._integrator.integrate(w, dw, lin_op)
        sim.end_loop()
        dts.append(sim._integrator._dt)
        if sim._loop_counter in snaps:
            snapshots[f"w_step{sim._loop_counter}"] = w[:].copy()
    out.update(snapshots)
    out["ke_t"] = np.array(sim._trackers[0]._times)
    out["ke"] = np.array(sim._trackers[0]._values)
    out["dt_series"] = np.array(dts)
    if store_final:
        out["w_final"] = w[:].copy()
    save(f"loop_{tag}.npz", nx=nx, nz=nz, lx=lx, lz=lz, coef=coef, dt=dt,
         nsteps=nsteps, order=order, int_order=int_order, **out)


def ic_tg(p):
    x = np.linspace(0, p.lx, p.nx, endpoint=False)
    z = np.linspace(0, p.lz, p.nz, endpoint=False)
    X, Z = np.meshgrid(x, z, indexing="ij")
    return -2 * np.cos(X) * np.cos(Z)


def ic_kh(p):
    x = np.linspace(0, p.lx, p.nx, endpoint=False)
    z = np.linspace(0, p.lz, p.nz, endpoint=False)
    X, Z = np.meshgrid(x, z, indexing="ij")
    R = np.sqrt((X - (p.lx / 2)) ** 2 + (Z - 0.5) ** 2)
    rng = np.random.default_rng(0)
    w0_p = np.power(sech((R - 0.25) / 0.1), 2) / 0.1
    w0_p += 0.01 * (2 * rng.random((p.nx, p.nz)) - 1.0)
    return w0_p


def loop_ddc(nx, nz, nsteps, snaps):
    """examples/double_diffusive_convection.py:100-126"""
    LX = 335.0 * 0.25
    d = base_params(nx, nz, LX, 9.0 / 16 * LX, initial_dt=1e-3, Pr=7.0, R0=1.1,
                    tau=1.0 / 3.0, tracker_cadence=1, final_time=1e9,
                    save_cadence=1e9, dump_cadence=1e9,
                    spatial_derivative_order=2, integrator_order=2,
                    integrator="semi-implicit")
    p, sim = _sim(d)
    (w, tmp, xi), (dw, dtmp, dxi), psi, ux, uz = _common_setup(
        sim, p, [CE, CE], ["w", "tmp", "xi"], ["dw", "dtmp", "dxi"])

    def nusselt():
        return 1.0 - np.mean(tmp.getp() * uz.getp())

    sim.config_scalar_trackers(
        {"ke": partial(calc_kinetic_energy, ux, uz, np, p), "nu": nusselt})
    for v in (w, tmp, xi):
        init_var_with_noise(v, 1e-2)
    out = {}
    while sim._loop_counter < nsteps:
        calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
        lin_op = p.Pr * w.lap()
        dw[:] = (-w.vec_dot_nabla(ux.getp(), uz.getp())
                 + p.Pr * xi.sddx() - p.Pr * tmp.sddx())
        sim._integrator.integrate(w, dw, lin_op)
        lin_op = tmp.lap()
        dtmp[:] = -tmp.vec_dot_nabla(ux.getp(), uz.getp()) - uz[:]
        sim._integrator.integrate(tmp, dtmp, lin_op)
        lin_op = p.tau * xi.lap()
        dxi[:] = -xi.vec_dot_nabla(ux.getp(), uz.getp()) - uz[:] / p.R0
        sim._integrator.integrate(xi, dxi, lin_op)
        tmp[:, 0] = 0.0
        xi[:, 0] = 0.0
        sim.end_loop()
        if sim._loop_counter in snaps:
            k = sim._loop_counter
            out[f"w_step{k}"] = w[:].copy()
            out[f"tmp_step{k}"] = tmp[:].copy()
            out[f"xi_step{k}"] = xi[:].copy()
    out["ke"] = np.array(sim._trackers[0]._values)
    out["nu"] = np.array(sim._trackers[1]._values)
    out["t"] = np.array(sim._trackers[0]._times)
    save(f"loop_ddc_{nx}x{nz}.npz", nx=nx, nz=nz, lx=d["lx"], lz=d["lz"],
         dt=1e-3, Pr=7.0, R0=1.1, tau=1.0 / 3.0, nsteps=nsteps, **out)


def loop_tearing(nx, nz, nsteps, snaps):
    """examples/resistive_tearing_instability.py:125-148"""
    d = base_params(nx, nz, 16.0 / 9, 1.0, initial_dt=1e-4, Re=1e6, S=1e6,
                    tracker_cadence=1, final_time=1e9, save_cadence=1e9,
                    dump_cadence=1e9, spatial_derivative_order=2,
                    integrator_order=2, integrator="semi-implicit")
    p, sim = _sim(d)
    (w, j), (dw, dj), psi, ux, uz = _common_setup(
        sim, p, [CE, CE], ["w", "j"], ["dw", "dj"])
    phi = sim.make_variable("phi", [CE, CE])
    bx = sim.make_variable("bx", [CE, CE])
    bz = sim.make_variable("bz", [CE, CE])
    sim.config_scalar_trackers(
        {"ke": partial(calc_kinetic_energy, ux, uz, np, p)})
    x = np.linspace(0, p.lx, p.nx, endpoint=False)
    z = np.linspace(0, p.lz, p.nz, endpoint=False)
    X, Z = np.meshgrid(x, z, indexing="ij")
    rng = np.random.default_rng(0)
    # a wider sheet than the example's 0.01 so a 64-point grid resolves it
    width = 0.01 if nz >= 1024 else 0.1
    j0 = -np.power(sech((Z - 0.5) / width), 2) / width
    j0 += 0.01 * (2 * rng.random((p.nx, p.nz)) - 1.0)
    j.load(j0, is_physical=True)
    out = {"j0_phys": j0}
    solver = sim.get_laplacian_solver()
    while sim._loop_counter < nsteps:
        calc_velocity_from_vorticity(w, psi, ux, uz, solver)
        calc_velocity_from_vorticity(j, phi, bx, bz, solver)
        lin_op = 1.0 / p.Re * w.lap()
        dw[:] = (-w.vec_dot_nabla(ux.getp(), uz.getp())
                 + j.vec_dot_nabla(bx.getp(), bz.getp()))
        sim._integrator.integrate(w, dw, lin_op)
        lin_op = 1.0 / p.S * j.lap()
        dj[:] = (-j.vec_dot_nabla(ux.getp(), uz.getp())
                 + w.vec_dot_nabla(bx.getp(), bz.getp()))
        sim._integrator.integrate(j, dj, lin_op)
        sim.end_loop()
        if sim._loop_counter in snaps:
            k = sim._loop_counter
            out[f"w_step{k}"] = w[:].copy()
            out[f"j_step{k}"] = j[:].copy()
    out["ke"] = np.array(sim._trackers[0]._values)
    save(f"loop_tearing_{nx}x{nz}.npz", nx=nx, nz=nz, lx=d["lx"], lz=d["lz"],
         dt=1e-4, Re=1e6, S=1e6, width=width, nsteps=nsteps, **out)


def loop_rbc(nx, nz, order, int_order, nsteps, snaps):
    """examples/rayleigh_benard_convection.py:95-145"""
    d = base_params(nx, nz, 2.44, 1.0, initial_dt=1e-6, Pr=0.5, Ra=1e6,
                    tracker_cadence=1, final_time=1e9, save_cadence=1e9,
                    dump_cadence=1e9, spatial_derivative_order=order,
                    integrator_order=int_order, integrator="explicit",
                    discretisation=["spectral", "fdm"])
    p, sim = _sim(d)
    basis = [CE, FDM]
    (w, tmp), (dw, dtmp), psi, ux, uz = _common_setup(
        sim, p, basis, ["w", "tmp"], ["dw", "dtmp"])
    sim.config_scalar_trackers(
        {"ke": partial(calc_kinetic_energy, ux, uz, np, p)})
    x = np.linspace(0, p.lx, p.nx, endpoint=False)
    z = np.linspace(0, p.lz, p.nz, endpoint=False)
    X, Z = np.meshgrid(x, z, indexing="ij")
    tmp.load(1 - Z + 1e-2 * (np.sin(np.pi * X / 2.44)), is_physical=True)
    init_var_with_noise(w, 1e-2)
    out = {}
    while sim._loop_counter < nsteps:
        calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
        diffusion_term = p.Pr * w.snabla2()
        dw[:] = (-w.vec_dot_nabla(ux.getp(), uz.getp())
                 - p.Pr * p.Ra * tmp.sddx())
        sim._integrator.integrate(w, dw, diffusion_term)
        diffusion_term = tmp.snabla2()
        dtmp[:] = -tmp.vec_dot_nabla(ux.getp(), uz.getp())
        sim._integrator.integrate(tmp, dtmp, diffusion_term)
        k = 1 if order == 2 else 2
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
        sim.end_loop()
        if sim._loop_counter in snaps:
            s = sim._loop_counter
            out[f"w_step{s}"] = w[:].copy()
            out[f"tmp_step{s}"] = tmp[:].copy()
            out[f"psi_step{s}"] = psi[:].copy()
    out["ke"] = np.array(sim._trackers[0]._values)
    save(f"loop_rbc_{nx}x{nz}_o{order}_ab{int_order}.npz", nx=nx, nz=nz,
         lx=2.44, lz=1.0, dt=1e-6, Pr=0.5, Ra=1e6, order=order,
         int_order=int_order, nsteps=nsteps, **out)


def ops_trig_bases(nx, nz, lx, lz):
    """COSINE / SINE bases (SpectralTransformer.py:90-199, SpatialDifferentiator.py:50-74):
    every pair of spectral bases except the all-Fourier one, on random data."""
    out = {}
    rng = np.random.default_rng(4321)
    p = Parameters(base_params(nx, nz, lx, lz), validate=False)
    af = ArrayFactory(p, np)
    st = SpectralTransformer(p, np, af)
    sd = SpatialDifferentiator(p, np, af)
    phys = rng.standard_normal((nx, nz))
    spec_in = (rng.standard_normal(p.spectral_shape) + 1j * rng.standard_normal(p.spectral_shape))
    out["phys_in"], out["spec_in"] = phys, spec_in
    for bx in (BasisFunctions.COMPLEX_EXP, BasisFunctions.COSINE, BasisFunctions.SINE):
        for bz in (BasisFunctions.COMPLEX_EXP, BasisFunctions.COSINE, BasisFunctions.SINE):
            if bx is CE and bz is CE:
                continue
            tag = f"b{int(bx)}{int(bz)}"
            spec = st.to_spectral(phys.copy(), basis_functions=[bx, bz])
            out[f"{tag}_to_spectral"] = spec.copy()
            # to_physical doubles the mean mode of a cosine axis inside its argument (:123-126):
            # hand it copies, store what it returns
            out[f"{tag}_roundtrip"] = st.to_physical(spec.copy(), basis_functions=[bx, bz]).copy()
            out[f"{tag}_to_physical"] = st.to_physical(spec_in.copy(), basis_functions=[bx, bz]).copy()
            out[f"{tag}_lap"] = np.asarray(sd.calc_lap([bx, bz]), dtype=np.float64)
    for b in (BasisFunctions.COSINE, BasisFunctions.SINE):
        out[f"sddx_b{int(b)}"] = sd.sddx(spec_in, b)
        out[f"sddz_b{int(b)}"] = sd.sddz(spec_in, b)
        out[f"sd2dx2_b{int(b)}"] = sd.sd2dx2(spec_in, b)
        out[f"sd2dz2_b{int(b)}"] = sd.sd2dz2(spec_in, b)
    save(f"ops_trig_{nx}x{nz}.npz", lx=lx, lz=lz, **out)


def main():
    if "--only-trig" in sys.argv:
        ops_trig_bases(64, 32, 1.5, 1.0)
        return
    with tempfile.TemporaryDirectory() as scratch:
        os.chdir(scratch)   # params.json / tracker files land here
        ops_trig_bases(64, 32, 1.5, 1.0)
        ops_fully_spectral(64, 32, 1.5, 1.0)
        ops_fdm(64, 32, 2.44, 1.0)
        integrator_vectors(32, 32)
        loop_single_scalar("tg_64x64", 64, 64, 2 * np.pi, 2 * np.pi, "nu", 0.25,
                           1e-3, 20, ic_tg, snaps=(1, 2, 10, 20))
        loop_single_scalar("tg_64x64_o4_ab4", 64, 64, 2 * np.pi, 2 * np.pi,
                           "nu", 0.25, 1e-3, 12, ic_tg, snaps=(1, 5, 12),
                           order=4, int_order=4)
        # config 1 exactly: 256^2, 1000 steps, KE sampled every step
        loop_single_scalar("tg_256x256_1000", 256, 256, 2 * np.pi, 2 * np.pi,
                           "nu", 0.25, 1e-3, 1000, ic_tg, snaps=(1, 100, 1000),
                           store_final=False)
        loop_single_scalar("kh_128x64", 128, 64, 16.0 / 9.0, 1.0, "invRe", 1e-5,
                           0.05 * (16.0 / 9.0) / 128, 30, ic_kh,
                           snaps=(1, 10, 30))
        loop_ddc(64, 64, 20, snaps=(1, 10, 20))
        loop_tearing(64, 64, 20, snaps=(1, 10, 20))
        loop_rbc(64, 32, 2, 2, 20, snaps=(1, 10, 20))
        loop_rbc(64, 32, 4, 4, 20, snaps=(1, 10, 20))


if __name__ == "__main__":
    main()
