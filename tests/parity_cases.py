"""Drivers that run the example scripts' loop bodies through the drop-in `melvin`
package (the public API a user calls).  Shared by the GPU parity tests and by the
CPU emulation tests; the loop bodies are those of the reference's examples
(cited per function) so the tests read like the reference's own scripts."""
import contextlib
import io
import os
import tempfile
from functools import partial

import numpy as np

import melvin
from melvin import BasisFunctions, Parameters, Simulation
from melvin import b200 as xp
from melvin.utility import calc_kinetic_energy, calc_velocity_from_vorticity, init_var_with_noise

CE = BasisFunctions.COMPLEX_EXP
FDM = BasisFunctions.FDM


@contextlib.contextmanager
def scratch_cwd():
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


def base_params(nx, nz, lx, lz, **kw):
    d = {"nx": nx, "nz": nz, "lx": lx, "lz": lz, "final_time": 1e9, "precision": "double",
         "save_cadence": 1e9, "dump_cadence": 1e9, "tracker_cadence": 1}
    d.update(kw)
    return d


def make_sim(pdict, names, dnames, basis):
    p = Parameters(pdict)
    sim = Simulation(p, xp)
    vs = [sim.make_variable(n, basis) for n in names]
    ds = [sim.make_derivative(n) for n in dnames]
    psi = sim.make_variable("psi", basis)
    ux = sim.make_variable("ux", basis)
    uz = sim.make_variable("uz", basis)
    with contextlib.redirect_stdout(io.StringIO()):
        sim.init_laplacian_solver(basis)
    sim.config_cfl(ux, uz)
    return p, sim, vs, ds, psi, ux, uz


def host(a):
    return np.asarray(a.get() if hasattr(a, "get") else a)


def run_single_scalar(nx, nz, lx, lz, coef, dt, nsteps, w0_phys, snaps=(), order=2, int_order=2,
                      tracker_cadence=1, strict_reads=False):
    """examples/taylor_green_vortex.py:85-95, examples/kelvin_helmholtz_instability.py:115-131"""
    d = base_params(nx, nz, lx, lz, initial_dt=dt, spatial_derivative_order=order,
                    integrator_order=int_order, integrator="semi-implicit", cfl_cutoff=0.5,
                    tracker_cadence=tracker_cadence)
    p, sim, (w,), (dw,), psi, ux, uz = make_sim(d, ["w"], ["dw"], [CE, CE])
    sim.config_scalar_trackers({"ke": partial(calc_kinetic_energy, ux, uz, xp, p)})
    w.load(w0_phys, is_physical=True)
    out = {}
    while sim._loop_counter < nsteps:
        calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
        lin_op = coef * w.lap()
        dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
        sim._integrator.integrate(w, dw, lin_op)
        if strict_reads and sim._loop_counter == 0:
            # reads after the update must still see the pre-update fields
            out["psi_after_step1"] = host(psi[:])
            out["ux_p_after_step1"] = host(ux.getp())
        sim.end_loop()
        if sim._loop_counter in snaps:
            out[f"w_step{sim._loop_counter}"] = host(w[:])
    out["w_final"] = host(w[:])
    out["ke"] = np.array([float(v) for v in sim._trackers[0]._values])
    out["ke_t"] = np.array(sim._trackers[0]._times)
    out["dt"] = sim._integrator._dt
    return out


def run_ddc(nx, nz, lx, lz, dt, nsteps, Pr, R0, tau, snaps=()):
    """examples/double_diffusive_convection.py:100-126"""
    d = base_params(nx, nz, lx, lz, initial_dt=dt, Pr=Pr, R0=R0, tau=tau,
                    spatial_derivative_order=2, integrator_order=2, integrator="semi-implicit")
    p, sim, (w, tmp, xi), (dw, dtmp, dxi), psi, ux, uz = make_sim(
        d, ["w", "tmp", "xi"], ["dw", "dtmp", "dxi"], [CE, CE])

    def nusselt():
        return 1.0 - xp.mean(tmp.getp() * uz.getp())

    sim.config_scalar_trackers({"ke": partial(calc_kinetic_energy, ux, uz, xp, p), "nu": nusselt})
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
        k = sim._loop_counter
        if k in snaps:
            out[f"w_step{k}"], out[f"tmp_step{k}"], out[f"xi_step{k}"] = host(w[:]), host(tmp[:]), host(xi[:])
    out["ke"] = np.array([float(v) for v in sim._trackers[0]._values])
    out["nu"] = np.array([float(v) for v in sim._trackers[1]._values])
    return out


def run_tearing(nx, nz, lx, lz, dt, nsteps, Re, S, j0_phys, snaps=(), w0_phys=None):
    """examples/resistive_tearing_instability.py:125-148 (w0_phys: optional non-zero initial vorticity)"""
    d = base_params(nx, nz, lx, lz, initial_dt=dt, Re=Re, S=S, spatial_derivative_order=2,
                    integrator_order=2, integrator="semi-implicit")
    p, sim, (w, j), (dw, dj), psi, ux, uz = make_sim(d, ["w", "j"], ["dw", "dj"], [CE, CE])
    phi = sim.make_variable("phi", [CE, CE])
    bx = sim.make_variable("bx", [CE, CE])
    bz = sim.make_variable("bz", [CE, CE])
    sim.config_scalar_trackers({"ke": partial(calc_kinetic_energy, ux, uz, xp, p)})
    j.load(j0_phys, is_physical=True)
    if w0_phys is not None:
        w.load(w0_phys, is_physical=True)
    solver = sim.get_laplacian_solver()
    out = {}
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
        k = sim._loop_counter
        if k in snaps:
            out[f"w_step{k}"], out[f"j_step{k}"] = host(w[:]), host(j[:])
    out["ke"] = np.array([float(v) for v in sim._trackers[0]._values])
    return out


def run_rbc(nx, nz, order, int_order, dt, nsteps, Pr, Ra, snaps=()):
    """examples/rayleigh_benard_convection.py:95-145"""
    d = base_params(nx, nz, 2.44, 1.0, initial_dt=dt, Pr=Pr, Ra=Ra,
                    spatial_derivative_order=order, integrator_order=int_order,
                    integrator="explicit", discretisation=["spectral", "fdm"])
    basis = [CE, FDM]
    p, sim, (w, tmp), (dw, dtmp), psi, ux, uz = make_sim(d, ["w", "tmp"], ["dw", "dtmp"], basis)
    sim.config_scalar_trackers({"ke": partial(calc_kinetic_energy, ux, uz, xp, p)})
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
        if p.spatial_derivative_order == 2:
            w[1:, 0] = 0.0
            w[1:, -1] = 0.0
            psi[1:, 0] = 0.0
            psi[1:, -1] = 0.0
            tmp[0, 0] = 1.0
            tmp[0, -1] = 0.0
            tmp[1:, 0] = 0.0
            tmp[1:, -1] = 0.0
        else:
            w[1:, :2] = 0.0
            w[1:, -2:] = 0.0
            psi[1:, :2] = 0.0
            psi[1:, -2:] = 0.0
            tmp[0, :2] = 1.0
            tmp[0, -2:] = 0.0
            tmp[1:, :2] = 0.0
            tmp[1:, -2:] = 0.0
        psi[0, :] = 0.0
        w[0, :] = 0.0
        sim.end_loop()
        k = sim._loop_counter
        if k in snaps:
            out[f"w_step{k}"], out[f"tmp_step{k}"], out[f"psi_step{k}"] = host(w[:]), host(tmp[:]), host(psi[:])
    out["ke"] = np.array([float(v) for v in sim._trackers[0]._values])
    return out
