"""Orderings the reference allows because it evaluates eagerly, and the deferred-evaluation
layer must therefore reproduce: a right-hand side assigned *before* one of
its sources is integrated or written in place, writes through views / augmented assignment,
and array handles held across steps.  Each case runs the public `melvin` API next to the
oracle's eager restatement of the same statements; shared by the emulation (CPU) and the
GPU test modules."""
import numpy as np

import parity_cases as pc
from melvin import b200 as xp
from melvin.utility import calc_velocity_from_vorticity
from oracle import melvin_oracle as mo

CE = pc.CE


def _setup(nx, nz, dt, names, dnames):
    d = pc.base_params(nx, nz, 2 * np.pi, 2 * np.pi, initial_dt=dt, spatial_derivative_order=2,
                       integrator_order=2, integrator="semi-implicit", cfl_cutoff=0.5)
    return pc.make_sim(d, names, dnames, [CE, CE])


def _ics(g):
    rng = np.random.default_rng(7)
    w0 = mo.ic_taylor_green(g) + 0.1 * rng.standard_normal(g.physical_shape)
    t0 = rng.standard_normal(g.physical_shape)
    return w0, t0


def swapped_integrate_order(nx=32, nz=32, nsteps=5, inplace_write=True):
    """dw's right-hand side reads tmp; tmp is integrated (and fixed up in place) first."""
    dt, c, nu, kappa = 1e-3, 0.7, 0.05, 0.02
    g = mo.Grid(nx, nz, 2 * np.pi, 2 * np.pi)
    w0, t0 = _ics(g)
    with pc.scratch_cwd():
        p, sim, (w, tmp), (dw, dtmp), psi, ux, uz = _setup(nx, nz, dt, ["w", "tmp"], ["dw", "dtmp"])
        w.load(w0, is_physical=True)
        tmp.load(t0, is_physical=True)
        held = tmp.sddx()                      # a user-held deferred expression of the old tmp
        held_want = mo.sddx(g, mo.to_spectral(g, t0))
        for _ in range(nsteps):
            calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
            dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp()) + c * tmp.sddx()
            dtmp[:] = -tmp.vec_dot_nabla(ux.getp(), uz.getp()) - uz[:]
            sim._integrator.integrate(tmp, dtmp, kappa * tmp.lap())
            if inplace_write:
                tmp[:, 0] = 0.0
            sim._integrator.integrate(w, dw, nu * w.lap())
            sim.end_loop()
        got = {"w": pc.host(w[:]), "tmp": pc.host(tmp[:]), "held": pc.host(held[:])}
    # oracle: the same statements, evaluated eagerly
    ws, ts = mo.to_spectral(g, w0), mo.to_spectral(g, t0)
    hw, ht = mo.History(g), mo.History(g)
    lap = mo.lap_symbol(g)
    for _ in range(nsteps):
        vel = mo.velocity_from_vorticity(g, ws)
        hw.set_current(-mo.vec_dot_nabla(g, ws, vel["ux_p"], vel["uz_p"])[0] + c * mo.sddx(g, ts))
        ht.set_current(-mo.vec_dot_nabla(g, ts, vel["ux_p"], vel["uz_p"])[0] - vel["uz_s"])
        ts = mo.integrate_semi_implicit(g, ts, ht, kappa * lap, dt)
        if inplace_write:
            ts[:, 0] = 0.0
        ws = mo.integrate_semi_implicit(g, ws, hw, nu * lap, dt)
    return got, {"w": ws, "tmp": ts, "held": held_want}


def augmented_assignment_after_velocity(nx=32, nz=32):
    """`w[:] += kick` after calc_velocity_from_vorticity: ux/uz/psi keep the un-kicked field."""
    dt, nu = 1e-3, 0.05
    g = mo.Grid(nx, nz, 2 * np.pi, 2 * np.pi)
    w0, f0 = _ics(g)
    with pc.scratch_cwd():
        p, sim, (w, f), (dw, _), psi, ux, uz = _setup(nx, nz, dt, ["w", "f"], ["dw", "df"])
        w.load(w0, is_physical=True)
        f.load(f0, is_physical=True)
        calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
        w[:] += 0.05 * f[:]
        dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
        got = {"ux_p": pc.host(ux.getp()), "psi": pc.host(psi[:])}
        sim._integrator.integrate(w, dw, nu * w.lap())
        got["w"] = pc.host(w[:])
        # a view taken from gets() writes through, with the same flush
        calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
        v = w.gets()[:, 1:3]
        v[...] = 0.0
        got["uz_s_after_view_write"] = pc.host(uz[:])
        got["w_after_view_write"] = pc.host(w[:])
    ws, fs = mo.to_spectral(g, w0), mo.to_spectral(g, f0)
    vel = mo.velocity_from_vorticity(g, ws)
    ws = ws + 0.05 * fs
    h = mo.History(g)
    h.set_current(-mo.vec_dot_nabla(g, ws, vel["ux_p"], vel["uz_p"])[0])
    w1 = mo.integrate_semi_implicit(g, ws, h, nu * mo.lap_symbol(g), dt)
    vel1 = mo.velocity_from_vorticity(g, w1)
    w2 = w1.copy()
    w2[:, 1:3] = 0.0
    want = {"ux_p": vel["ux_p"], "psi": vel["psi_s"], "w": w1, "uz_s_after_view_write": vel1["uz_s"],
            "w_after_view_write": w2}
    return got, want


def held_handle_stays_current(nx=32, nz=32, nsteps=3):
    """`ws = w.gets()` taken outside the loop is the live state on every step (the reference
    returns the same ndarray each time), and writes through it reach the Variable."""
    dt, nu = 1e-3, 0.05
    g = mo.Grid(nx, nz, 2 * np.pi, 2 * np.pi)
    w0, _ = _ics(g)
    with pc.scratch_cwd():
        p, sim, (w,), (dw,), psi, ux, uz = _setup(nx, nz, dt, ["w"], ["dw"])
        w.load(w0, is_physical=True)
        ws = w.gets()
        row = w.gets()[0]
        same, rows_ok = [], []
        for _ in range(nsteps):
            calc_velocity_from_vorticity(w, psi, ux, uz, sim.get_laplacian_solver())
            dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
            sim._integrator.integrate(w, dw, nu * w.lap())
            sim.end_loop()
            same.append(ws is w.gets() and np.array_equal(pc.host(ws), pc.host(w[:])))
            rows_ok.append(np.array_equal(pc.host(row), pc.host(w[:])[0]))
        ws[:, 0] = 0.0
        col0 = pc.host(w[:])[:, 0]
    return same, rows_ok, col0
