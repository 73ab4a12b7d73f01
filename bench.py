#!/usr/bin/env python3
"""Benchmark of the per-timestep pseudo-spectral hot path (BASELINE.json metric:
grid-point-timesteps/s, fp64).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config kh|rbc|ddc|tearing]

Headline workload: BASELINE configs[1] -- Kelvin-Helmholtz, fully spectral, 4096x4096, AB2 +
semi-implicit diffusion; the loop body is the example script's
(reference examples/kelvin_helmholtz_instability.py:115-131).
  N = 1 : the loop through the drop-in `melvin` API (which calls libmelvin_b200.so);
  N > 1 : the same grid slab-decomposed over the GPUs (melvin/sharded.py) -- strong scaling;
          a parity preflight against the committed goldens of the unmodified reference runs
          first and the process exits non-zero if it fails.
Every line also carries a `large_grid` block: BASELINE configs[4] (resistive tearing, 16384^2)
and configs[3] (double-diffusive convection, 8192^2) on the same N GPUs.
`--impl reference`: the UNMODIFIED reference (pip-installed into baseline/_ref by
tools/install_reference.sh) on the host cores, NumPy float64, same loop and initial condition.
One "step" = one time step of the whole field.  Prints ONE JSON line ("Measurement").
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "melvin.py_b200")
REFDIR = os.path.join(ROOT, "baseline", "_ref")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "grid-point-timesteps/sec (fp64)"
UNIT = "grid-point-steps/s"
NVLINK_GBPS = 770.0            # measured peer copy, per direction and GPU (B200_PROFILING.md)

# ------------------------------------------------------------------ workloads
CONFIGS = {
    # name: default grid, BASELINE.json entry, label (identical in both arms: the driver compares it)
    "kh": ((4096, 4096), "Kelvin-Helmholtz {nx}x{nz} fully spectral, AB2 + semi-implicit diffusion "
                         "(BASELINE configs[1]; loop of examples/kelvin_helmholtz_instability.py:115-131)"),
    "rbc": ((4096, 2048), "Rayleigh-Benard {nx}x{nz} Fourier-x / 4th-order finite-difference z, AB4 explicit "
                          "(BASELINE configs[2]; loop of examples/rayleigh_benard_convection.py:95-145)"),
    "ddc": ((8192, 8192), "double-diffusive convection {nx}x{nz} fully spectral, 3 fields, AB2 + semi-implicit "
                          "(BASELINE configs[3]; loop of examples/double_diffusive_convection.py:100-126)"),
    "tearing": ((16384, 16384), "resistive tearing (MHD) {nx}x{nz} fully spectral, AB2 + semi-implicit "
                                "(BASELINE configs[4]; loop of examples/resistive_tearing_instability.py:125-148)"),
}


def workload(config, nx, nz):
    return CONFIGS[config][1].format(nx=nx, nz=nz)


def run_params(config, nx, nz):
    """Parameter dict of the example script at the BASELINE grid."""
    base = {"nx": nx, "nz": nz, "final_time": 1e9, "save_cadence": 1e9, "dump_cadence": 1e9,
            "precision": "double", "cfl_cutoff": 0.5}
    if config == "kh":
        lx = 16.0 / 9.0
        base.update(lx=lx, lz=1.0, Re=1e5, spatial_derivative_order=2, integrator_order=2,
                    integrator="semi-implicit", initial_dt=0.05 * lx / nx)
    elif config == "rbc":
        base.update(lx=2.44, lz=1.0, Pr=0.5, Ra=1e6, spatial_derivative_order=4, integrator_order=4,
                    integrator="explicit", discretisation=["spectral", "fdm"],
                    # the example's dt = 1e-6 (nz = 13) is 40x beyond the stability limit of explicit
                    # AB4 diffusion at nz = 2048 (the reference blows up in ~10 steps as well)
                    initial_dt=min(1e-6, 0.05 / (nz * nz)))
    elif config == "ddc":
        lx = 83.75
        base.update(lx=lx, lz=9.0 * lx / 16.0, Pr=7.0, R0=1.1, tau=1.0 / 3.0, spatial_derivative_order=2,
                    integrator_order=2, integrator="semi-implicit", initial_dt=1e-3)
    elif config == "tearing":
        lx = 16.0 / 9.0
        base.update(lx=lx, lz=1.0, Re=1e6, S=1e6, spatial_derivative_order=2, integrator_order=2,
                    integrator="semi-implicit", initial_dt=0.01 * 0.05 * lx / nx)
    return base


def byte_model(config, nx, nz):
    """Algorithmic bytes per step of SURVEY section 8(d): 2*I*(N_inv + N_fwd) + S*N_state."""
    nn, nm = (nx - 1) // 3, (nz - 1) // 3
    S = 16 * (2 * nn + 1) * nm
    I = 16 * nx * nm
    if config == "rbc":
        Sf = 16 * nn * nz
        return {"S_f": Sf, "step": 23 * Sf}
    n_inv, n_fwd, n_state = {"kh": (3, 1, 5), "ddc": (5, 3, 15), "tearing": (7, 4, 10)}[config]
    bm = {"S": S, "I": I, "step": 2 * I * (n_inv + n_fwd) + S * n_state,
          "transposes": n_inv + n_fwd}
    if config == "kh":
        bm.update({"mlv_x_inverse": S + 3 * I,       # read w-hat, write the 3 x-transformed fields
                   "mlv_advect_z": 4 * I,            # read 3 I, write 1 I (the model's single forward field)
                   "mlv_x_forward": I + 4 * S})      # read 1 I, w-hat, f(-1); write w-hat, f(0)
    return bm


def csrc_digest():
    h = hashlib.sha1()
    d = os.path.join(PKG, "csrc")
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), "rb") as fp:
            h.update(fp.read())
    return h.hexdigest()[:16]


def measured_traffic(nx, nz):
    """DRAM bytes per launch from the newest committed ncu capture -- only if it was taken on
    exactly these kernel sources (digest recorded next to it) and this grid; else None."""
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True):
        if not (name.endswith("_traffic.json")):
            continue
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fp:
                d = json.load(fp)
            if list(d["grid"]) == [nx, nz] and d.get("csrc_digest") == csrc_digest():
                return d["bytes_per_launch"], d["source"]
        except (OSError, ValueError, KeyError):
            continue
    return {}, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:     # wait for the first sample
                time.sleep(0.02)
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self, t_begin=None, t_end=None):
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = [ln for ts, ln in self.lines
                  if (t_begin is None or ts >= t_begin) and (t_end is None or ts <= t_end + 0.06)]
        if len(window) < 2:          # very short timed region: use every sample taken under load
            window = [ln for _, ln in self.lines]
        for line in window:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smmax.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smmax)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------- initial conditions (host, as the scripts)
def initial_fields(config, nx, nz, pd):
    """Physical-space initial conditions of the example script (oracle restatements, host)."""
    from oracle import melvin_oracle as mo
    g = mo.Grid(nx, nz, pd["lx"], pd["lz"], fdm_z=(config == "rbc"))
    if config == "kh":
        return {"w": mo.ic_kelvin_helmholtz(g)}
    if config == "tearing":
        return {"j": mo.ic_tearing_current(g)}
    if config == "ddc":
        return {"noise": mo.ic_noise(g)}                # w, tmp, xi: same seed 0 (utility.py:31-39)
    return {"w": mo.ic_noise(g), "tmp": mo.ic_rbc_temperature(g)}


# ------------------------------------------------------------- CPU arms
def reference_loop(nx, nz, nwarm, nsteps):
    """Kelvin-Helmholtz loop on the UNMODIFIED reference (baseline/_ref), xp = numpy, float64.
    Returns (gp-steps/s, s/step, kind)."""
    pd = run_params("kh", nx, nz)
    w0 = initial_fields("kh", nx, nz, pd)["w"]
    if os.path.isdir(os.path.join(REFDIR, "melvin")):
        sys.path.insert(0, REFDIR)
        from functools import partial
        import melvin
        assert os.path.dirname(os.path.abspath(melvin.__file__)).startswith(REFDIR), "reference arm imported the wrong melvin"
        from melvin import BasisFunctions, Parameters, Simulation
        from melvin.utility import calc_kinetic_energy, calc_velocity_from_vorticity
        os.chdir(tempfile.mkdtemp(prefix="mlvref"))
        params = Parameters(pd)
        sim = Simulation(params, np)
        basis = [BasisFunctions.COMPLEX_EXP, BasisFunctions.COMPLEX_EXP]
        w = sim.make_variable("w", basis)
        dw = sim.make_derivative("dw")
        psi, ux, uz = (sim.make_variable(n, basis) for n in ("psi", "ux", "uz"))
        sim.init_laplacian_solver(basis)
        sim.config_cfl(ux, uz)
        sim.config_scalar_trackers({"kinetic_energy.npz": partial(calc_kinetic_energy, ux, uz, np, params)})
        w.load(w0, is_physical=True)
        solver = sim.get_laplacian_solver()

        def step():        # examples/kelvin_helmholtz_instability.py:115-131
            calc_velocity_from_vorticity(w, psi, ux, uz, solver)
            lin_op = 1.0 / params.Re * w.lap()
            dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
            sim._integrator.integrate(w, dw, lin_op)
            sim.end_loop()
        kind = "reference"
    else:
        from oracle import melvin_oracle as mo
        g = mo.Grid(nx, nz, pd["lx"], pd["lz"])
        run = mo.Run(g, pd["initial_dt"])
        state = {"w": mo.to_spectral(g, w0)}
        hist = mo.History(g)

        def step():
            state["w"] = mo.step_single_scalar(g, run, state["w"], hist, 1.0 / pd["Re"])
        kind = "port"
    for _ in range(nwarm):
        step()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        step()
    dt = time.perf_counter() - t0
    return nx * nz * nsteps / dt, dt / nsteps, kind


def reference_arm(args, rank):
    """`--impl reference`: the reference's own CPU implementation of the path (NumPy float64,
    pocketfft: single-threaded by construction) on the same grid, loop and initial condition.
    The grid is never shrunk; a run that would take too long times fewer steps and says so."""
    if rank != 0:
        return
    nx, nz = args.nx, args.nz
    est = 3.3 * (nx * nz) / 4096.0 ** 2                 # s per step
    budget = 240.0
    warm = max(1, min(args.warmup, 2))
    nsteps = int(max(2, min(args.steps, (budget - warm * est) // est)))
    value, s_per_step, kind = reference_loop(nx, nz, warm, nsteps)
    what = ("the unmodified reference (baseline/_ref, xp = numpy, float64)" if kind == "reference"
            else "the oracle port (oracle/melvin_oracle.py; baseline/_ref is not installed)")
    sample = (f"{nsteps} timed + {warm} warm-up steps of {what} on the full {nx}x{nz} grid"
              + ("" if nsteps == args.steps else f" ({args.steps} requested; capped to keep the run within minutes)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_timed": nsteps,
        "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload("kh", nx, nz), "grid": [nx, nz]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def cpu_baseline_subprocess(nx, nz):
    """The reference arm in a fresh process (this one has the drop-in `melvin` imported)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--nx", str(nx), "--nz", str(nz)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    for ln in reversed(out.stdout.splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError("reference arm failed: " + out.stderr[-2000:])


# ----------------------------------------------------------- public-API loops (N = 1)
def build_public_loop(config, nx, nz):
    """The example script's set-up and loop body on the drop-in package.  Returns
    (step function, dict of the objects the caller may want)."""
    from functools import partial
    from melvin import BasisFunctions, Parameters, Simulation
    from melvin import b200 as xp
    from melvin.utility import calc_kinetic_energy, calc_velocity_from_vorticity
    pd = run_params(config, nx, nz)
    ic = initial_fields(config, nx, nz, pd)
    os.chdir(tempfile.mkdtemp(prefix="mlvbench"))
    params = Parameters(pd)
    sim = Simulation(params, xp)
    CE, FDM = BasisFunctions.COMPLEX_EXP, BasisFunctions.FDM
    basis = [CE, FDM] if config == "rbc" else [CE, CE]
    mk = lambda n: sim.make_variable(n, basis)           # noqa: E731
    psi, ux, uz = mk("psi"), mk("ux"), mk("uz")
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        sim.init_laplacian_solver(basis)
    sim.config_cfl(ux, uz)
    trackers = {"kinetic_energy.npz": partial(calc_kinetic_energy, ux, uz, xp, params)}
    solver = sim.get_laplacian_solver()
    integ = sim._integrator
    o = {"sim": sim, "params": params, "ux": ux, "uz": uz, "psi": psi, "xp": xp}

    if config == "kh":
        w, dw = mk("w"), sim.make_derivative("dw")
        sim.config_scalar_trackers(trackers)
        w.load(ic["w"], is_physical=True)
        o.update(w=w, dw=dw)

        def step():        # examples/kelvin_helmholtz_instability.py:115-131
            calc_velocity_from_vorticity(w, psi, ux, uz, solver)
            lin_op = 1.0 / params.Re * w.lap()
            dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
            integ.integrate(w, dw, lin_op)
            sim.end_loop()
    elif config == "tearing":
        w, j, dw, dj = mk("w"), mk("j"), sim.make_derivative("dw"), sim.make_derivative("dj")
        phi, bx, bz = mk("phi"), mk("bx"), mk("bz")
        sim.config_scalar_trackers(trackers)
        j.load(ic["j"], is_physical=True)
        o.update(w=w, j=j)

        def step():        # examples/resistive_tearing_instability.py:125-148
            calc_velocity_from_vorticity(w, psi, ux, uz, solver)
            calc_velocity_from_vorticity(j, phi, bx, bz, solver)
            lin_op = 1.0 / params.Re * w.lap()
            dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp()) + j.vec_dot_nabla(bx.getp(), bz.getp())
            integ.integrate(w, dw, lin_op)
            lin_op = 1.0 / params.S * j.lap()
            dj[:] = -j.vec_dot_nabla(ux.getp(), uz.getp()) + w.vec_dot_nabla(bx.getp(), bz.getp())
            integ.integrate(j, dj, lin_op)
            sim.end_loop()
    elif config == "ddc":
        w, tmp, xi = mk("w"), mk("tmp"), mk("xi")
        dw, dtmp, dxi = (sim.make_derivative(n) for n in ("dw", "dtmp", "dxi"))
        trackers["nusselt_number.npz"] = lambda: 1.0 - xp.mean(tmp.getp() * uz.getp())
        sim.config_scalar_trackers(trackers)
        for v in (w, tmp, xi):
            v.load(ic["noise"], is_physical=True)
        o.update(w=w, tmp=tmp, xi=xi)

        def step():        # examples/double_diffusive_convection.py:100-126
            calc_velocity_from_vorticity(w, psi, ux, uz, solver)
            lin_op = params.Pr * w.lap()
            dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp()) + params.Pr * xi.sddx() - params.Pr * tmp.sddx()
            integ.integrate(w, dw, lin_op)
            lin_op = tmp.lap()
            dtmp[:] = -tmp.vec_dot_nabla(ux.getp(), uz.getp()) - uz[:]
            integ.integrate(tmp, dtmp, lin_op)
            lin_op = params.tau * xi.lap()
            dxi[:] = -xi.vec_dot_nabla(ux.getp(), uz.getp()) - uz[:] / params.R0
            integ.integrate(xi, dxi, lin_op)
            tmp[:, 0] = 0.0
            xi[:, 0] = 0.0
            sim.end_loop()
    else:                  # rbc
        w, tmp = mk("w"), mk("tmp")
        dw, dtmp = sim.make_derivative("dw"), sim.make_derivative("dtmp")
        sim.config_scalar_trackers(trackers)
        tmp.load(ic["tmp"], is_physical=True)
        w.load(ic["w"], is_physical=True)
        o.update(w=w, tmp=tmp)

        def step():        # examples/rayleigh_benard_convection.py:95-145
            calc_velocity_from_vorticity(w, psi, ux, uz, solver)
            diffusion_term = params.Pr * w.snabla2()
            dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp()) - params.Pr * params.Ra * tmp.sddx()
            integ.integrate(w, dw, diffusion_term)
            diffusion_term = tmp.snabla2()
            dtmp[:] = -tmp.vec_dot_nabla(ux.getp(), uz.getp())
            integ.integrate(tmp, dtmp, diffusion_term)
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
    return step, o


# ------------------------------------------------------------- sharded steppers (N > 1)
def build_sharded(config, nx, nz):
    """Slab-decomposed stepper of the same loop, initial condition transformed once on a
    full-size context (set-up, outside every timed region)."""
    import torch
    from melvin import _backend
    from melvin.sharded import (ShardedDoubleDiffusiveStepper, ShardedScalarStepper,
                                ShardedTearingStepper)
    pd = run_params(config, nx, nz)
    ic = initial_fields(config, nx, nz, pd)
    if config == "kh":
        st = ShardedScalarStepper(nx, nz, pd["lx"], pd["lz"], 1.0 / pd["Re"], pd["initial_dt"])
    elif config == "ddc":
        st = ShardedDoubleDiffusiveStepper(nx, nz, pd["lx"], pd["lz"], pd["Pr"], pd["R0"], pd["tau"], pd["initial_dt"])
    elif config == "tearing":
        st = ShardedTearingStepper(nx, nz, pd["lx"], pd["lz"], pd["Re"], pd["S"], pd["initial_dt"])
    else:
        raise SystemExit(f"bench.py: config {config} has no slab-decomposed form (DESIGN.md, multi-GPU)")
    full = _backend.Context(nx, nz, pd["lx"], pd["lz"], False, 2)
    phys = _backend.from_host(next(iter(ic.values())))
    spec = _backend.empty(full.spec_shape, np.complex128)
    full.call("mlv_to_spectral", ctypes.c_void_p(phys.data_ptr()),
              ctypes.c_void_p(full.scratch_i().data_ptr()), ctypes.c_void_p(spec.data_ptr()))
    zero = _backend.zeros(full.spec_shape, np.complex128) if config == "tearing" else None
    if config == "kh":
        st.load_spectral(spec)
    elif config == "ddc":
        st.load_spectral(spec, spec, spec)
    else:
        st.load_spectral(zero, spec)
    _backend.synchronize()
    del phys, spec, full, zero
    if _backend.is_cuda():
        torch.cuda.empty_cache()
    return st


def time_steps(step, warmup, steps, world, clk=None):
    """`warmup` untimed + exactly `steps` timed steps; barrier + synchronize on both sides; CUDA
    events; max over ranks.  Returns (ms total, library launches inside the timed region)."""
    import torch
    import torch.distributed as dist
    from melvin import _backend

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warmup):
        step()
    barrier()
    n_before = _backend.launches()
    marks = [clk.mark()] if clk else []
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    if clk:
        marks.append(clk.mark())
    launches = _backend.launches() - n_before
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, launches, marks


def large_grid_block(world, rank, peak, only=None):
    """BASELINE configs[4] (tearing 16384^2) and configs[3] (DDC 8192^2) on the same GPUs:
    N = 1 through the public API, N > 1 slab-decomposed.  Strong scaling: the driver's 1 -> 8
    curve of these entries is the one the north star grades."""
    import gc
    import torch
    out = {}
    for config, nsteps in (("tearing", 6), ("ddc", 10)):
        if only and config not in only:
            continue
        (nx, nz), _ = CONFIGS[config]
        bm = byte_model(config, nx, nz)
        entry = {"workload": workload(config, nx, nz), "grid": [nx, nz], "steps": nsteps, "warmup": 3}
        try:
            if world == 1:
                step, keep = build_public_loop(config, nx, nz)
                mode, sent = "none (1 GPU, public melvin API)", 0
                st = None
            else:
                st = build_sharded(config, nx, nz)
                step, keep = st.step, st
                mode, sent = st.mode, int(st.bytes_exchanged_per_step)
            ms, launches, _ = time_steps(step, 3, nsteps, world)
            msps = ms / nsteps
            gbps = bm["step"] / world / (msps * 1e-3) / 1e9
            nv_ms = sent / (NVLINK_GBPS * 1e9) * 1e3
            entry.update({
                "ms_per_step": msps, "value": nx * nz * nsteps / (ms * 1e-3), "unit": UNIT,
                "exchange_mode": mode, "sent_bytes_per_gpu_per_step": sent, "gpu_launches": launches,
                "hbm": {"algorithmic_bytes_per_step": bm["step"], "achieved_GBps_per_gpu": gbps,
                        "frac_of_measured_peak": gbps / peak},
                "nvlink": {"ms_at_770GBps": nv_ms, "frac_of_step": nv_ms / msps},
            })
            if st is not None:
                st.close()
            del step, keep, st
        except Exception as exc:                      # noqa: BLE001  (reported, never fatal for the headline)
            entry["error"] = f"{type(exc).__name__}: {exc}"[:400]
        from melvin import _backend
        _backend._contexts.clear()
        gc.collect()
        torch.cuda.empty_cache()
        out[config] = entry
    return out


# ------------------------------------------------------------------ parity preflights
def parity_single_gpu():
    """Taylor-Green 64^2, 20 steps through the public API against the golden of the unmodified
    reference (tests/golden/loop_tg_64x64.npz)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_cases as pc
    from oracle import melvin_oracle as mo
    gl = np.load(os.path.join(ROOT, "tests", "golden", "loop_tg_64x64.npz"))
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    with pc.scratch_cwd():
        out = pc.run_single_scalar(64, 64, g.lx, g.lz, float(gl["coef"]), float(gl["dt"]), 20,
                                   mo.ic_taylor_green(g), snaps=(1, 2, 10, 20))
    ferr = max(mo.relative_l2(out[f"w_step{k}"], gl[f"w_step{k}"]) for k in (1, 2, 10, 20))
    kerr = float(np.max(np.abs(out["ke"] / gl["ke"] - 1)))
    return {"ok": bool(ferr < 1e-12 and kerr < 1e-9),
            "cases": [{"case": "taylor_green_64x64", "exchange_mode": "none", "steps": 20,
                       "field_rel_l2": float(ferr), "ke_rel": kerr, "ok": bool(ferr < 1e-12 and kerr < 1e-9)}]}


def parity_sharded():
    """The three slab-decomposed steppers against the goldens of the unmodified reference, in
    the exchange modes the timed runs use (p2p at 4096^2, dma at the large grids) plus NCCL."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import sharded_parity as sp
    cases = [sp.check_taylor_green(mode="p2p"), sp.check_taylor_green(mode="dma"),
             sp.check_tearing(mode="dma"), sp.check_double_diffusive(mode="dma"),
             sp.check_tearing(mode="a2a")]
    return {"ok": all(c["ok"] for c in cases), "cases": cases}


# -------------------------------------------------------------- GPU arm, N = 1
def gpu_arm(args, rank, world):
    import torch
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (melvin-b200 has no CPU path)")
    torch.cuda.set_device(local)
    import __graft_entry__ as ge
    if ge._stale():
        ge.build()
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from melvin import _backend, _capi
    from melvin.utility import calc_kinetic_energy, calc_velocity_from_vorticity

    config, nx, nz = args.config, args.nx, args.nz
    peak, peak_src = peaks()
    parity = parity_single_gpu()
    if not parity["ok"]:
        emit({"error": "parity preflight failed", "parity": parity})
        raise SystemExit(3)

    step, o = build_public_loop(config, nx, nz)
    sim, params, w, ux, uz, xp = o["sim"], o["params"], o["w"], o["ux"], o["uz"], o["xp"]
    with ClockSampler(local) as clk:
        ms, launches, marks = time_steps(step, max(args.warmup, 3), args.steps, 1, clk)
    ms_per_step = ms / args.steps
    value = nx * nz * args.steps / (ms * 1e-3)
    bm = byte_model(config, nx, nz)

    # ---- per-entry-point device time (same buffers, back to back, > L2 working set)
    kern = {}

    def timed(name, fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        kern[name] = a.elapsed_time(b) / reps

    roofline = {"bound": "hbm", "kernel": "whole step", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                "achieved": bm["step"] / (ms_per_step * 1e-3) / 1e9, "traffic": None}
    roofline["frac"] = roofline["achieved"] / peak
    if config != "kh":
        traffic, traffic_src = measured_traffic(nx, nz)
        roofline["traffic"], roofline["traffic_source"] = traffic.get("whole step"), traffic_src
        roofline["algorithmic_bytes_per_step"] = bm["step"]
    if config == "kh":
        ctx = w._ctx
        psi, dw, solver = o["psi"], o["dw"], sim.get_laplacian_solver()
        calc_velocity_from_vorticity(w, psi, ux, uz, solver)
        nl = w.vec_dot_nabla(ux.getp(), uz.getp())             # leaves valid intermediates behind
        wptr = w.gets()._t.data_ptr()
        srcs = (ctypes.c_void_p * 3)(wptr, wptr, wptr)
        # field order of Variable.vec_dot_nabla (the scalar first: one source tile load per column tile)
        ops = (ctypes.c_int32 * 3)(_capi.OP_IDENT, _capi.OP_UX, _capi.OP_UZ)
        dsts = (ctypes.c_void_p * 3)(w._i.data_ptr(), ux._i.data_ptr(), uz._i.data_ptr())
        timed("mlv_x_inverse", lambda: ctx.call("mlv_x_inverse", 3, srcs, ops, dsts))
        ia, ib = nl.nls[0][1].ia, nl.nls[0][1].ib
        vp = ctypes.c_void_p
        timed("mlv_advect_z", lambda: ctx.call(
            "mlv_advect_z", vp(ux._i.data_ptr()), vp(uz._i.data_ptr()), vp(w._i.data_ptr()),
            vp(ia.data_ptr()), vp(ib.data_ptr()), None))
        scratch_q = _backend.empty(ctx.spec_shape, np.complex128)
        scratch_f = _backend.zeros(ctx.spec_shape, np.complex128)
        d = _capi.XFwd()
        d.nf, d.mode = 2, 1
        d.src[0], d.src[1] = ia.data_ptr(), ib.data_ptr()
        d.sym[0], d.sym[1] = _capi.SYM_FDX, _capi.SYM_FDZ
        d.coef[0] = d.coef[1] = -1.0
        d.lin = _capi.make_lin_terms([])
        d.integ.ab_order, d.integ.scheme = 2, _capi.SCHEME_SI_LAP
        d.integ.dt, d.integ.alpha, d.integ.lcoef = float(sim._integrator._dt), params.alpha, 1.0 / params.Re
        d.integ.q_in, d.integ.q_out = wptr, scratch_q.data_ptr()
        d.integ.f0, d.integ.fm1 = scratch_f.data_ptr(), dw._level(-1)._t.data_ptr()
        timed("mlv_x_forward", lambda: ctx.call("mlv_x_forward", ctypes.byref(d)))
        dw._pending = None
        dominant = max(kern, key=kern.get)
        achieved = bm[dominant] / (kern[dominant] * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic(nx, nz)
        roofline = {
            "bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic.get(dominant), "traffic_source": traffic_src,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bm[dominant], "ms_per_launch": kern[dominant],
            "kernels": {k: {"ms": v, "algorithmic_bytes": bm[k], "GBps": bm[k] / (v * 1e-3) / 1e9,
                            "frac": bm[k] / (v * 1e-3) / 1e9 / peak} for k, v in kern.items()},
            "step": {"algorithmic_bytes": bm["step"], "GBps": bm["step"] / (ms_per_step * 1e-3) / 1e9,
                     "frac": bm["step"] / (ms_per_step * 1e-3) / 1e9 / peak,
                     "bytes_per_grid_point_step": bm["step"] / (nx * nz)},
        }

    # ---- end to end through the public API with HOST buffers: every step uploads the
    #      spectral state(s) from pinned host memory, steps, and reads the new state(s) and the
    #      kinetic energy back.  (a) blocking: one simulation, Variable.load / on_host as the
    #      reference uses them; (b) streamed: an ensemble of independent simulations on their own
    #      streams (melvin/ensemble.py), every member still doing upload -> step -> read-back
    #      each pass, so the copies of one member overlap the time step of another.
    names = {"kh": ["w"], "tearing": ["w", "j"], "ddc": ["w", "tmp", "xi"], "rbc": ["w", "tmp"]}[config]
    hosts = {n: torch.from_numpy(o[n].on_host()).pin_memory() for n in names}
    sbytes = sum(h.numel() * 16 for h in hosts.values())
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        for n in names:
            o[n].load(hosts[n].numpy(), is_physical=False)         # H2D
        step()
        for n in names:
            o[n].on_host(out=hosts[n])                             # D2H
        torch.cuda.current_stream().synchronize()
        return float(calc_kinetic_energy(ux, uz, xp, params))      # D2H (4 doubles)

    sim.reductions = "always"          # this loop reads the kinetic energy after every step
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ke = e2e_step()
    torch.cuda.synchronize()
    blocking_s = time.perf_counter() - t0

    from melvin.ensemble import Ensemble
    # measured (tools/gpu_r4d.sh): 2 / 4 / 6 / 8 members -> 2.62 / 1.47 / 1.43 / 1.40 ms per pass (PCIe-bound)
    n_members = int(os.environ.get("MLV_E2E_MEMBERS", "8"))
    e2e_passes = max(3 * n_members, min(4 * args.steps, 120))

    def build_member(i):
        mstep, mo = build_public_loop(config, nx, nz)
        mo["sim"].reductions = "always"
        return {"step": mstep, "o": mo, "ke": None,
                "hosts": {n: torch.from_numpy(mo[n].on_host()).pin_memory() for n in names}}

    ens = Ensemble(build_member, members=n_members)

    def ensemble_pass(k):
        with ens.turn(k) as m:                  # the member's previous pass is complete here
            p = m.payload
            if m.passes:                        # consume its result: new state in p["hosts"] + energy
                p["ke"] = float(calc_kinetic_energy(p["o"]["ux"], p["o"]["uz"], xp, params))
            for n in names:
                p["o"][n].load(p["hosts"][n].numpy(), is_physical=False)    # H2D, queued
            p["step"]()
            for n in names:
                p["o"][n].on_host(out=p["hosts"][n])                        # D2H, queued

    # warm-up: at least two passes per member and one second of continuous traffic (the first CUDA process
    # on a fresh box measured 5-10 % slower PCIe round trips than a warm one: r02d 1.57 vs r02c 1.41 ms)
    n_warm, t_w = 0, time.perf_counter()
    while n_warm < 2 * n_members or time.perf_counter() - t_w < 1.0:
        ensemble_pass(n_warm)
        n_warm += 1
    ens.drain()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e2e_passes):
        ensemble_pass(n_warm + k)
    ens.drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": nx * nz * e2e_passes / e2e_s, "unit": UNIT, "h2d_bytes_per_step": sbytes,
           "d2h_bytes_per_step": sbytes + 32, "steps": e2e_passes, "warmup_passes": n_warm,
           "mode": f"streamed ensemble of {n_members} independent simulations (melvin/ensemble.py): every pass "
                   "uploads the member's state from pinned host memory, steps it and reads the new state back; "
                   "copies of one member overlap the step of another",
           "ms_per_step": e2e_s / e2e_passes * 1e3,
           "blocking": {"value": nx * nz * e2e_steps / blocking_s, "steps": e2e_steps,
                        "ms_per_step": blocking_s / e2e_steps * 1e3,
                        "mode": "one simulation, blocking Variable.load / on_host round trip per step"},
           "kinetic_energy": ke}
    del ens

    # free the headline workload before the large grids
    import gc
    del step, o, sim, w, ux, uz, hosts
    _backend._contexts.clear()
    gc.collect()
    torch.cuda.empty_cache()
    large = None if (args.no_large_grid or config != "kh") else large_grid_block(1, 0, peak)

    # ---- CPU baseline: the unmodified reference on this box's host cores (fresh process)
    cpu = None
    if not args.no_cpu_baseline and config == "kh":
        cpu = cpu_baseline_subprocess(nx, nz)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload(config, nx, nz), "grid": [nx, nz],
            "cfl_cadence": params.cfl_cadence, "tracker_cadence": params.tracker_cadence,
            "parallelism": "1 GPU, public melvin API",
            "l2": "working set ~1 GB per step and >= 0.3 GB per kernel launch, larger than the "
                  "126 MB L2; no flush between iterations",
        },
        "clocks": clk.summary(*marks),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": launches,
        "parity": parity,
        "large_grid": large,
    }
    emit(line)


# -------------------------------------------------------------- GPU arm, N > 1
def sharded_arm(args, rank, world):
    """N > 1: the same workload slab-decomposed over the GPUs of the box (kz-slabs / x-slabs,
    exchange of the x-transformed intermediates between the passes; melvin/sharded.py)."""
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if ge._stale() and rank == 0:
        ge.build()
    dist.barrier()
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    from melvin import _backend

    config, nx, nz = args.config, args.nx, args.nz
    peak, peak_src = peaks()
    parity = parity_sharded()
    if not parity["ok"]:
        if rank == 0:
            emit({"error": "sharded parity preflight failed", "n_gpus": world, "parity": parity})
        dist.destroy_process_group()
        raise SystemExit(3)

    st = build_sharded(config, nx, nz)
    with ClockSampler(local) as clk:
        ms, launches, marks = time_steps(st.step, max(args.warmup, 3), args.steps, world, clk)
    value = nx * nz * args.steps / (ms * 1e-3)          # strong scaling: one field for the whole job

    # end to end with host buffers: every step uploads this rank's slab(s) and reads them back
    states = [q for q in (getattr(st, "q", None) or {"w": st.w}).values()]
    hosts = [torch.from_numpy(_backend.to_host(q[st.cur])).pin_memory() for q in states]
    e2e_steps = max(3, min(args.steps, 20))
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for q, h in zip(states, hosts):
            q[st.cur].copy_(h, non_blocking=False)
        st.step()
        for q, h in zip(states, hosts):
            h.copy_(q[st.cur], non_blocking=False)
    dist.barrier()
    torch.cuda.synchronize()
    tt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    slab_bytes = sum(h.numel() * 16 for h in hosts)
    bm = byte_model(config, nx, nz)
    mode, xbytes = st.mode, int(st.bytes_exchanged_per_step)
    cfl_cadence, tracker_cadence = st.cfl_cadence, st.tracker_cadence
    st.close()
    del st, states, hosts
    _backend._contexts.clear()
    torch.cuda.empty_cache()
    large = None if (args.no_large_grid or config != "kh") else large_grid_block(world, rank, peak)
    if rank == 0:
        ms_per_step = ms / args.steps
        nv_ms = xbytes / (NVLINK_GBPS * 1e9) * 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload(config, nx, nz), "grid": [nx, nz],
                "cfl_cadence": cfl_cadence, "tracker_cadence": tracker_cadence,
                "parallelism": f"kz-slabs / x-slabs over {world} GPUs; exchange of the x-transformed "
                + {"p2p": "intermediates fused into the producer kernels (stores into peer memory over "
                          "NVLink), producers and consumers ordered across ranks on the device",
                   "dma": "intermediates by copy-engine transfers into peer memory (one contiguous block per "
                          "peer and field, side streams), the copies of an inverse field overlapping the x "
                          "pass of the next field",
                   "a2a": "intermediates by one asynchronous NCCL all-to-all per field; "
                          "the transfer of an inverse field overlaps the x pass of the next field"}[mode],
                "exchange_mode": mode,
                "l2": "working set per rank and step larger than the 126 MB L2; no flush",
            },
            "clocks": clk.summary(*marks),
            "roofline": {
                "bound": "hbm", "kernel": "whole step (per GPU)", "unit": "GB/s", "peak": peak,
                "achieved": bm["step"] / world / (ms_per_step * 1e-3) / 1e9,
                "frac": bm["step"] / world / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None,
                "peak_source": peak_src,
                "nvlink": {"sent_bytes_per_gpu_per_step": xbytes, "ms_at_770GBps": nv_ms,
                           "frac_of_step": nv_ms / ms_per_step},
            },
            "cpu_baseline": None,
            "e2e": {"value": nx * nz * e2e_steps / float(tt.item()), "unit": UNIT,
                    "h2d_bytes_per_step": slab_bytes, "d2h_bytes_per_step": slab_bytes, "steps": e2e_steps},
            "gpu_launches": launches,
            "parity": parity,
            "large_grid": large,
        }
        emit(line)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="kh", choices=sorted(CONFIGS))
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-grid", action="store_true")
    args = ap.parse_args()
    (dnx, dnz), _ = CONFIGS[args.config]
    args.nx = args.nx or dnx
    args.nz = args.nz or dnz
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    # the contract is ONE JSON line on stdout: route everything else (NCCL prints a version
    # banner to fd 1) to stderr and keep a private handle on the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        reference_arm(args, rank)
    elif world > 1:
        sharded_arm(args, rank, world)
    else:
        gpu_arm(args, rank, world)


if __name__ == "__main__":
    main()
