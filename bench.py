#!/usr/bin/env python3
"""Benchmark of the per-timestep pseudo-spectral hot path (BASELINE.json metric:
grid-point-timesteps/s, fp64).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (N=1): BASELINE configs[1] -- Kelvin-Helmholtz, fully spectral, 4096x4096,
AB2 + semi-implicit diffusion; the loop body is the example script's
(reference examples/kelvin_helmholtz_instability.py:115-131) driven through the
drop-in `melvin` API, which calls libmelvin_b200.so.  One "step" = one time step of
the whole 4096^2 field.  Prints ONE JSON line (see DESIGN.md section "Measurement").
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "melvin.py_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "grid-point-timesteps/sec (fp64)"
UNIT = "grid-point-steps/s"
RE = 1e5                       # examples/kelvin_helmholtz_instability.py:64


def kh_params(nx, nz):
    lx, lz = 16.0 / 9.0, 1.0
    return {
        "nx": nx, "nz": nz, "lx": lx, "lz": lz, "Re": RE, "final_time": 1e9,
        "save_cadence": 1e9, "dump_cadence": 1e9, "precision": "double",
        "spatial_derivative_order": 2, "integrator_order": 2, "integrator": "semi-implicit",
        "cfl_cutoff": 0.5, "initial_dt": 0.05 * lx / nx,
    }


def byte_model(nx, nz):
    """Algorithmic bytes of SURVEY section 8(d) for one advected scalar (config 2)."""
    nn, nm = (nx - 1) // 3, (nz - 1) // 3
    S = 16 * (2 * nn + 1) * nm
    I = 16 * nx * nm
    return {
        "S": S, "I": I,
        "step": 5 * S + 8 * I,               # 2*I*(N_inv+N_fwd) + S*N_state, N_inv=3 N_fwd=1 N_state=5
        "mlv_x_inverse": S + 3 * I,          # read w-hat, write the 3 x-transformed fields
        "mlv_advect_z": 4 * I,               # read 3 I, write 1 I (the model's single forward field)
        "mlv_x_forward": I + 4 * S,          # read 1 I, w-hat, f(-1); write w-hat, f(0)
    }


def measured_traffic(nx, nz):
    """DRAM bytes per launch from the committed ncu capture (same grid only), else None."""
    path = os.path.join(ROOT, "profiles", "r01c_traffic.json")
    try:
        with open(path) as fp:
            d = json.load(fp)
        if list(d["grid"]) == [nx, nz]:
            return d["bytes_per_launch"], d["source"]
    except (OSError, ValueError, KeyError):
        pass
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


# ------------------------------------------------------------- CPU arms
def oracle_steps(nx, nz, nwarm, nsteps):
    """Time `nsteps` steps of the oracle port (NumPy float64, pocketfft, 1 thread)
    of the same loop on the same initial condition.  Returns (gp-steps/s, s/step)."""
    from oracle import melvin_oracle as mo
    d = kh_params(nx, nz)
    g = mo.Grid(nx, nz, d["lx"], d["lz"])
    run = mo.Run(g, d["initial_dt"])
    w = mo.to_spectral(g, mo.ic_kelvin_helmholtz(g))
    dw = mo.History(g)
    for _ in range(nwarm):
        w = mo.step_single_scalar(g, run, w, dw, 1.0 / RE)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        w = mo.step_single_scalar(g, run, w, dw, 1.0 / RE)
    dt = time.perf_counter() - t0
    return nx * nz * nsteps / dt, dt / nsteps


def reference_arm(args, rank):
    """`--impl reference`: the reference's own CPU algorithm (NumPy port in oracle/,
    because the pure-Python reference checkout does not exist on the GPU box)."""
    if rank != 0:
        return
    total = args.steps + args.warmup
    # bounded sample: keep the whole run within a few minutes (about 3.3 / 0.7 / 0.18 s per step)
    n = 4096 if total <= 50 else (2048 if total <= 250 else 1024)
    n = min(n, args.nx)
    value, s_per_step = oracle_steps(n, n, args.warmup, args.steps)
    sample = (f"{args.steps} timed + {args.warmup} warm-up steps of the oracle port "
              f"(oracle/melvin_oracle.py, numpy.fft) on a {n}x{n} Kelvin-Helmholtz grid")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Kelvin-Helmholtz 4096x4096 fully spectral AB2 semi-implicit "
                               "(BASELINE configs[1])", "sample_grid": [n, n]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# -------------------------------------------------------------- GPU arm
def gpu_arm(args, rank, world):
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (melvin-b200 has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as ge
    if ge._stale():
        ge.build()
    from functools import partial
    import melvin
    from melvin import BasisFunctions, Parameters, Simulation, _backend, _capi
    from melvin import b200 as xp
    from melvin.utility import calc_kinetic_energy, calc_velocity_from_vorticity
    from oracle import melvin_oracle as mo        # initial condition only (host, like the scripts)

    nx, nz = args.nx, args.nz
    pd = kh_params(nx, nz)
    scratch = tempfile.mkdtemp(prefix="mlvbench")
    os.chdir(scratch)
    params = Parameters(pd)
    sim = Simulation(params, xp)
    basis = [BasisFunctions.COMPLEX_EXP, BasisFunctions.COMPLEX_EXP]
    w = sim.make_variable("w", basis)
    dw = sim.make_derivative("dw")
    psi = sim.make_variable("psi", basis)
    ux = sim.make_variable("ux", basis)
    uz = sim.make_variable("uz", basis)
    sim.init_laplacian_solver(basis)
    sim.config_cfl(ux, uz)
    sim.config_scalar_trackers({"kinetic_energy.npz": partial(calc_kinetic_energy, ux, uz, xp, params)})
    g = mo.Grid(nx, nz, pd["lx"], pd["lz"])
    w0 = torch.from_numpy(mo.ic_kelvin_helmholtz(g)).pin_memory()
    w.load(w0.numpy(), is_physical=True)
    solver = sim.get_laplacian_solver()

    def step():
        # examples/kelvin_helmholtz_instability.py:115-131
        calc_velocity_from_vorticity(w, psi, ux, uz, solver)
        lin_op = 1.0 / params.Re * w.lap()
        dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
        sim._integrator.integrate(w, dw, lin_op)
        sim.end_loop()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (state in HBM, CUDA events, max over ranks)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        n_before = _backend.launches()
        t_begin = clk.mark()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        t_end = clk.mark()
        launches = _backend.launches() - n_before
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * nx * nz * args.steps / (ms * 1e-3)     # replicas: every rank steps a full field

    # ---- per-entry-point device time (same buffers, back to back, > L2 working set)
    bm = byte_model(nx, nz)
    peak, peak_src = peaks()
    ctx = w._ctx
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

    import ctypes
    calc_velocity_from_vorticity(w, psi, ux, uz, solver)
    nl = w.vec_dot_nabla(ux.getp(), uz.getp())             # leaves valid intermediates behind
    wptr = w.gets()._t.data_ptr()
    srcs = (ctypes.c_void_p * 3)(wptr, wptr, wptr)
    ops = (ctypes.c_int32 * 3)(_capi.OP_UX, _capi.OP_UZ, _capi.OP_IDENT)
    dsts = (ctypes.c_void_p * 3)(ux._i.data_ptr(), uz._i.data_ptr(), w._i.data_ptr())
    timed("mlv_x_inverse", lambda: ctx.call("mlv_x_inverse", 3, srcs, ops, dsts))
    ia, ib = nl.nls[0][1].ia, nl.nls[0][1].ib
    red4 = torch.empty(4, dtype=torch.float64, device="cuda")
    timed("mlv_advect_z", lambda: ctx.call(
        "mlv_advect_z", ctypes.c_void_p(ux._i.data_ptr()), ctypes.c_void_p(uz._i.data_ptr()),
        ctypes.c_void_p(w._i.data_ptr()), ctypes.c_void_p(ia.data_ptr()),
        ctypes.c_void_p(ib.data_ptr()), ctypes.c_void_p(red4.data_ptr())))
    scratch_q = torch.empty_like(w.gets()._t)
    scratch_f = torch.zeros_like(w.gets()._t)
    d = _capi.XFwd()
    d.nf, d.mode = 2, 1
    d.src[0], d.src[1] = ia.data_ptr(), ib.data_ptr()
    d.sym[0], d.sym[1] = _capi.SYM_FDX, _capi.SYM_FDZ
    d.coef[0] = d.coef[1] = -1.0
    d.lin = _capi.make_lin_terms([])
    d.integ.ab_order, d.integ.scheme = 2, _capi.SCHEME_SI_LAP
    d.integ.dt, d.integ.alpha, d.integ.lcoef = float(sim._integrator._dt), params.alpha, 1.0 / RE
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
    #      spectral state from pinned host memory, steps, and reads the new state and the
    #      kinetic energy back.
    S = bm["S"]
    host_state = torch.from_numpy(w.on_host()).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        w.load(host_state.numpy(), is_physical=False)
        step()
        host_state.copy_(w.gets()._t, non_blocking=False)
    barrier()
    t0 = time.perf_counter()
    ke = 0.0
    for _ in range(e2e_steps):
        w.load(host_state.numpy(), is_physical=False)          # H2D  (S bytes)
        step()
        host_state.copy_(w.gets()._t, non_blocking=False)      # D2H  (S bytes)
        ke = float(calc_kinetic_energy(ux, uz, xp, params))    # D2H  (2 doubles)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * nx * nz * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": S,
           "d2h_bytes_per_step": S + 16, "steps": e2e_steps, "kinetic_energy": ke}

    # ---- CPU baseline: oracle port on this box's host cores (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cv, cs = oracle_steps(nx, nz, 1, 3)
        cpu = {"value": cv, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"3 timed + 1 warm-up steps of the same {nx}x{nz} workload, oracle port "
                         f"(NumPy float64, pocketfft, single thread; host has {os.cpu_count()} cores)",
               "s_per_step": cs}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"Kelvin-Helmholtz {nx}x{nz} fully spectral, AB2 + semi-implicit diffusion "
                            "(BASELINE configs[1]; loop of examples/kelvin_helmholtz_instability.py)",
                "grid": [nx, nz], "cfl_cadence": params.cfl_cadence,
                "tracker_cadence": params.tracker_cadence,
                "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas "
                               "(z-slab sharded path: see DESIGN.md, multi-GPU)",
                "l2": "working set ~1 GB per step and >= 0.4 GB per kernel launch, larger than the "
                      "126 MB L2; no flush between iterations",
            },
            "clocks": clk.summary(t_begin, t_end),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def sharded_arm(args, rank, world):
    """N > 1: the same Kelvin-Helmholtz workload, slab-decomposed over the GPUs of the
    box (kz-slabs / x-slabs, NCCL all-to-all between the passes; melvin/sharded.py)."""
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if ge._stale() and rank == 0:
        ge.build()
    dist.barrier()
    from melvin import _backend
    from melvin.sharded import ShardedScalarStepper
    from oracle import melvin_oracle as mo

    nx, nz = args.nx, args.nz
    pd = kh_params(nx, nz)
    st = ShardedScalarStepper(nx, nz, pd["lx"], pd["lz"], 1.0 / RE, pd["initial_dt"])
    g = mo.Grid(nx, nz, pd["lx"], pd["lz"])
    # initial condition: transformed once on a full-size context, then only this rank's
    # column slab is kept (set-up, outside every timed region)
    full_ctx = _backend.Context(nx, nz, pd["lx"], pd["lz"], False, 2)
    import ctypes
    phys = _backend.from_host(mo.ic_kelvin_helmholtz(g))
    spec = _backend.empty(full_ctx.spec_shape, np.complex128)
    full_ctx.call("mlv_to_spectral", ctypes.c_void_p(phys.data_ptr()),
                  ctypes.c_void_p(full_ctx.scratch_i().data_ptr()), ctypes.c_void_p(spec.data_ptr()))
    st.load_spectral(_backend.to_host(spec))
    del phys, spec, full_ctx
    torch.cuda.empty_cache()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        for _ in range(max(args.warmup, 3)):
            st.step()
        barrier()
        n_before = _backend.launches()
        t_begin = clk.mark()
        e0.record()
        for _ in range(args.steps):
            st.step()
        e1.record()
        barrier()
        t_end = clk.mark()
        launches = _backend.launches() - n_before
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = nx * nz * args.steps / (ms * 1e-3)          # strong scaling: one field for the whole job

    # end to end with host buffers: every step uploads this rank's slab and reads it back
    host = torch.from_numpy(_backend.to_host(st.w[st.cur])).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        st.w[st.cur].copy_(host, non_blocking=False)
        st.step()
        host.copy_(st.w[st.cur], non_blocking=False)
    barrier()
    tt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    slab_bytes = st.rows * st.nml * 16
    bm = byte_model(nx, nz)
    peak, peak_src = peaks()
    if rank == 0:
        ms_per_step = ms / args.steps
        xbytes = st.bytes_exchanged_per_step                # sent per rank and step
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"Kelvin-Helmholtz {nx}x{nz} fully spectral, AB2 + semi-implicit diffusion "
                            "(BASELINE configs[1]), slab-decomposed",
                "grid": [nx, nz],
                "parallelism": f"kz-slabs / x-slabs over {world} GPUs; exchange of the x-transformed "
                + {"p2p": "intermediates fused into the producer kernels (stores into peer memory over "
                          "NVLink), ordered by 2 one-element all-reduces per step",
                   "dma": "intermediates by copy-engine transfers into peer memory (one contiguous block per "
                          "peer and field, side stream), the copies of an inverse field overlapping the x "
                          "pass of the next field; ordered by 2 one-element all-reduces per step",
                   "a2a": "intermediates by one asynchronous NCCL all-to-all per field (3 + 2 per step); "
                          "the transfer of an inverse field overlaps the x pass of the next field"}[st.mode],
                "exchange_mode": st.mode,
                "cfl_cadence": st.cfl_cadence,
                "tracker_cadence": st.tracker_cadence,
                "l2": "working set per rank and step larger than the 126 MB L2; no flush",
            },
            "clocks": clk.summary(t_begin, t_end),
            "roofline": {
                "bound": "hbm", "kernel": "whole step (per GPU)", "unit": "GB/s", "peak": peak,
                "achieved": bm["step"] / world / (ms_per_step * 1e-3) / 1e9,
                "frac": bm["step"] / world / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None,
                "peak_source": peak_src,
                "nvlink": {"sent_bytes_per_gpu_per_step": xbytes,
                           "ms_at_770GBps": xbytes / 770e9 * 1e3},
            },
            "cpu_baseline": None,
            "e2e": {"value": nx * nz * e2e_steps / float(tt.item()), "unit": UNIT,
                    "h2d_bytes_per_step": slab_bytes, "d2h_bytes_per_step": slab_bytes, "steps": e2e_steps},
            "gpu_launches": launches,
        }
        emit(line)
    st.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--nz", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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
    else:
        if True:
            if world > 1:
                sharded_arm(args, rank, world)
            else:
                gpu_arm(args, rank, world)


if __name__ == "__main__":
    main()
