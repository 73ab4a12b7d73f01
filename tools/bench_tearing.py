#!/usr/bin/env python3
"""BASELINE config 5 (resistive tearing, MHD, fully spectral) at a chosen grid: grid-point-steps/s
of the reference loop (examples/resistive_tearing_instability.py:125-148).
  1 GPU : the loop body through the public melvin API (as tests/parity_cases.run_tearing)
  N GPUs: melvin.sharded.ShardedTearingStepper under torch.distributed.run
usage: bench_tearing.py [--nx 16384 --nz 16384 --steps 10 --warmup 3]; prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "melvin.py_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=16384)
ap.add_argument("--nz", type=int, default=16384)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
real_stdout = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

from oracle import melvin_oracle as mo  # noqa: E402  (initial condition only)

nx, nz, lx, lz, Re, S = args.nx, args.nz, 16.0 / 9.0, 1.0, 1e6, 1e6
g = mo.Grid(nx, nz, lx, lz)
dt = 0.01 * 0.05 * lx / nx
j0 = mo.ic_tearing_current(g)

if world == 1:
    from functools import partial
    import parity_cases as pc
    from melvin import b200 as xp
    from melvin.utility import calc_kinetic_energy, calc_velocity_from_vorticity
    os.chdir(tempfile.mkdtemp(prefix="mlvtear"))
    d = pc.base_params(nx, nz, lx, lz, initial_dt=dt, Re=Re, S=S, spatial_derivative_order=2,
                       integrator_order=2, integrator="semi-implicit", tracker_cadence=100)
    p, sim, (w, j), (dw, dj), psi, ux, uz = pc.make_sim(d, ["w", "j"], ["dw", "dj"], [pc.CE, pc.CE])
    phi, bx, bz = (sim.make_variable(n, [pc.CE, pc.CE]) for n in ("phi", "bx", "bz"))
    sim.config_scalar_trackers({"ke": partial(calc_kinetic_energy, ux, uz, xp, p)})
    j.load(j0, is_physical=True)
    solver = sim.get_laplacian_solver()

    def step():
        calc_velocity_from_vorticity(w, psi, ux, uz, solver)
        calc_velocity_from_vorticity(j, phi, bx, bz, solver)
        lin_op = 1.0 / p.Re * w.lap()
        dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp()) + j.vec_dot_nabla(bx.getp(), bz.getp())
        sim._integrator.integrate(w, dw, lin_op)
        lin_op = 1.0 / p.S * j.lap()
        dj[:] = -j.vec_dot_nabla(ux.getp(), uz.getp()) + w.vec_dot_nabla(bx.getp(), bz.getp())
        sim._integrator.integrate(j, dj, lin_op)
        sim.end_loop()
    mode = "public API, 1 GPU"
else:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from melvin import _backend
    from melvin.sharded import ShardedTearingStepper
    import ctypes
    st = ShardedTearingStepper(nx, nz, lx, lz, Re, S, dt)
    full = _backend.Context(nx, nz, lx, lz, False, 2)
    phys = _backend.from_host(j0)
    spec = _backend.empty(full.spec_shape, np.complex128)
    full.call("mlv_to_spectral", ctypes.c_void_p(phys.data_ptr()), ctypes.c_void_p(full.scratch_i().data_ptr()),
              ctypes.c_void_p(spec.data_ptr()))
    jhat = _backend.to_host(spec)
    del phys, spec, full
    torch.cuda.empty_cache()
    st.load_spectral(np.zeros_like(jhat), jhat)
    step = st.step
    mode = f"slab-decomposed, {world} GPUs, exchange {st.mode}"


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for _ in range(max(3, args.warmup)):
    step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
barrier()
ms = e0.elapsed_time(e1)
if world > 1:
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
if rank == 0:
    real_stdout.write(json.dumps({
        "metric": "grid-point-timesteps/sec (fp64)", "value": nx * nz * args.steps / (ms * 1e-3),
        "unit": "grid-point-steps/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
        "scaling": "strong", "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"resistive tearing (MHD) {nx}x{nz}, AB2 + semi-implicit (BASELINE configs[4])",
                   "parallelism": mode}}) + "\n")
    real_stdout.flush()
if world > 1:
    st.close()
    dist.destroy_process_group()
