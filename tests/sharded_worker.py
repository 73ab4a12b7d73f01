"""Worker of the slab-decomposition tests: run under torch.distributed.run.
  backend "emu": CPU emulation build of the kernels + gloo (host-logic test, CPU box)
  backend "cuda": real library + NCCL (GPU box, one rank per GPU)
Runs the sharded Taylor-Green loop and checks the gathered state against the golden
of the unmodified reference (tests/golden/loop_tg_64x64.npz)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"),
          os.path.join(ROOT, "melvin.py_b200")):
    sys.path.insert(0, p)

backend = sys.argv[1] if len(sys.argv) > 1 else "emu"
case = sys.argv[2] if len(sys.argv) > 2 else "tg64"
from melvin import _backend  # noqa: E402

if backend == "emu":
    import emu_harness as eh
    _backend._install(eh.lib(), "cpu")
    dist.init_process_group("gloo")
else:
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))

from melvin.sharded import (ShardedDoubleDiffusiveStepper, ShardedScalarStepper,  # noqa: E402
                            ShardedTearingStepper)
from oracle import melvin_oracle as mo  # noqa: E402

rank, world = dist.get_rank(), dist.get_world_size()


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / np.linalg.norm(np.asarray(b).ravel())


def report(res):
    if rank == 0:
        print(f"SHARDED world={world} " + " ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}"
                                                   for k, v in res.items()) + (" OK" if res["ok"] else " FAIL"),
              flush=True)
    return res["ok"]


st = None
if case == "tg64":
    import sharded_parity as sp
    ok = report(sp.check_taylor_green())
elif case in ("kh", "khlong"):
    # uneven column split (nm not a multiple of the rank count), order-2 KH vs the oracle;
    # "khlong": the long-line kernels (16384-point lines) forced onto a small grid
    nx, nz = (128, 64) if backend == "emu" else (1024, 512)
    if case == "khlong":
        os.environ["MLV_FORCE_SPLIT"] = "3"
        nx, nz = 128, 256
    g = mo.Grid(nx, nz, 16.0 / 9.0, 1.0)
    w0 = mo.ic_kelvin_helmholtz(g)
    dt = 0.05 * g.lx / nx
    want, run, _ = mo.run_single_scalar(g, w0, 1e-5, dt, 12, tracker_cadence=1)
    st = ShardedScalarStepper(nx, nz, g.lx, g.lz, 1e-5, dt, tracker_cadence=1)
    if case == "khlong":
        assert st.ctx.lib.mlv_long_lines(st.ctx.handle) == 3
    st.load_spectral(mo.to_spectral(g, w0))
    for _ in range(12):
        st.step()
    err = rel(st.gather_spectral(), want)
    ke_err = float(np.max(np.abs(np.array(st.ke) / np.array(run.ke) - 1)))
    ok = err < 1e-12 and ke_err < 1e-9 and abs(st.dt - run.dt) < 1e-18
    if rank == 0:
        print(f"SHARDED world={world} field_err={err:.2e} ke_err={ke_err:.2e} {'OK' if ok else 'FAIL'}",
              flush=True)
elif case.startswith("api_"):
    # N4: the UNCHANGED public-API loops of tests/parity_cases.py under the process group --
    # Simulation / Variable shard themselves (melvin/_dist.py) -- vs the reference goldens
    import parity_cases as pc
    G = lambda name: np.load(os.path.join(ROOT, "tests", "golden", name))   # noqa: E731
    errs = {}
    if case == "api_tg":
        gl = G("loop_tg_64x64.npz")
        g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
        with pc.scratch_cwd():
            out = pc.run_single_scalar(64, 64, g.lx, g.lz, float(gl["coef"]), float(gl["dt"]), 20,
                                       mo.ic_taylor_green(g), snaps=(1, 2, 10, 20), strict_reads=True)
        errs["field"] = max(rel(out[f"w_step{k}"], gl[f"w_step{k}"]) for k in (1, 2, 10, 20))
        errs["ke"] = float(np.max(np.abs(out["ke"] / gl["ke"] - 1)))
        vel = mo.velocity_from_vorticity(g, mo.to_spectral(g, mo.ic_taylor_green(g)))
        errs["psi_read_after_update"] = rel(out["psi_after_step1"], vel["psi_s"])
        errs["ux_read_after_update"] = rel(out["ux_p_after_step1"], vel["ux_p"])
        ok = errs["field"] < 1e-12 and errs["ke"] < 1e-9 and errs["psi_read_after_update"] < 1e-12 \
            and errs["ux_read_after_update"] < 1e-12
    elif case == "api_ddc":
        gl = G("loop_ddc_64x64.npz")
        with pc.scratch_cwd():
            out = pc.run_ddc(64, 64, float(gl["lx"]), float(gl["lz"]), float(gl["dt"]), 10,
                             float(gl["Pr"]), float(gl["R0"]), float(gl["tau"]), snaps=(1, 10))
        errs["field"] = max(rel(out[f"{nm}_step{k}"], gl[f"{nm}_step{k}"]) for k in (1, 10) for nm in ("w", "tmp", "xi"))
        errs["ke"] = float(np.max(np.abs(out["ke"] / gl["ke"][:10] - 1)))
        errs["nu"] = float(np.max(np.abs((out["nu"] - 1) - (gl["nu"][:10] - 1)) / np.maximum(np.abs(gl["nu"][:10] - 1), 1e-17)))
        ok = errs["field"] < 1e-12 and errs["ke"] < 1e-9 and errs["nu"] < 1e-6
    else:
        gl = G("loop_tearing_64x64.npz")
        with pc.scratch_cwd():
            out = pc.run_tearing(64, 64, float(gl["lx"]), float(gl["lz"]), float(gl["dt"]), 10,
                                 float(gl["Re"]), float(gl["S"]), gl["j0_phys"], snaps=(1, 10))
        errs["j"] = max(rel(out[f"j_step{k}"], gl[f"j_step{k}"]) for k in (1, 10))
        errs["w"] = max(rel(out[f"w_step{k}"], gl[f"w_step{k}"]) for k in (1, 10))
        ok = errs["j"] < 1e-12 and errs["w"] < 1e-9
    if rank == 0:
        print(f"SHARDED world={world} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items())
              + (" OK" if ok else " FAIL"), flush=True)
elif case == "bench":
    # bench.py's multi-GPU host logic: the parity preflight and the three workload builders
    import bench
    res = bench.parity_sharded()
    ok = res["ok"] and len(res["cases"]) == 5
    for cfg in ("kh", "ddc", "tearing"):
        bst = bench.build_sharded(cfg, 64, 64)
        for _ in range(3):
            bst.step()
        ok = ok and all(np.all(np.isfinite(a)) for a in np.atleast_1d(bst.gather_spectral()))
        bst.close()
    if rank == 0:
        print(f"SHARDED world={world} bench preflight + builders {'OK' if ok else 'FAIL'}", flush=True)
elif case == "ddc":
    import sharded_parity as sp
    ok = report(sp.check_double_diffusive())
elif case == "tearing":
    import sharded_parity as sp
    ok = report(sp.check_tearing())
if st is not None:
    st.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
