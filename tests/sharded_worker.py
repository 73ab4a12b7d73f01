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


if case == "tg64":
    gl = np.load(os.path.join(ROOT, "tests", "golden", "loop_tg_64x64.npz"))
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    st = ShardedScalarStepper(64, 64, g.lx, g.lz, float(gl["coef"]), float(gl["dt"]), tracker_cadence=1)
    st.load_spectral(mo.to_spectral(g, mo.ic_taylor_green(g)))
    errs = {}
    for k in range(1, 21):
        st.step()
        if k in (1, 2, 10, 20):
            errs[k] = rel(st.gather_spectral(), gl[f"w_step{k}"])
    ke_err = float(np.max(np.abs(np.array(st.ke) / gl["ke"] - 1)))
    ok = all(e < 1e-12 for e in errs.values()) and ke_err < 1e-9
    if rank == 0:
        print(f"SHARDED world={world} field_err={max(errs.values()):.2e} ke_err={ke_err:.2e} "
              f"{'OK' if ok else 'FAIL'}", flush=True)
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
elif case == "ddc":
    # three coupled scalars (BASELINE config 4 loop) vs the golden of the unmodified reference
    gl = np.load(os.path.join(ROOT, "tests", "golden", "loop_ddc_64x64.npz"))
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    st = ShardedDoubleDiffusiveStepper(64, 64, g.lx, g.lz, float(gl["Pr"]), float(gl["R0"]), float(gl["tau"]),
                                       float(gl["dt"]), tracker_cadence=1)
    noise = mo.to_spectral(g, mo.ic_noise(g))
    st.load_spectral(noise, noise, noise)
    errs = {}
    for k in range(1, 21):
        st.step()
        if k in (1, 10, 20):
            got = st.gather_spectral()
            errs[k] = max(rel(a, gl[f"{nm}_step{k}"]) for a, nm in zip(got, ("w", "tmp", "xi")))
    ke_err = float(np.max(np.abs(np.array(st.ke) / gl["ke"] - 1)))
    nu_err = float(np.max(np.abs((np.array(st.nu) - 1) - (gl["nu"] - 1)) / np.maximum(np.abs(gl["nu"] - 1), 1e-17)))
    ok = all(e < 1e-12 for e in errs.values()) and ke_err < 1e-9 and nu_err < 1e-6
    if rank == 0:
        print(f"SHARDED world={world} field_err={max(errs.values()):.2e} ke_err={ke_err:.2e} "
              f"nu_err={nu_err:.2e} {'OK' if ok else 'FAIL'}", flush=True)
elif case == "tearing":
    # MHD loop (BASELINE config 5 loop), two exchange rounds per step, vs the reference golden
    gl = np.load(os.path.join(ROOT, "tests", "golden", "loop_tearing_64x64.npz"))
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    st = ShardedTearingStepper(64, 64, g.lx, g.lz, float(gl["Re"]), float(gl["S"]), float(gl["dt"]),
                               tracker_cadence=1)
    st.load_spectral(np.zeros(g.spectral_shape, complex), mo.to_spectral(g, gl["j0_phys"]))
    jerr, werr = 0.0, 0.0
    for k in range(1, 21):
        st.step()
        if k in (1, 10, 20):
            w, j = st.gather_spectral()
            jerr = max(jerr, rel(j, gl[f"j_step{k}"]))
            werr = max(werr, rel(w, gl[f"w_step{k}"]))
    # w starts from exactly zero and is driven by rounding-level asymmetries of j
    ok = jerr < 1e-12 and werr < 1e-9
    if rank == 0:
        print(f"SHARDED world={world} j_err={jerr:.2e} w_err={werr:.2e} {'OK' if ok else 'FAIL'}", flush=True)
st.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
