"""Parity checks of the slab-decomposed steppers against the committed goldens of the
unmodified reference (tests/golden/loop_*_64x64.npz): shared by tests/sharded_worker.py
(gloo on the emulation build, NCCL on GPUs) and by the parity preflight of bench.py's
multi-GPU arm.  Every function runs on all ranks of an initialised process group and
returns a dict with the worst errors and ``ok``."""
import os

import numpy as np

from melvin.sharded import (ShardedDoubleDiffusiveStepper, ShardedScalarStepper,
                            ShardedTearingStepper)
from oracle import melvin_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(np.asarray(b).ravel()))


def check_taylor_green(mode=None, steps=20):
    """examples/taylor_green_vortex.py:85-95 at 64^2: fields 1e-12, kinetic energy 1e-9."""
    gl = np.load(os.path.join(GOLDEN, "loop_tg_64x64.npz"))
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    st = ShardedScalarStepper(64, 64, g.lx, g.lz, float(gl["coef"]), float(gl["dt"]), tracker_cadence=1,
                              mode=mode)
    st.load_spectral(mo.to_spectral(g, mo.ic_taylor_green(g)))
    errs = []
    for k in range(1, steps + 1):
        st.step()
        if k in (1, 2, 10, 20):
            errs.append(rel(st.gather_spectral(), gl[f"w_step{k}"]))
    ke_err = float(np.max(np.abs(np.array(st.ke) / gl["ke"][:steps] - 1)))
    res = {"case": "taylor_green_64x64", "exchange_mode": st.mode, "steps": steps,
           "field_rel_l2": max(errs), "ke_rel": ke_err}
    res["ok"] = bool(res["field_rel_l2"] < 1e-12 and ke_err < 1e-9)
    st.close()
    return res


def check_double_diffusive(mode=None, steps=20):
    """examples/double_diffusive_convection.py:100-126 at 64^2 (three coupled scalars)."""
    gl = np.load(os.path.join(GOLDEN, "loop_ddc_64x64.npz"))
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    st = ShardedDoubleDiffusiveStepper(64, 64, g.lx, g.lz, float(gl["Pr"]), float(gl["R0"]), float(gl["tau"]),
                                       float(gl["dt"]), tracker_cadence=1, mode=mode)
    noise = mo.to_spectral(g, mo.ic_noise(g))
    st.load_spectral(noise, noise, noise)
    errs = []
    for k in range(1, steps + 1):
        st.step()
        if k in (1, 10, 20):
            got = st.gather_spectral()
            errs.append(max(rel(a, gl[f"{nm}_step{k}"]) for a, nm in zip(got, ("w", "tmp", "xi"))))
    ke_err = float(np.max(np.abs(np.array(st.ke) / gl["ke"][:steps] - 1)))
    nu, nu_ref = np.array(st.nu) - 1, gl["nu"][:steps] - 1
    nu_err = float(np.max(np.abs(nu - nu_ref) / np.maximum(np.abs(nu_ref), 1e-17)))
    res = {"case": "double_diffusive_64x64", "exchange_mode": st.mode, "steps": steps,
           "field_rel_l2": max(errs), "ke_rel": ke_err, "nusselt_minus_one_rel": nu_err}
    res["ok"] = bool(res["field_rel_l2"] < 1e-12 and ke_err < 1e-9 and nu_err < 1e-6)
    st.close()
    return res


def check_tearing(mode=None, steps=20):
    """examples/resistive_tearing_instability.py:125-148 at 64^2 (two exchange rounds per step).
    The vorticity starts from exactly zero and is driven by rounding-level asymmetries of j,
    hence the looser gate on w."""
    gl = np.load(os.path.join(GOLDEN, "loop_tearing_64x64.npz"))
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    st = ShardedTearingStepper(64, 64, g.lx, g.lz, float(gl["Re"]), float(gl["S"]), float(gl["dt"]),
                               tracker_cadence=1, mode=mode)
    st.load_spectral(np.zeros(g.spectral_shape, complex), mo.to_spectral(g, gl["j0_phys"]))
    jerr = werr = 0.0
    for k in range(1, steps + 1):
        st.step()
        if k in (1, 10, 20):
            w, j = st.gather_spectral()
            jerr = max(jerr, rel(j, gl[f"j_step{k}"]))
            werr = max(werr, rel(w, gl[f"w_step{k}"]))
    res = {"case": "tearing_64x64", "exchange_mode": st.mode, "steps": steps,
           "field_rel_l2": jerr, "w_rel_l2": werr}
    res["ok"] = bool(jerr < 1e-12 and werr < 1e-9)
    st.close()
    return res
