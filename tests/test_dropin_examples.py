"""North-star claim "existing example scripts and JSON parameter files run unchanged":
the reference's own example scripts are executed UNMODIFIED (runpy) on the drop-in package
-- `import cupy; xp = cupy` resolves to melvin.py_b200/shims/cupy, `"precision": "single"` is
promoted to float64 -- and, in a second process, on the unmodified reference with NumPy
float64.  Only run parameters (grid size, number of steps, cadences) are overridden, through a
patched ``Parameters`` class; see tests/example_runner.py.  Outputs are compared file by file.

Runs in the CPU container on the host emulation build of the kernels (development harness);
skipped where the reference checkout is absent (the GPU box).  The same comparison on a GPU:
``python tests/example_runner.py --impl gpu ...``.
"""
import glob
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
RUNNER = os.path.join(ROOT, "tests", "example_runner.py")

pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF, "examples")) or shutil.which("g++") is None,
    reason="needs the reference checkout (/root/reference) and g++ for the emulation build")

CASES = {
    # script: (overrides, steps)
    "taylor_green_vortex": ({"nx": 64, "nz": 64, "tracker_cadence": 1, "save_cadence": 4e-3}, 12),
    "kelvin_helmholtz_instability": ({"nx": 64, "nz": 32, "tracker_cadence": 1, "save_cadence": 1e-2,
                                      "initial_dt": 1e-3}, 10),
    "double_diffusive_convection": ({"nx": 64, "nz": 32, "tracker_cadence": 1, "save_cadence": 4e-3,
                                     "dump_cadence": 6e-3}, 10),
    "rayleigh_benard_convection": ({"nx": 64, "nz": 13, "tracker_cadence": 1, "save_cadence": 4e-6,
                                    "dump_cadence": 6e-6}, 10),
    "vortex_pairs": ({"nx": 64, "nz": 32, "tracker_cadence": 1, "save_cadence": 1e-2,
                      "initial_dt": 1e-3}, 8),
    # hard-codes `xp = np` (examples/resistive_tearing_instability.py:19-20): runs on the device under
    # MELVIN_B200_NUMPY_IS_DEVICE=1 (without the switch numpy as xp raises BackendUnavailable)
    "resistive_tearing_instability": ({"nx": 64, "nz": 64, "tracker_cadence": 1}, 8),
}
NUMPY_XP = {"resistive_tearing_instability"}


def run(impl, script, overrides, steps, out, numpy_is_device=False):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONPATH", None)
    env.pop("MELVIN_B200_NUMPY_IS_DEVICE", None)
    if numpy_is_device:
        env["MELVIN_B200_NUMPY_IS_DEVICE"] = "1"
    subprocess.run([sys.executable, RUNNER, "--impl", impl, "--script", script, "--overrides",
                    json.dumps(overrides), "--steps", str(steps), "--out", out],
                   check=True, env=env, timeout=600)


def rel(a, b):
    den = np.linalg.norm(np.asarray(b).ravel())
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (den if den > 0 else 1.0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_example_runs_unmodified(name, tmp_path):
    overrides, steps = CASES[name]
    script = os.path.join(REF, "examples", name + ".py")
    ours, ref = str(tmp_path / "ours"), str(tmp_path / "ref")
    if name in NUMPY_XP:
        with pytest.raises(subprocess.CalledProcessError):        # no CPU path: numpy as xp is refused ...
            run("emu", script, overrides, steps, ours)
    run("emu", script, overrides, steps, ours, numpy_is_device=name in NUMPY_XP)     # ... unless asked for
    run("reference", script, overrides, steps, ref)
    assert int(open(os.path.join(ours, "launches.txt")).read()) > 0
    if name in NUMPY_XP:
        assert "xp = numpy requested, computing on the device" in open(os.path.join(ours, "warnings.txt")).read()
    assert "precision='single' is computed in float64" in open(os.path.join(ours, "warnings.txt")).read()
    # the JSON parameter file the script wrote is the same, and reloads into the same Parameters
    pj = json.load(open(os.path.join(ours, "params.json")))
    assert pj["precision"] == "single"        # as the script wrote it; the reference run is the double path
    assert pj == dict(json.load(open(os.path.join(ref, "params.json"))), precision="single")
    sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))
    from melvin import Parameters
    with pytest.warns(UserWarning):
        p2 = Parameters(dict(pj, **overrides))
    assert (p2.nx, p2.nz) == (overrides["nx"], overrides["nz"]) and p2.float is np.float64
    # every file the reference run produced exists in ours, with the same content
    compared = 0
    for path in sorted(glob.glob(os.path.join(ref, "*.np[yz]"))):
        fname = os.path.basename(path)
        mine = os.path.join(ours, fname)
        assert os.path.exists(mine), f"{fname} missing from the drop-in run"
        if fname.endswith(".npy"):
            a, b = np.load(mine), np.load(path)
            assert a.shape == b.shape and rel(a, b) < 1e-10, fname
            compared += 1
            continue
        za, zb = np.load(mine, allow_pickle=True), np.load(path, allow_pickle=True)
        for key in zb.files:
            if key == "params" or zb[key].dtype == object:
                continue
            a, b = np.asarray(za[key]), np.asarray(zb[key])
            assert a.shape == b.shape, (fname, key)
            if b.size and np.issubdtype(b.dtype, np.number):
                tol = 1e-9 if fname.startswith(("kinetic", "nusselt")) else 1e-10
                # (Taylor-Green: the nonlinear term vanishes analytically, dw is rounding noise)
                assert rel(a, b) < tol or np.max(np.abs(a - b)) < 1e-15, (fname, key, rel(a, b))
                compared += 1
    assert compared >= 2
