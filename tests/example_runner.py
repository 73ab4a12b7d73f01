#!/usr/bin/env python3
"""Runs one of the reference's example scripts UNMODIFIED (runpy), either on the drop-in
package or on the reference itself, with run parameters overridden through a patched
``Parameters`` class (the script file is never edited):

    example_runner.py --impl emu|gpu|reference --script /root/reference/examples/X.py \
        --overrides '{"nx": 64, "nz": 64, "tracker_cadence": 1}' --steps 10 --out DIR

impl = emu / gpu : `melvin` is melvin.py_b200/melvin and `import cupy` resolves to
                   melvin.py_b200/shims/cupy (emu: kernels on the host emulation build, the
                   CPU development harness; gpu: libmelvin_b200.so on cuda:0).
impl = reference : `melvin` is the unmodified reference checkout and `import cupy` resolves to
                   NumPy (xp.__name__ == "numpy" selects the reference's CPU path), precision
                   forced to "double" -- the path the drop-in promotes "single" to.
Outputs (kinetic_energy.npz, w0000.npy, dump*.npz, params.json ...) land in DIR.
"""
import argparse
import contextlib
import io
import json
import os
import runpy
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def patch(base, overrides, steps):
    """Subclass of the package's Parameters that applies `overrides` to the script's dict."""
    global Parameters

    def __init__(self, params, validate=True):
        params = dict(params)
        params.update(overrides)
        if steps:
            params["final_time"] = (steps - 0.5) * params["initial_dt"]
        base.__init__(self, params, validate)

    Parameters = type("Parameters", (base,), {"__init__": __init__, "__module__": __name__})
    return Parameters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["emu", "gpu", "reference"])
    ap.add_argument("--script", required=True)
    ap.add_argument("--overrides", default="{}")
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    overrides = json.loads(args.overrides)

    if args.impl == "reference":
        import numpy
        sys.modules["cupy"] = numpy                      # xp = cupy -> the NumPy module
        sys.path.insert(0, args.reference)
        overrides = dict(overrides, precision="double")
    else:
        sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200", "shims"))
        sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))
        if args.impl == "emu":
            sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
            import emu_harness
            from melvin import _backend
            _backend._install(emu_harness.lib(), "cpu")
    import melvin
    import example_runner as er        # importable home of the patched class (dumps pickle it)
    melvin.Parameters = er.patch(melvin.Parameters, overrides, args.steps)
    os.makedirs(args.out, exist_ok=True)
    os.chdir(args.out)
    log = io.StringIO()
    with contextlib.redirect_stdout(log), warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        runpy.run_path(args.script, run_name="__main__")
    with open("stdout.txt", "w") as fp:
        fp.write(log.getvalue())
    with open("warnings.txt", "w") as fp:
        fp.write("\n".join(str(w.message) for w in caught))
    if args.impl != "reference":
        from melvin import _backend
        with open("launches.txt", "w") as fp:
            fp.write(str(_backend.launches()))


if __name__ == "__main__":
    main()
