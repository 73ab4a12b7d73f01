"""Every public class / method / utility function of the reference package exists under the same
name on the drop-in package.  Introspects the reference checkout in a subprocess
(both packages are called `melvin`), so it is skipped where /root/reference is absent."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MELVIN_REFERENCE", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "melvin")),
                                reason="needs the reference checkout (/root/reference)")

PROBE = r"""
import importlib, inspect, json, pkgutil, sys
sys.path.insert(0, sys.argv[1])
import melvin
surface = {"__all__": sorted(n for n in dir(melvin) if not n.startswith("_"))}
for info in pkgutil.iter_modules(melvin.__path__):
    mod = importlib.import_module("melvin." + info.name)
    for name, obj in vars(mod).items():
        if name.startswith("_") or getattr(obj, "__module__", None) != mod.__name__:
            continue
        if inspect.isclass(obj):
            surface[name] = sorted(n for n, _ in inspect.getmembers(obj, inspect.isfunction)
                                   if not n.startswith("_") or n in ("__getitem__", "__setitem__"))
        elif inspect.isfunction(obj) and info.name == "utility":
            surface.setdefault("utility", []).append(name)
print(json.dumps(surface))
"""


def probe(path):
    out = subprocess.run([sys.executable, "-c", PROBE, path], capture_output=True, text=True, check=True,
                         env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    return json.loads(out.stdout.splitlines()[-1])


def test_public_surface_of_the_reference_is_mirrored():
    ref = probe(REF)
    sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))
    import melvin
    from melvin import fields, operators, simulation, utility
    homes = (melvin, simulation, operators, fields)
    missing = []
    for name in ref.pop("__all__"):
        if name[0].isupper() and not hasattr(melvin, name):
            missing.append(name)
    for fn in ref.pop("utility", []):
        if not hasattr(utility, fn):
            missing.append("utility." + fn)
    assert len(ref) >= 10                                   # the probe saw the reference's classes
    for cls, methods in ref.items():
        home = next((getattr(h, cls) for h in homes if hasattr(h, cls)), None)
        if home is None:
            missing.append(cls)
            continue
        missing += [f"{cls}.{m}" for m in methods if not hasattr(home, m)]
    assert not missing, missing
