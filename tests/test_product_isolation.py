"""The oracle is test infrastructure: nothing in the product package may import, call or
execute it, and the product has no CPU fallback (it must fail loudly without the CUDA library)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "melvin.py_b200")


def _product_sources():
    for base, _dirs, files in os.walk(PKG):
        if "_lib" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(base, f)


def test_product_never_touches_the_oracle():
    bad = []
    for path in _product_sources():
        with open(path, encoding="utf-8") as fp:
            for n, line in enumerate(fp, 1):
                if re.search(r"^\s*(from|import)\s+oracle\b|melvin_oracle|oracle/", line):
                    bad.append(f"{os.path.relpath(path, ROOT)}:{n}: {line.strip()}")
    assert not bad, "\n".join(bad)


def test_no_numpy_compute_fallback_in_backend():
    """_backend raises instead of computing on the host when the library or a GPU is missing."""
    with open(os.path.join(PKG, "melvin", "_backend.py"), encoding="utf-8") as fp:
        src = fp.read()
    assert "BackendUnavailable" in src and "no CPU fallback" in src
    assert "numpy.fft" not in src and "np.fft" not in src
