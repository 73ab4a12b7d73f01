"""pytest configuration: markers and import paths.

``-m "not gpu"`` : oracle vs golden fixtures, host logic, C-ABI symbol check
                   (runs in the CPU-only dev container in a few minutes).
``-m gpu``       : parity tests proper -- CUDA path through the C ABI vs the
                   oracle and the committed goldens (run on a B200).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "melvin.py_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def golden_loader():
    return golden


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    return num / den if den > 0 else num
