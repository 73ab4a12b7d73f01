"""pytest plugin for tests/test_reference_suite.py: makes `melvin` the drop-in package on the host
emulation build of the kernels (the CPU development harness) before the reference's own conftest
imports it.  Not product code."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import emu_harness  # noqa: E402
from melvin import _backend  # noqa: E402

_backend._install(emu_harness.lib(), "cpu")
