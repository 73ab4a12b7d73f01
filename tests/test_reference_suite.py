"""The reference's OWN test suite (/root/reference/test/*_test.py: transforms on all basis pairs,
array shapes, the FDM Laplacian solve, Variable derivatives), run UNMODIFIED against the drop-in
package.  Its fixtures pass `np` as the array namespace (test/conftest.py:38-49), so the run sets
MELVIN_B200_NUMPY_IS_DEVICE=1; kernels execute on the host emulation build here (the GPU box has no
reference checkout)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MELVIN_REFERENCE", "/root/reference")

pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF, "test")) or shutil.which("g++") is None,
    reason="needs the reference checkout (/root/reference) and g++ for the emulation build")


def test_reference_test_suite_passes_on_the_drop_in(tmp_path):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", MELVIN_B200_NUMPY_IS_DEVICE="1",
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), os.path.join(ROOT, "melvin.py_b200")]))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REF, "test"), "-p", "reference_suite_plugin",
                        "-p", "no:cacheprovider", "-q", "-W", "ignore::UserWarning"],
                       cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    tail = r.stdout[-2000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    ntests = sum(len(re.findall(r"^def test_", open(os.path.join(REF, "test", f)).read(), re.M))
                 for f in os.listdir(os.path.join(REF, "test")) if f.endswith("_test.py"))
    assert m and int(m.group(1)) == ntests and "failed" not in r.stdout and "skipped" not in r.stdout, tail
    # ... and it really was the drop-in: the same command without the switch is refused (no CPU path)
    env.pop("MELVIN_B200_NUMPY_IS_DEVICE")
    r2 = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REF, "test", "ArrayFactory_test.py"), "-p",
                         "reference_suite_plugin", "-p", "no:cacheprovider", "-q"],
                        cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    assert r2.returncode != 0 and "BackendUnavailable" in r2.stdout + r2.stderr
