"""bench.py's workload builders (the example scripts' loop bodies at the BASELINE parameters)
stepped at tiny grids on the emulation build: the benchmark cannot be run in the CPU
container, its host logic can."""
import os
import shutil
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "melvin.py_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

if shutil.which("g++") is None:  # pragma: no cover
    pytest.skip("g++ not available for the emulation build", allow_module_level=True)

import emu_harness as eh  # noqa: E402
from melvin import _backend  # noqa: E402
from test_emu_python_api import emu_backend  # noqa: E402,F401  (module-scoped autouse fixture)


@pytest.mark.parametrize("config,nx,nz", [("kh", 64, 64), ("tearing", 64, 32), ("ddc", 32, 64), ("rbc", 64, 24)])
def test_bench_public_loops_step(config, nx, nz):
    import bench
    cwd = os.getcwd()
    try:
        step, o = bench.build_public_loop(config, nx, nz)
        for _ in range(12):                      # crosses a CFL tick (loop 11) and the first tracker tick
            step()
        assert o["sim"]._loop_counter == 12
        assert np.all(np.isfinite(o["w"].on_host()))
        assert len(o["sim"]._trackers[0]._values) == 1 and np.isfinite(float(o["sim"]._trackers[0]._values[0]))
    finally:
        os.chdir(cwd)


def test_bench_parity_preflight_single():
    import bench
    res = bench.parity_single_gpu()
    assert res["ok"] and res["cases"][0]["field_rel_l2"] < 1e-12


def test_streamed_ensemble_matches_blocking_loop():
    """melvin/ensemble.py: members stepped round-robin with host-resident states (Variable.load from
    a host buffer, on_host(out=...)) give the states of the same simulations stepped one by one."""
    import bench
    from melvin.ensemble import Ensemble
    cwd = os.getcwd()
    try:
        def build(i):
            step, o = bench.build_public_loop("kh", 32, 32)
            return {"step": step, "o": o, "host": o["w"].on_host().copy()}
        ens = Ensemble(build, members=3)
        assert len({id(m.payload["o"]["w"]._ctx) for m in ens.members}) == (3 if _backend.is_cuda() else 1)
        for k in range(9):
            with ens.turn(k) as m:
                p = m.payload
                p["o"]["w"].load(p["host"], is_physical=False)
                p["step"]()
                p["o"]["w"].on_host(out=p["host"])
        ens.drain()
        step, o = bench.build_public_loop("kh", 32, 32)
        for _ in range(3):
            step()
        want = o["w"].on_host()
        for m in ens.members:
            assert m.passes == 3
            np.testing.assert_allclose(m.payload["host"], want, rtol=0, atol=1e-14 * np.abs(want).max())
        with pytest.raises(ValueError):
            o["w"].on_host(out=np.zeros((3, 3), dtype=np.complex128))
    finally:
        os.chdir(cwd)
