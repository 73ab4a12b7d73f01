"""Host logic of the slab-decomposed (multi-GPU) step on CPU: world_size 2 and 4 over
gloo, kernels from the emulation build (development harness).  The same worker runs
with NCCL on GPUs in tests/test_gpu_sharded.py."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if shutil.which("g++") is None:  # pragma: no cover
    pytest.skip("g++ not available for the emulation build", allow_module_level=True)


def run_world(n, case, port, **extra_env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_worker.py"), "emu", case]
    env = dict(os.environ, OMP_NUM_THREADS="1", **extra_env)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("SHARDED")]
    assert out.returncode == 0 and lines and lines[-1].endswith("OK"), out.stdout[-2000:] + out.stderr[-2000:]


@pytest.fixture(scope="module", autouse=True)
def emulation_build():
    """build the emulation library ONCE, here, if it is stale (the ranks of a world must not race for it)"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import emu_harness
    emu_harness.lib()


@pytest.mark.parametrize("world", [2, 4])          # (world 1: test_multi_field_steppers[ddc-1])
def test_taylor_green_sharded_matches_reference_golden(world):
    run_world(world, "tg64", 29600 + world)


def test_kelvin_helmholtz_sharded_uneven_split():
    run_world(2, "kh", 29611)


def test_long_line_kernels_sharded():
    """split x passes + real-row z stage under the slab decomposition (world 2)"""
    run_world(2, "khlong", 29612)


@pytest.mark.parametrize("case", ["khlong", "tearing"])
def test_forward_exchange_in_row_blocks(case):
    """forward buffers cut into row blocks, z stage launched block by block (the layout the
    copy-engine exchange pipelines on GPUs)"""
    run_world(2, case, 29613, MLV_FWD_CHUNKS="4")


@pytest.mark.parametrize("case,world", [("ddc", 1), ("ddc", 2), ("tearing", 2), ("tearing", 4)])
def test_multi_field_steppers(case, world):
    """double-diffusive (3 coupled scalars) and MHD tearing (2 exchange rounds per step)
    slab-decomposed, vs the goldens of the unmodified reference"""
    run_world(world, case, 29620 + world)


def test_bench_multi_gpu_host_logic():
    """bench.py's sharded parity preflight and workload builders over gloo (world 2)"""
    run_world(2, "bench", 29631)


@pytest.mark.parametrize("case,world", [("api_tg", 4), ("api_ddc", 2), ("api_tearing", 2)])
def test_public_api_shards_itself(case, world):
    """N4: the UNCHANGED loops of tests/parity_cases.py (Simulation / Variable / Integrator API)
    under a process group: the package decomposes the fields into slabs itself (melvin/_dist.py),
    results gathered on read; vs the goldens of the unmodified reference"""
    run_world(world, case, 29640 + world)
