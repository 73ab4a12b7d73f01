"""Slab-decomposed step on real GPUs over NCCL (needs >= 2 devices; the CPU/gloo
version of the same worker runs in tests/test_sharded_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("mode", ["p2p", "dma", "a2a"])
@pytest.mark.parametrize("case", ["tg64", "kh", "khlong"])
def test_sharded_step_nccl(case, mode):
    """all exchange modes: stores into peer memory fused in the producer kernels, copy-engine
    transfers into peer memory, asynchronous NCCL all-to-all per field"""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29701",
           os.path.join(ROOT, "tests", "sharded_worker.py"), "cuda", case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, MLV_EXCHANGE=mode))
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("SHARDED")]
    assert out.returncode == 0 and lines and lines[-1].endswith("OK"), out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("mode", ["p2p", "dma", "a2a"])
@pytest.mark.parametrize("case", ["ddc", "tearing"])
def test_multi_field_steppers_nccl(case, mode):
    """double-diffusive and MHD tearing loops slab-decomposed over 2 GPUs, every exchange mode,
    vs the goldens of the unmodified reference"""
    if _ngpu() < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29702",
           os.path.join(ROOT, "tests", "sharded_worker.py"), "cuda", case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, MLV_EXCHANGE=mode))
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("SHARDED")]
    assert out.returncode == 0 and lines and lines[-1].endswith("OK"), out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("case", ["api_tg", "api_ddc", "api_tearing"])
def test_public_api_shards_itself_nccl(case):
    """N4: the unchanged public-API loops under torchrun on GPUs (slabs + NCCL all-to-all behind
    Simulation / Variable), vs the goldens of the unmodified reference"""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29703",
           os.path.join(ROOT, "tests", "sharded_worker.py"), "cuda", case]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("SHARDED")]
    assert out.returncode == 0 and lines and lines[-1].endswith("OK"), out.stdout[-2000:] + out.stderr[-2000:]


def test_single_rank_stepper_matches_python_api_path():
    """The rank-local stepper (world 1) against the golden: same kernels as the public API."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "melvin.py_b200"))
    from conftest import golden, rel_l2
    from melvin.sharded import ShardedScalarStepper
    from oracle import melvin_oracle as mo
    gl = golden("loop_tg_64x64.npz")
    g = mo.Grid(64, 64, float(gl["lx"]), float(gl["lz"]))
    st = ShardedScalarStepper(64, 64, g.lx, g.lz, float(gl["coef"]), float(gl["dt"]), tracker_cadence=1)
    st.load_spectral(mo.to_spectral(g, mo.ic_taylor_green(g)))
    for k in range(1, 21):
        st.step()
        if k in (1, 10, 20):
            assert rel_l2(st.gather_spectral(), gl[f"w_step{k}"]) < 1e-12
    np.testing.assert_allclose(st.ke, gl["ke"], rtol=1e-9)
