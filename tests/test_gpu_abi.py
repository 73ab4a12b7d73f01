"""Parity tests of the CUDA kernels through the C ABI (device run of abi_cases.py),
plus full-size checks on the BASELINE grid (4096^2) against the oracle and through
size-independent properties."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import abi_cases as ac  # noqa: E402
from conftest import rel_l2  # noqa: E402
from oracle import melvin_oracle as mo  # noqa: E402


@pytest.fixture(scope="module")
def H():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    import gpu_harness
    return gpu_harness


@pytest.mark.parametrize("nx,nz", ac.SIZES_2D + [(1024, 1024)])
def test_transforms_2d(H, nx, nz):
    ac.case_transforms_2d(H, nx, nz)


@pytest.mark.parametrize("nx,nz", ac.SIZES_1D + [(2048, 512)])
def test_transforms_1d_fdm(H, nx, nz):
    ac.case_transforms_1d_fdm(H, nx, nz)


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("nx,nz", ac.SIZES_FUSED + [(1024, 512)])
def test_fused_advection_step(H, nx, nz, order):
    ac.case_fused_advection_step(H, nx, nz, order)


@pytest.mark.parametrize("nx,nz,grid", [(64, 512, 3), (32, 2048, 5), (16, 4096, 0), (16, 4096, 3), (1024, 4096, 0),
                                        (2048, 1024, 0)])
def test_fused_advection_persistent_grouped(H, nx, nz, grid, monkeypatch):
    """three-pass line lengths: persistent CTAs (several row pairs per CTA, also forced on small
    grids) with grouped-order transforms, and the classic one-pair-per-CTA kernel, vs the oracle"""
    monkeypatch.setenv("MLV_ZADV_GROUPED", "1")
    if grid:
        monkeypatch.setenv("MLV_ZADV_GRID", str(grid))
    ac.case_fused_advection_step(H, nx, nz, 2)
    monkeypatch.delenv("MLV_ZADV_GROUPED")
    ac.case_fused_advection_step(H, nx, nz, 2)


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("nx,nz,bits", ac.SIZES_SPLIT)
def test_split_lines_forced(H, nx, nz, bits, order):
    """long-line kernels forced onto small grids (same cases as the emulation build)"""
    ac.case_split_lines(H, nx, nz, bits, order)


@pytest.mark.parametrize("nx,nz", [(16384, 32), (32, 16384)])
def test_lines_of_16384_points(H, nx, nz):
    """BASELINE config-5 line length: transforms and the fused advection step vs the oracle"""
    ac.case_transforms_2d(H, nx, nz)
    ac.case_fused_advection_step(H, nx, nz, 2)


def test_pointwise_and_stencils(H):
    ac.case_pointwise_and_stencils(H)


def test_fdm_solver_and_stencils(H):
    ac.case_fdm_solver_and_stencils(H)


@pytest.mark.parametrize("nx,nz,order", [(64, 13, 2), (64, 40, 4), (32, 300, 4), (16, 2048, 2), (16, 2500, 4),
                                         (4096, 2048, 4)])
def test_fdm_fused_step(H, nx, nz, order):
    """Fourier-x / FDM-z kernels (batched scan solver + velocities, fused 1-D advection, row-wise
    right-hand side and update) vs the oracle, incl. the BASELINE config-3 grid 4096 x 2048 and the
    F8 gate of the solver against an extended-precision solution and the reference's SuperLU."""
    ac.case_fdm_fused_step(H, nx, nz, order)


def test_integrate_and_array_ops(H):
    ac.case_integrate_and_array_ops(H)


def test_full_size_4096_transforms_and_properties(H):
    """BASELINE config-2 grid: transform parity vs the oracle (pocketfft), round trip
    (idempotence on the retained band) and linearity."""
    nx = nz = 4096
    g = mo.Grid(nx, nz, 16.0 / 9.0, 1.0)
    ctx = H.Ctx(nx, nz, g.lx, g.lz)
    rng = np.random.default_rng(42)
    phys = rng.standard_normal((nx, nz))
    spec = np.zeros(g.spectral_shape, complex)
    I = ctx.ibuf()
    ctx.call("mlv_to_spectral", H.ptr(phys), H.ptr(I), H.ptr(spec))
    assert rel_l2(spec, mo.to_spectral(g, phys)) < 1e-13
    back = np.zeros((nx, nz))
    ctx.call("mlv_to_physical", H.ptr(spec), H.ptr(I), H.ptr(back))
    assert rel_l2(back, mo.to_physical(g, spec)) < 1e-13
    spec2 = np.zeros_like(spec)
    ctx.call("mlv_to_spectral", H.ptr(back), H.ptr(I), H.ptr(spec2))
    assert rel_l2(spec2, spec) < 1e-13                       # band-limited round trip
    other = rng.standard_normal((nx, nz))
    so = np.zeros_like(spec)
    ctx.call("mlv_to_spectral", H.ptr(other), H.ptr(I), H.ptr(so))
    comb = np.zeros_like(spec)
    mix = 0.3 * phys - 1.7 * other
    ctx.call("mlv_to_spectral", H.ptr(mix), H.ptr(I), H.ptr(comb))
    assert rel_l2(comb, 0.3 * spec - 1.7 * so) < 1e-13       # linearity
    ctx.close()


@pytest.mark.parametrize("nx,nz,grid", [(4096, 16, 0), (4096, 64, 3), (4096, 1024, 0), (4096, 4096, 0)])
def test_inverse_x_pass_column_serial(H, nx, nz, grid, monkeypatch):
    """4096-point x lines: column-serial persistent inverse x pass (two 256-thread CTAs per SM,
    16-byte-wide tensor loads / stores, swizzled exchange buffer), several columns per CTA when the
    grid is forced small, and the classic two-column kernel, vs the oracle"""
    if grid:
        monkeypatch.setenv("MLV_XINV_GRID", str(grid))
    ac.case_transforms_2d(H, nx, nz)
    if nz <= 1024:
        ac.case_fused_advection_step(H, nx, nz, 2)
    monkeypatch.setenv("MLV_XINV_COLS", "1")
    ac.case_transforms_2d(H, nx, nz)
    if nz <= 1024:
        ac.case_fused_advection_step(H, nx, nz, 2)
