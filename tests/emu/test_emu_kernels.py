"""Kernel index-logic tests on the CPU emulation build (development harness).

These run the *same kernel sources* as the GPU build, compiled for the host
(threads of a CTA = fibers), through the same C ABI, and compare with the
oracle.  They complement -- never replace -- the ``-m gpu`` parity tests
(tests/test_gpu_abi.py runs the identical cases on the device).
"""
import os
import shutil
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "melvin.py_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

if shutil.which("g++") is None:  # pragma: no cover
    pytest.skip("g++ not available for the emulation build", allow_module_level=True)

import abi_cases as ac  # noqa: E402
import emu_harness as H  # noqa: E402


@pytest.mark.parametrize("nx,nz", ac.SIZES_2D)
def test_transforms_2d(nx, nz):
    ac.case_transforms_2d(H, nx, nz)


@pytest.mark.parametrize("nx,nz", ac.SIZES_1D)
def test_transforms_1d_fdm(nx, nz):
    ac.case_transforms_1d_fdm(H, nx, nz)


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("nx,nz", ac.SIZES_FUSED)
def test_fused_advection_step(nx, nz, order):
    ac.case_fused_advection_step(H, nx, nz, order)


@pytest.mark.parametrize("nx,nz,grid", [(64, 512, 3), (32, 2048, 5), (16, 4096, 0), (16, 4096, 3)])
def test_fused_advection_persistent_grouped(nx, nz, grid, monkeypatch):
    """three-pass line lengths: persistent CTAs over several row pairs (grid forced small) with
    grouped-order transforms, incl. a trip count that leaves lines of the last trip idle"""
    monkeypatch.setenv("MLV_ZADV_GROUPED", "1")
    if grid:
        monkeypatch.setenv("MLV_ZADV_GRID", str(grid))
    ac.case_fused_advection_step(H, nx, nz, 2)
    monkeypatch.delenv("MLV_ZADV_GROUPED")
    ac.case_fused_advection_step(H, nx, nz, 2)


@pytest.mark.parametrize("order", [2, 4])
@pytest.mark.parametrize("nx,nz,bits", ac.SIZES_SPLIT)
def test_split_lines(nx, nz, bits, order):
    ac.case_split_lines(H, nx, nz, bits, order)


def test_pointwise_and_stencils():
    ac.case_pointwise_and_stencils(H)


def test_fdm_solver_and_stencils():
    ac.case_fdm_solver_and_stencils(H)


def test_integrate_and_array_ops():
    ac.case_integrate_and_array_ops(H)


@pytest.mark.parametrize("nx,nz,order", [(64, 13, 2), (64, 40, 4), (32, 300, 4), (16, 2048, 2), (16, 2500, 4)])
def test_fdm_fused_step(nx, nz, order):
    """batched scan solver (one and several unknowns per thread, 256 and 512 threads), fused
    1-D advection and the row-wise right-hand side / update"""
    ac.case_fdm_fused_step(H, nx, nz, order)


@pytest.mark.parametrize("nx,nz,grid", [(4096, 16, 0), (4096, 64, 3)])
def test_inverse_x_pass_column_serial(nx, nz, grid, monkeypatch):
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
