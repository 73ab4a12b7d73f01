"""melvin -- drop-in, B200-native implementation of Melvin.py's per-timestep
pseudo-spectral hot path.

Same public names as the reference package (reference ``melvin/__init__.py``);
the array namespace to pass as ``xp`` is ``melvin.b200`` (or the ``cupy`` shim in
``melvin.py_b200/shims`` so that the reference's example scripts run unchanged).
"""
from .basis import BasisFunctions
from .parameters import Parameters
from .operators import (ArrayFactory, Integrator, LaplacianSolver, SpatialDifferentiator,
                        SpectralTransformer)
from .fields import TimeDerivative, Variable
from .simulation import DataTransferer, ScalarTracker, Simulation, Ticker, Timer
from .utility import load_scipy_sparse, load_scipy_sparse_linalg
from . import b200, utility

__all__ = [
    "ArrayFactory", "BasisFunctions", "DataTransferer", "Integrator", "LaplacianSolver",
    "Parameters", "ScalarTracker", "SpatialDifferentiator", "SpectralTransformer", "Ticker",
    "TimeDerivative", "Timer", "Variable", "Simulation", "b200", "utility",
    "load_scipy_sparse", "load_scipy_sparse_linalg",
]
