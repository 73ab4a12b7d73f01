"""Basis-function tags and their derivative factors.

API mirror of the reference's ``melvin/BasisFunctions.py`` (enum :5-9, factors
:26-59).  Only COMPLEX_EXP (Fourier) and FDM are computed by the CUDA path;
COSINE / SINE are accepted as tags but their transforms are "next tier"
 and raise NotImplementedError when used.
"""
from enum import IntEnum

import numpy as np


class BasisFunctions(IntEnum):
    COMPLEX_EXP = 0
    COSINE = 1
    SINE = 2
    FDM = 3


_SPECTRAL = (BasisFunctions.COMPLEX_EXP, BasisFunctions.COSINE, BasisFunctions.SINE)


def is_spectral(bs):
    return any(bs is s for s in _SPECTRAL)


def is_fully_spectral(bs1, bs2):
    return is_spectral(bs1) and is_spectral(bs2)


_WAVELENGTH = {
    BasisFunctions.COMPLEX_EXP: 1j * 2 * np.pi,
    BasisFunctions.SINE: np.pi,
    BasisFunctions.COSINE: -np.pi,
}


def calc_diff_wavelength(basis_fn):
    for key, val in _WAVELENGTH.items():
        if basis_fn is key:
            return val
    return 0


def calc_diff2_wavelength(basis_fn):
    return -np.abs(calc_diff_wavelength(basis_fn)) ** 2


def calc_diff_factor(basis_fn, length):
    return calc_diff_wavelength(basis_fn) / length


def gen_diff_factors(length):
    return [calc_diff_wavelength(b) / length for b in BasisFunctions]


def gen_diff2_factors(length):
    return [calc_diff2_wavelength(b) / length ** 2 for b in BasisFunctions]
