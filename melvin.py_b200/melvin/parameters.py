"""Run parameters: a plain dict turned into attributes plus derived sizes.

API mirror of the reference's ``melvin/Parameters.py`` (defaults :9-23, derived
values :63-92, ``save`` :94-96).  Host-only glue; the derived ``nn``/``nm``/
shapes define every device layout.
"""
import json
import os
import sys
import warnings

import numpy as np

_DEFAULTS = dict(
    integrator_order=2,
    integrator="semi-implicit",
    spatial_derivative_order=2,
    alpha=0.51,
    cfl_cutoff=0.5,
    cfl_cadence=10,        # loops between CFL checks
    tracker_cadence=100,   # loops between scalar-tracker samples
    save_cadence=100,      # simulated time between saves
    load_from=None,
    discretisation=["spectral", "spectral"],
    precision="double",
    laplacian_order=2,     # EXTENSION (not a reference key): 4 = pentadiagonal FDM-z Laplacian solve
    nx=None, nz=None, lx=None, lz=None, final_time=None,
)


class Parameters:
    required_params = ["nx", "nz", "lx", "lz", "final_time"]

    def __init__(self, params, validate=True):
        for key, val in _DEFAULTS.items():
            setattr(self, key, list(val) if isinstance(val, list) else val)
        if validate and not self.is_valid(params):
            sys.exit(-1)
        self.load_from_dict(params)
        self._original_params = params

    def load_from_dict(self, params):
        for key, val in params.items():
            setattr(self, key, val)
        self.set_derived_params(params)

    def is_valid(self, params):
        ok = True
        for key in self.required_params:
            if key not in params:
                print(key, "missing from input parameters.")
                ok = False
        # the reference inspects the class defaults here, before the dict is applied
        # (SURVEY App. A-14), so this check never fires; kept for identical behaviour
        if "fdm" in self.discretisation and self.integrator == "explicit":
            print("FDM and implicit method currently not supported.")
            ok = False
        return int(ok)

    def is_fully_spectral(self):
        return all(d == "spectral" for d in self.discretisation[:2])

    def set_derived_params(self, params):
        if "dump_cadence" not in params:
            self.dump_cadence = 0.1 * self.final_time
        disc = self.discretisation
        if disc[0] == "spectral":
            self.nn = (self.nx - 1) // 3
        if disc[1] == "spectral":
            self.nm = (self.nz - 1) // 3
        if self.is_fully_spectral():
            self.spectral_shape = (2 * self.nn + 1, self.nm)
        elif disc[0] == "fdm":
            self.spectral_shape = (self.nx, self.nm)
        elif disc[1] == "fdm":
            self.spectral_shape = (self.nn, self.nz)
        self.physical_shape = (self.nx, self.nz)
        self.dx = self.lx / self.nx
        self.dz = self.lz / self.nz
        if self.precision == "double":
            self.complex, self.float = np.complex128, np.float64
        elif self.precision == "single":
            # Every reference example asks for "single" (examples/taylor_green_vortex.py:40).
            # The B200 kernels are float64/complex128 only (BASELINE: the fp64 path is the
            # one that is parity-checked), so the request is *promoted*: same API, results at
            # least as accurate, arrays come back as float64.  MLV_STRICT_PRECISION=1 refuses.
            if os.environ.get("MLV_STRICT_PRECISION"):
                raise NotImplementedError(
                    "melvin-b200 computes in float64/complex128 only (precision='double')")
            warnings.warn("melvin-b200: precision='single' is computed in float64/complex128 "
                          "(MLV_STRICT_PRECISION=1 to refuse)", stacklevel=3)
            self.complex, self.float = np.complex128, np.float64
        if "initial_dt" not in params:
            self.initial_dt = 0.2 * min(self.dx, self.dz)

    def save(self, fname="params.json"):
        with open(fname, "w") as fp:
            json.dump(self._original_params, fp)
