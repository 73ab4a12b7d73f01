"""Simulation bookkeeping: object factory, tickers, trackers, save / dump.

API mirror of the reference's ``melvin/Simulation.py``, ``Ticker.py``,
``ScalarTracker.py``, ``Timer.py`` and ``DataTransferer.py`` (host-side control;
SURVEY section 2 marks these out of the compute scope, but scripts rely on their
exact semantics).
"""
import time

import numpy as np

from . import _backend
from .b200 import DeviceArray
from .fields import TimeDerivative, Variable
from .operators import (ArrayFactory, Integrator, LaplacianSolver, SpatialDifferentiator,
                        SpectralTransformer, _require_device_namespace)


class Timer:
    """Keeps track of current runtime (reference melvin/Timer.py)."""

    def __init__(self):
        self._start_time = 0.0
        self.diff = 0.0

    def start(self):
        self._start_time = time.time()

    def split(self):
        now = time.time()
        self.diff = now - self._start_time
        self._start_time = now


class Ticker:
    """Fires `fn` whenever its counter falls behind the loop counter (or the
    simulated time); fires on the very first loop (reference melvin/Ticker.py:14-25)."""

    def __init__(self, cadence, fn, dump_name="", is_loop_counter=True):
        self._cadence = cadence
        self._fn = fn
        self._counter = 0
        self.times_fired = 0
        self._dump_name = dump_name
        self._by_loop = is_loop_counter

    def tick(self, t, loop_counter):
        if self._counter < (loop_counter if self._by_loop else t):
            self._fn(self)
            self._counter += self._cadence
            self.times_fired += 1

    def due(self, t, loop_counter):
        return self._counter < (loop_counter if self._by_loop else t)

    def dump(self):
        return {"counter": self._counter, "times_fired": self.times_fired}

    def restore(self, data):
        data = data.item() if hasattr(data, "item") and not isinstance(data, dict) else data
        self._counter = data["counter"]
        self.times_fired = data["times_fired"]

    def get_name(self):
        return self._dump_name


class ScalarTracker:
    """Tracks a scalar through time (reference melvin/ScalarTracker.py)."""

    def __init__(self, params, xp, filename):
        self._xp = _require_device_namespace(xp)
        self._params = params
        self._filename = filename
        self._values = []
        self._times = []

    def append(self, t, val):
        self._times += [t]
        self._values += [val]

    def save(self, _index):
        from . import _dist
        if _dist.rank() != 0:
            return
        np.savez(self._filename, t=np.array(self._times, dtype=np.float64),
                 values=np.array([float(v) for v in self._values]))


class DataTransferer:
    """Host <-> device copies (reference melvin/DataTransferer.py)."""

    def __init__(self, xp):
        self._xp = _require_device_namespace(xp)

    def to_host(self, data):
        return data.get() if hasattr(data, "get") else np.asarray(data)

    def from_host(self, data):
        if isinstance(data, DeviceArray):
            return data
        return DeviceArray(_backend.from_host(np.asarray(data)))


class Simulation:
    """Factory + bookkeeping (reference melvin/Simulation.py:18-221)."""

    def __init__(self, params, xp):
        self._params = params
        xp = _require_device_namespace(xp)
        self._xp = xp
        self._data_trans = DataTransferer(xp)
        self._integrator = Integrator(params, xp)
        self._array_factory = ArrayFactory(params, xp)
        self._spatial_diff = SpatialDifferentiator(params, xp, self._array_factory)
        self._spectral_trans = SpectralTransformer(params, xp, self._array_factory)
        self._t = 0
        self._loop_counter = 0
        self.reductions = "auto"             # see end_loop
        self._ctx = _backend.context_for(params)
        self._dump_vars = []
        self._dump_dvars = []
        self._dump_idx = 0
        self._save_vars = []
        self._tickers = []
        self._timer = Timer()
        self._wallclock_remaining = 0.0
        self._wallclock_ticker = Ticker(100, self.calc_time_remaining, is_loop_counter=True)
        self.register_ticker(self._wallclock_ticker)

    # -- factories
    def make_variable(self, name, basis_fns):
        return Variable(self._params, self._xp, sd=self._spatial_diff, st=self._spectral_trans,
                        dt=self._data_trans, array_factory=self._array_factory,
                        dump_name=name, basis_functions=basis_fns)

    def make_derivative(self, name):
        return TimeDerivative(self._params, self._xp, dump_name=name)

    def init_laplacian_solver(self, basis_fns):
        self._laplacian_solver = LaplacianSolver(
            self._params, self._xp, basis_fns, spatial_diff=self._spatial_diff,
            array_factory=self._array_factory)

    def get_laplacian_solver(self):
        return self._laplacian_solver

    # -- ticker configuration
    def config_dump(self, variables, derivatives):
        self._dump_vars = variables
        self._dump_dvars = derivatives
        self._dump_ticker = Ticker(self._params.dump_cadence, self.dump,
                                   dump_name="dump_ticker", is_loop_counter=False)
        self.register_ticker(self._dump_ticker)

    def config_save(self, variables):
        self._save_vars = variables
        self._save_ticker = Ticker(self._params.save_cadence, self.save,
                                   dump_name="save_ticker", is_loop_counter=False)
        self.register_ticker(self._save_ticker)

    def config_cfl(self, ux, uz):
        self._ux = ux
        self._uz = uz
        self._cfl_ticker = Ticker(self._params.cfl_cadence, self.set_dt,
                                  dump_name="cfl_ticker", is_loop_counter=True)
        self.register_ticker(self._cfl_ticker)

    def config_scalar_trackers(self, trackers):
        self._tracker_fns = [trackers[fname] for fname in trackers]
        self._trackers = [ScalarTracker(self._params, self._xp, fname) for fname in trackers]
        self._tracker_ticker = Ticker(self._params.tracker_cadence, self.track_scalars,
                                      dump_name="tracker_ticker", is_loop_counter=True)
        self.register_ticker(self._tracker_ticker)
        self._save_vars += self._trackers

    # -- ticker callbacks
    def track_scalars(self, ticker):
        for tracker, func in zip(self._trackers, self._tracker_fns):
            tracker.append(self._t, func())

    def set_dt(self, ticker):
        self._integrator.set_dt(self._ux, self._uz)

    def save(self, ticker):
        self.print_info()
        for var in self._save_vars:
            var.save(ticker.times_fired)

    def print_info(self):
        hours = int(self._wallclock_remaining / 3600)
        minutes = int((self._wallclock_remaining / 3600.0 - hours) * 60)
        print(f"{self._t / self._params.final_time * 100:.2f}% complete",
              f"t = {self._t:.2e}", f"dt = {self._integrator._dt:.2e}",
              f"Remaining: {hours} hr, {minutes} min")

    def form_dumpname(self, index):
        return f"dump{index:04d}.npz"

    def dump(self, ticker):
        """Checkpoint in the reference's dump format (Simulation.py:155-181)."""
        fname = self.form_dumpname(ticker.times_fired)
        data = {v.get_name(): self._data_trans.to_host(v[:]) for v in self._dump_vars}
        data.update({d.get_name(): self._data_trans.to_host(d.get_all()) for d in self._dump_dvars})
        tickers = {t.get_name(): t.dump() for t in self._tickers}
        from . import _dist
        if _dist.rank() != 0:                   # slabs were gathered above by every rank; one writer
            return
        np.savez(fname, **data, **tickers, curr_idx=self._dump_dvars[0].get_curr_idx(),
                 dt=self._integrator._dt, t=self._t, loop_counter=self._loop_counter,
                 params=self._params._original_params)

    def load(self, index):
        """Restart from a dump (Simulation.py:183-197; works here)."""
        arrays = np.load(self.form_dumpname(index), allow_pickle=True)
        for var in self._dump_vars:
            var.load(arrays[var._dump_name])          # re-sampled if the resolution changed
        for dvar in self._dump_dvars:
            dvar.load(arrays[dvar._dump_name])
            dvar.set_curr_idx(int(arrays["curr_idx"]))
        for ticker in self._tickers:
            if ticker._dump_name and ticker._dump_name in arrays.files:
                ticker.restore(arrays[ticker._dump_name])
        self._integrator._dt = float(arrays["dt"])
        self._t = float(arrays["t"])
        self._loop_counter = int(arrays["loop_counter"])

    # -- main loop plumbing
    def is_running(self):
        return self._t < self._params.final_time

    def register_ticker(self, ticker):
        self._tickers.append(ticker)

    def end_loop(self):
        self._loop_counter += 1
        self._t += self._integrator._dt
        for ticker in self._tickers:
            ticker.tick(self._t, self._loop_counter)
        # the fused z stage produces the CFL maxima and energy sums the tickers read; it only has to
        # for a step whose end_loop will fire one (every cfl_cadence / tracker_cadence loops).
        # `reductions = "always"` keeps them on for loops that read them outside the tickers
        # (a reader that finds none falls back to an explicit reduction: same numbers, two more kernels)
        if self.reductions == "always":
            want = True
        else:
            dt = self._integrator._dt
            want = any(t.due(self._t + dt, self._loop_counter + 1) for t in self._tickers)
        self._ctx.want_reductions = want

    def calc_time_remaining(self, ticker):
        self._timer.split()
        per_step = self._timer.diff / ticker._cadence
        self._wallclock_remaining = (per_step * (self._params.final_time - self._t)
                                     / self._integrator._dt)
