"""Streamed ensembles: several independent simulations of one grid, each on its own CUDA
stream, stepped round-robin with HOST-resident states.

The reference keeps one simulation per process and moves its state with blocking copies
(``Variable.load`` / ``Variable.on_host``, reference melvin/Variable.py:67-85,135-136 and
melvin/DataTransferer.py).  When the states live on the host -- parameter sweeps, ensembles,
or a state larger than one wants resident -- a blocking round trip leaves the GPU idle for two
PCIe transfers per step.  Here every member owns a stream; its upload, time step and read-back
are queued on that stream without blocking the host, so the upload of member k+1 and the
read-back of member k-1 travel (on the two copy engines) while member k computes.

    ens = Ensemble(build, members=3)          # build(i) -> anything; runs under member i's stream
    for k in range(passes):
        with ens.turn(k) as m:                # waits for m's previous pass, enters its stream
            consume(m.payload)                # previous results are complete here
            ...load(pinned) / step / on_host(out=pinned)...
    ens.drain()

Inside a turn everything is ordinary public API (Variable.load, the example loop body,
Variable.on_host(out=...)); the helper only owns streams and events.  There is no CPU path.
"""
import contextlib

import torch

from . import _backend


class Member:
    def __init__(self, index):
        self.index = index
        self.stream = torch.cuda.Stream() if _backend.is_cuda() else None
        self.done = torch.cuda.Event() if _backend.is_cuda() else None
        self.passes = 0                 # passes queued so far
        self.payload = None

    def wait(self):
        """Block the host until everything queued in this member's last turn has finished."""
        if self.done is not None and self.passes:
            self.done.synchronize()


class Ensemble:
    def __init__(self, build, members=3):
        if members < 1:
            raise ValueError("an ensemble needs at least one member")
        self.members = [Member(i) for i in range(members)]
        for m in self.members:
            with self._on(m):
                # contexts (plans, scratch pools) are keyed by the creating stream
                # (_backend.context_for): every member gets its own, so members never share
                # scratch buffers across streams
                m.payload = build(m.index)

    @contextlib.contextmanager
    def _on(self, m):
        if m.stream is None:
            yield
        else:
            with torch.cuda.stream(m.stream):
                yield

    @contextlib.contextmanager
    def turn(self, k):
        """Member k mod M: wait for its previous pass, then queue on its stream; on exit an
        event marks the end of this pass."""
        m = self.members[k % len(self.members)]
        m.wait()
        with self._on(m):
            yield m
            m.passes += 1
            if m.done is not None:
                m.done.record(m.stream)

    def drain(self):
        for m in self.members:
            m.wait()
