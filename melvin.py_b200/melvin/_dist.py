"""Slab decomposition behind the public API: when the process runs under an
initialised ``torch.distributed`` group of more than one rank (one process per GPU, e.g.
``torchrun``), every fully spectral ``Simulation`` is sharded -- spectral arrays in kz-slabs
(each rank owns a contiguous run of columns m), physical arrays in x-slabs (rows) -- and the
x-transformed intermediates are exchanged between the two passes of every transform.

The reference is single-device; nothing in a script changes: initial conditions are given as
full arrays on every rank (each keeps its slab), ``var[...]`` indexing uses *global* mode
numbers, reductions (``xp.max/sum/mean``, CFL, trackers) are global, and reading an array to the
host (``.get()``, dumps, saves) gathers it.  ``MLV_SHARD=0`` keeps N independent replicas.

Collectives go through torch.distributed (NCCL on GPUs, gloo in the CPU tests); the exchange of
the intermediates is one all-to-all per field.  (The example-specific steppers in
``melvin/sharded.py`` add the peer-memory and copy-engine exchanges used by bench.py.)
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def world():
    if os.environ.get("MLV_SHARD", "1") == "0":
        return 1
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if world() > 1 else 0


def all_to_all(recv, send):
    """Blocks of `send` ([peer][block], complex) to their peers; stream-ordered."""
    dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send))


def all_reduce_host(values, op):
    """Element-wise reduction of a small host array over the ranks (op: 'max', 'min', 'sum')."""
    from . import _backend
    t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=_backend.device()).clone()
    if op in ("max", "min"):
        # NaN must survive (numpy.max semantics, Integrator.py:41): reduce the NaN flags too
        nan = torch.isnan(t).to(torch.float64)
        t = torch.nan_to_num(t, nan=float("-inf") if op == "max" else float("inf"))
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.MIN)
        dist.all_reduce(nan, op=dist.ReduceOp.MAX)
        out = _backend.to_host(t)
        out[_backend.to_host(nan) > 0] = np.nan
        return out
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return _backend.to_host(t)


def gather(t, axis):
    """Concatenate the local slabs of all ranks along `axis` on the host of every rank."""
    from . import _backend
    t = t.contiguous()
    parts = [torch.empty_like(t) for _ in range(world())]
    if t.is_complex():
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(t))
    else:
        dist.all_gather(parts, t)
    return np.concatenate([_backend.to_host(p) for p in parts], axis=axis)
