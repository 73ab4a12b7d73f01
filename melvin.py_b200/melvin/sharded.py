"""Slab-decomposed fused step: one process per GPU, spectral state in kz-slabs
(column slabs), physical side in x-slabs (row slabs), one all-to-all of the
x-transformed intermediate per direction and step.

The reference is single-device; this module is the multi-GPU form of the loop body of
its single-scalar examples (examples/taylor_green_vortex.py:85-95,
examples/kelvin_helmholtz_instability.py:115-131):

    calc_velocity_from_vorticity(w, psi, ux, uz, solver)
    lin_op = coef * w.lap()
    dw[:] = -w.vec_dot_nabla(ux.getp(), uz.getp())
    integrator.integrate(w, dw, lin_op)
    simulation.end_loop()          # CFL every cfl_cadence loops, tracker every tracker_cadence

Per step and rank:  mlv_x_inverse (local columns, 3 fields)  ->  exchange  ->
mlv_advect_z (local rows)  ->  exchange  ->  mlv_x_forward (local columns, fused RHS +
AB + theta-scheme).  The x stencil of the conservative form is applied along the complete
x lines of the forward pass, so no halo exchange exists.  Reductions (CFL max, kinetic
energy) are 4-double all-reduces at ticker cadence.

Exchange buffers are laid out [field][peer][block]; three ways to move the blocks:
  * "p2p"  -- the producer kernels store every peer's block straight into that peer's
    receive buffer (CUDA IPC mapping, NVLink); only a one-element all-reduce orders
    producers and consumers.  Lowest latency: the choice for small grids, where a step is
    a few hundred microseconds.
  * "dma"  -- blocks are written locally and every contiguous peer block is moved by an
    asynchronous device-to-device copy into the peer's IPC-mapped receive buffer
    (mlv_p2p_copy on a side stream): copy engines drive NVLink with large transfers and
    take no SM from the kernels; the copies of field f overlap the inverse x pass of field
    f+1.  The choice for large grids (at 16384^2 the 16-byte pieces of the direct peer
    stores reach only ~200-300 GB/s per GPU).
  * "a2a"  -- same, with one asynchronous NCCL all-to-all per field (also the form the CPU
    tests run over gloo).
The mode is picked from the bytes a rank sends per step (MLV_EXCHANGE=p2p|dma|a2a overrides).
Collectives go through torch.distributed (NCCL on GPUs; gloo in the CPU tests of the host
logic).
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _backend, _capi
from .b200 import DeviceArray, sum_of_product


class ShardedScalarStepper:
    # exchange-buffer slots (fields) per direction, and fields actually moved per step
    N_INV, N_FWD = 3, 2
    MOVED_INV, MOVED_FWD = 3, 2

    def __init__(self, nx, nz, lx, lz, coef, dt, fd_order=2, ab_order=2, alpha=0.51,
                 cfl_cutoff=0.5, cfl_cadence=10, tracker_cadence=100, group=None, p2p=None, mode=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nx, self.nz, self.lx, self.lz = nx, nz, lx, lz
        self.coef, self.dt, self.alpha, self.order = float(coef), float(dt), float(alpha), int(ab_order)
        self.cfl_cutoff, self.cfl_cadence, self.tracker_cadence = cfl_cutoff, cfl_cadence, tracker_cadence
        self.dx, self.dz = lx / nx, lz / nz
        self.ctx = _backend.Context(nx, nz, lx, lz, False, fd_order)
        self.ctx.call("mlv_set_sharding", self.rank, self.world, 1, 1, count=False)
        info = _capi.Info()
        _capi.check(self.ctx.lib, self.ctx.lib.mlv_get_info(self.ctx.handle, ctypes.byref(info)))
        self.nn, self.nm = info.nn, info.nm
        self.rows, self.nml, self.nm_local = info.spec_rows, info.spec_cols, info.nm_local
        self.m_off = self.rank * self.nml if self.world > 1 else 0
        self.nxl = nx // self.world
        field = info.ibytes // 16                      # elements of one field of an exchange buffer
        self.inv_field = self.nxl * info.ipitch        # one peer block of an inverse field
        tiles = field // (self.world * self.nxl)       # tpr * ct
        self.fwd_field = tiles * self.nxl              # one peer block of a forward field
        self.inv_stride = self.world * self.inv_field  # elements between consecutive fields
        self.fwd_stride = self.world * self.fwd_field
        cplx = np.complex128
        # bytes this rank sends to its peers per step (3 inverse + 2 forward fields)
        self.bytes_exchanged_per_step = 16 * (self.MOVED_INV * self.inv_field + self.MOVED_FWD * self.fwd_field) * (self.world - 1)
        mode = mode or os.environ.get("MLV_EXCHANGE", "")
        if os.environ.get("MLV_NO_P2P"):
            mode = "a2a"
        if p2p is not None:
            mode = "p2p" if p2p else "a2a"
        if mode not in ("p2p", "dma", "a2a"):
            # latency-bound steps: direct peer stores (32-byte pieces, ~200-300 GB/s of NVLink, but no
            # extra launch).  Copy engines win once the contiguous block a rank sends to a peer is
            # large: measured at KH 4096^2 on 2 GPUs (22 MB blocks) 0.482 ms (dma) vs 0.523 ms (p2p),
            # on 8 GPUs (1.4 MB blocks) 0.673 ms (dma) vs 0.259 ms (p2p)
            big_total = self.bytes_exchanged_per_step >= 256 * 1024 * 1024
            big_block = self.world == 2 and 16 * self.inv_field >= 8 * 1024 * 1024
            mode = "dma" if (big_total or big_block) else "p2p"
        if self.world == 1 or not _backend.is_cuda():
            mode = "a2a"                               # (world 1: no exchange at all)
        self.mode = mode
        self.p2p = mode == "p2p"
        self._flags = False
        self._ipc = []
        if mode == "p2p":
            # receive buffers live in peer-mapped memory; the producer kernels of every rank
            # store straight into them over NVLink (compute + exchange in one kernel)
            self.inv_recv_ptr = self._setup_peers(0, self.N_INV * self.inv_stride * 16)[self.rank]
            self.fwd_recv_ptr = self._setup_peers(1, self.N_FWD * self.fwd_stride * 16)[self.rank]
            self.inv_send_ptr, self.fwd_send_ptr = self.inv_recv_ptr, self.fwd_recv_ptr
            self._sync = _backend.zeros((1,), np.float64)
            # producers and consumers of the exchange are ordered across ranks ON THE DEVICE:
            # every producer CTA bumps an arrival counter in each rank's memory, consumer CTAs
            # spin on their own (mlv_set_peer_flags).  MLV_P2P_BARRIER=nccl: the older form, a
            # one-element all-reduce between the kernels
            self._flags = os.environ.get("MLV_P2P_BARRIER", "flags") != "nccl"
            if self._flags:
                ptrs = self._setup_peers(None, 64, register=False)
                arr = (ctypes.c_void_p * self.world)(*ptrs)
                _capi.check(self.ctx.lib, self.ctx.lib.mlv_set_peer_flags(self.ctx.handle, arr))
                _backend.synchronize()
                dist.barrier(group=self.group)
        elif mode == "dma":
            # peer-mapped receive buffers filled by copy engines from local send buffers
            self.inv_peers = self._setup_peers(0, self.N_INV * self.inv_stride * 16, register=False)
            self.fwd_peers = self._setup_peers(1, self.N_FWD * self.fwd_stride * 16, register=False)
            self.inv_recv_ptr, self.fwd_recv_ptr = self.inv_peers[self.rank], self.fwd_peers[self.rank]
            self.inv_send = _backend.zeros((self.N_INV, self.inv_stride), cplx)
            self.fwd_send = _backend.zeros((self.N_FWD, self.fwd_stride), cplx)
            self.inv_send_ptr, self.fwd_send_ptr = self.inv_send.data_ptr(), self.fwd_send.data_ptr()
            self._sync = _backend.zeros((1,), np.float64)
            # several copy streams: copies to different peers run on different copy engines
            ncs = int(os.environ.get("MLV_COPY_STREAMS", "4"))
            # (copies done by CTAs, MLV_COPY_CTAS: high-priority streams, so that they get the next free
            # SM slot instead of queueing behind a whole transform kernel)
            prio = -1 if int(os.environ.get("MLV_COPY_CTAS", "0") or 0) > 0 else 0
            self._copy_streams = [torch.cuda.Stream(priority=prio) for _ in range(max(1, min(ncs, self.world)))]
            self._ev = [torch.cuda.Event() for _ in range(6)]
            self._join_ev = [torch.cuda.Event() for _ in self._copy_streams]
        else:
            self.inv_send = _backend.zeros((self.N_INV, self.inv_stride), cplx)
            self.fwd_send = _backend.zeros((self.N_FWD, self.fwd_stride), cplx)
            if self.world > 1:
                self.inv_recv = _backend.zeros((self.N_INV, self.inv_stride), cplx)
                self.fwd_recv = _backend.zeros((self.N_FWD, self.fwd_stride), cplx)
            else:
                self.inv_recv, self.fwd_recv = self.inv_send, self.fwd_send
            self.inv_send_ptr, self.inv_recv_ptr = self.inv_send.data_ptr(), self.inv_recv.data_ptr()
            self.fwd_send_ptr, self.fwd_recv_ptr = self.fwd_send.data_ptr(), self.fwd_recv.data_ptr()
        # forward exchange pipelined by row blocks: the z stage runs block by block and every
        # finished block is shipped while the next one is computed (dma mode; a2a keeps one
        # transfer per field but uses the same block layout when asked to, for the CPU tests)
        nch = int(os.environ.get("MLV_FWD_CHUNKS", "4" if mode == "dma" else "1"))
        while nch > 1 and (self.nxl % nch or (self.nxl // nch) % 2 or self.nxl // nch < 2):
            nch //= 2
        self.nchunks = max(1, nch) if self.world > 1 else 1
        self.chunk_rows = self.nxl // self.nchunks
        if self.nchunks > 1:
            self.ctx.call("mlv_set_forward_blocks", self.chunk_rows, count=False)
        self.fwd_block = self.fwd_field // self.nchunks      # one row block of one peer and field
        self.cur = 0
        self.hidx = 0
        # CUDA graphs of the step (scalar stepper; peer-store exchange ordered on the device, or one GPU)
        graphable = (type(self) is ShardedScalarStepper and _backend.is_cuda()
                     and (self.world == 1 or (self.p2p and self._flags)))
        self._graphs = {} if graphable and os.environ.get("MLV_GRAPH", "1") != "0" else None
        self.red4 = _backend.zeros((4,), np.float64)
        self._alloc_state()
        self.t = 0.0
        self.loop = 0
        self._cfl_counter = 0
        self._trk_counter = 0
        self.ke_times, self.ke = [], []
        self._prebuild()

    def _alloc_state(self):
        cplx = np.complex128
        self.w = [_backend.zeros((self.rows, self.nml), cplx), _backend.zeros((self.rows, self.nml), cplx)]
        self.hist = _backend.zeros((self.order, self.rows, self.nml), cplx)

    def _setup_peers(self, which, nbytes, register=True):
        """Allocate this rank's receive buffer, exchange CUDA IPC handles, open the peers'.
        Returns the list of mapped pointers (own entry = own buffer); `register` hands them to
        the library so that its kernels store into them directly."""
        lib, h = self.ctx.lib, self.ctx.handle
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        _capi.check(lib, lib.mlv_p2p_alloc(h, nbytes, ctypes.byref(own), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=self.group)
        ptrs = (ctypes.c_void_p * self.world)()
        for r, raw in enumerate(handles):
            if r == self.rank:
                ptrs[r] = own.value
            else:
                peer = ctypes.c_void_p()
                _capi.check(lib, lib.mlv_p2p_open(h, ctypes.create_string_buffer(raw, 64), ctypes.byref(peer)))
                ptrs[r] = peer.value
                self._ipc.append((peer.value, 1))
        self._ipc.append((own.value, 0))
        if register:
            _capi.check(lib, lib.mlv_set_peer_buffers(h, which, ptrs))
        return [int(ptrs[r]) for r in range(self.world)]

    def close(self):
        """Release peer mappings (call on every rank, after a barrier)."""
        if self._ipc:
            _backend.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for ptr, opened in self._ipc:
                self.ctx.lib.mlv_p2p_close(self.ctx.handle, ctypes.c_void_p(ptr), opened)
            self._ipc = []

    # ------------------------------------------------------------------ state
    def load_spectral(self, w_full):
        """Keep this rank's column slab of a full (2nn+1, nm) spectral array (host)."""
        self._load_slab(self.w[self.cur], w_full)
        DeviceArray(self.hist)[...] = 0.0
        self.hidx = 0

    def gather_spectral(self):
        """Full (2nn+1, nm) spectral state on the host of every rank."""
        return self._gather(self.w[self.cur])

    # ------------------------------------------------------------------- step
    def _a2a(self, recv, send, f):
        """Asynchronous all-to-all of field f ([peer][block] on both sides): ordered after the
        work already on the compute stream, overlaps what is launched next."""
        return dist.all_to_all_single(torch.view_as_real(recv[f]), torch.view_as_real(send[f]),
                                      group=self.group, async_op=True)

    def _dma(self, which, f, ev, chunk=None):
        """Copy engine exchange of field f: after the work already on the compute stream, move
        block h of the local send buffer into slot `rank` of peer h's receive buffer (forward
        buffers: optionally only row block `chunk` of it)."""
        comp = torch.cuda.current_stream()
        ev.record(comp)
        streams = self._copy_streams
        for cs in streams:
            cs.wait_event(ev)
        block, stride = (self.inv_field, self.inv_stride) if which == 0 else (self.fwd_field, self.fwd_stride)
        peers, send = (self.inv_peers, self.inv_send_ptr) if which == 0 else (self.fwd_peers, self.fwd_send_ptr)
        sub, n = (0, block) if chunk is None else (chunk * self.fwd_block, self.fwd_block)
        lib, hnd = self.ctx.lib, self.ctx.handle
        vp = ctypes.c_void_p
        for i in range(self.world):
            h = (self.rank + 1 + i) % self.world       # spread the first copies over the peers
            dst = peers[h] + 16 * (f * stride + self.rank * block + sub)
            src = send + 16 * (f * stride + h * block + sub)
            cs = streams[i % len(streams)]
            _capi.check(lib, lib.mlv_p2p_copy(hnd, vp(dst), vp(src), 16 * n, vp(cs.cuda_stream)))

    def _dma_join(self, ev):
        """Compute stream waits for this rank's outgoing copies; the all-reduce that follows
        then tells every rank that all incoming blocks have landed."""
        comp = torch.cuda.current_stream()
        for cs, e in zip(self._copy_streams, self._join_ev):
            e.record(cs)
            comp.wait_event(e)
        dist.all_reduce(self._sync, group=self.group)

    # ------------------------------------------------ generic exchange rounds (multi-field steppers)
    def _slot(self, which, recv, slot):
        base = {(0, 0): self.inv_send_ptr, (0, 1): self.inv_recv_ptr,
                (1, 0): self.fwd_send_ptr, (1, 1): self.fwd_recv_ptr}[(which, int(recv))]
        return base + 16 * slot * (self.inv_stride if which == 0 else self.fwd_stride)

    def _load_slab(self, dst, full):
        """dst (rows, nml) := this rank's column slab of a full (2nn+1, nm) spectral array (host
        array, or a device tensor of a full-size context)."""
        out = DeviceArray(dst)
        out[...] = 0.0
        if self.nm_local <= 0:
            return
        cols = slice(self.m_off, self.m_off + self.nm_local)
        if isinstance(full, torch.Tensor):
            out[:, :self.nm_local] = DeviceArray(full)[:, cols]
        else:
            out[:, :self.nm_local] = np.ascontiguousarray(np.asarray(full)[:, cols])

    def _gather(self, local):
        if self.world == 1:
            return _backend.to_host(local)[:, : self.nm]
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(local), group=self.group)
        return np.concatenate([_backend.to_host(p) for p in parts], axis=1)[:, : self.nm]

    def _mark(self):
        """Events at the current end of every copy stream (copy-engine exchange)."""
        evs = []
        for cs in self._copy_streams:
            e = torch.cuda.Event()
            e.record(cs)
            evs.append(e)
        return evs

    def _join_mark(self, evs):
        """Compute stream waits for this rank's copies up to a mark; the all-reduce then tells every
        rank that the matching blocks of all ranks have landed."""
        comp = torch.cuda.current_stream()
        for e in evs:
            comp.wait_event(e)
        dist.all_reduce(self._sync, group=self.group)

    def _inverse_round(self, fields, marks_at=None):
        """fields: [(source spectral tensor, MLV_OP_*, inverse slot)].  Inverse x pass of every
        field on the local columns, then row block h of every field at rank h.
        marks_at (copy-engine exchange only): field counts after which the copy streams are marked;
        the round then does NOT join -- the caller joins mark by mark (`_join_mark`), so that the
        first consumers start while the copies of the later fields are still travelling."""
        ctx, vp = self.ctx, ctypes.c_void_p
        if self.mode == "dma" or not (self.world == 1 or self.p2p):
            works, marks = [], {}
            for i, (src, op, slot) in enumerate(fields):
                ctx.call("mlv_x_inverse", 1, (vp * 1)(src.data_ptr()), (ctypes.c_int32 * 1)(op),
                         (vp * 1)(self._slot(0, False, slot)))
                if self.mode == "dma":
                    self._dma(0, slot, self._ev[i % 3])
                    if marks_at and (i + 1) in marks_at:
                        marks[i + 1] = self._mark()
                else:
                    works.append(self._a2a(self.inv_recv, self.inv_send, slot))
            if self.mode == "dma" and not marks_at:
                self._dma_join(self._ev[3])
            for wk in works:
                wk.wait()
            return marks
        # world 1 / peer stores: batches of up to 4 fields share the column stash of their source
        for b in range(0, len(fields), 4):
            grp = fields[b:b + 4]
            n = len(grp)
            ctx.call("mlv_x_inverse", n, (vp * n)(*[g[0].data_ptr() for g in grp]),
                     (ctypes.c_int32 * n)(*[g[1] for g in grp]),
                     (vp * n)(*[self._slot(0, False, g[2]) for g in grp]))
        if self.world > 1 and not self._flags:
            dist.all_reduce(self._sync, group=self.group)

    def _advect_round(self, jobs, before=None):
        """jobs: [(inverse slots ux, uz, q, forward slots a, b, red4 tensor or None)].  Fused z
        stage on the local rows, then tile block h of every forward field at rank h.
        before: {job index: callable} run right before that job is launched (deferred joins)."""
        ctx, vp = self.ctx, ctypes.c_void_p
        chunked = self.mode == "dma" or not (self.world == 1 or self.p2p)
        works = []
        # the z stage computes its CFL / energy partials only on the steps whose tickers read them
        ctx.want_reductions = any(job[5] is not None for job in jobs)
        ctx.sync_reduction_mode()
        for k, (sux, suz, sq, fa, fb, red) in enumerate(jobs):
            if before and k in before:
                before[k]()
            args = (vp(self._slot(0, True, sux)), vp(self._slot(0, True, suz)), vp(self._slot(0, True, sq)),
                    vp(self._slot(1, False, fa)), vp(self._slot(1, False, fb)))
            redp = vp(red.data_ptr()) if red is not None else None
            if not chunked:
                ctx.call("mlv_advect_z", *args, redp)
                continue
            for c in range(self.nchunks):
                ctx.call("mlv_advect_z_rows", *args, c * self.chunk_rows, self.chunk_rows,
                         redp if c == self.nchunks - 1 else None)
                if self.mode == "dma":
                    for f in (fa, fb):
                        self._dma(1, f, self._ev[c % 3], chunk=c if self.nchunks > 1 else None)
            if self.mode != "dma":
                works += [self._a2a(self.fwd_recv, self.fwd_send, f) for f in (fa, fb)]
        if self.mode == "dma":
            self._dma_join(self._ev[4])
        elif not chunked and self.world > 1 and not self._flags:
            dist.all_reduce(self._sync, group=self.group)
        for wk in works:
            wk.wait()

    def _forward(self, slots, coefs, lin, lcoef, q_in, q_out, hist, hidx):
        """Forward x pass of the operand pairs in `slots` (d/dx, d/dz alternating) with the fused
        right-hand side (lin = [(coef, MLV_OP_*, tensor)]), AB predictor and theta-scheme."""
        d = _capi.XFwd()
        d.nf, d.mode = len(slots), 1
        for i, (sl, cf) in enumerate(zip(slots, coefs)):
            d.src[i] = self._slot(1, True, sl)
            d.sym[i] = _capi.SYM_FDX if i % 2 == 0 else _capi.SYM_FDZ
            d.coef[i] = cf
        d.lin = _capi.make_lin_terms([(c, op, t.data_ptr()) for (c, op, t) in lin])
        g = d.integ
        g.ab_order, g.scheme = self.order, _capi.SCHEME_SI_LAP
        g.alpha, g.lcoef, g.dt = self.alpha, lcoef, self.dt
        g.q_in, g.q_out = q_in.data_ptr(), q_out.data_ptr()
        lv = [hist[(hidx - k) % self.order].data_ptr() for k in range(self.order)]
        g.f0, g.fm1 = lv[0], lv[1]
        if self.order == 4:
            g.fm2, g.fm3 = lv[2], lv[3]
        self.ctx.call("mlv_x_forward", ctypes.byref(d))

    def _prebuild(self):
        """ctypes argument blocks that do not change from step to step."""
        vp = ctypes.c_void_p
        ops = (_capi.OP_IDENT, _capi.OP_UX, _capi.OP_UZ)
        self._ops = (ctypes.c_int32 * 3)(*ops)
        self._dsts = (vp * 3)(*[self.inv_send_ptr + 16 * f * self.inv_stride for f in range(3)])
        self._srcs = [(vp * 3)(w.data_ptr(), w.data_ptr(), w.data_ptr()) for w in self.w]
        # one field per launch (a2a mode: the exchange of a field overlaps the next launch)
        self._ops1 = [(ctypes.c_int32 * 1)(o) for o in ops]
        self._dsts1 = [(vp * 1)(self.inv_send_ptr + 16 * f * self.inv_stride) for f in range(3)]
        self._srcs1 = [(vp * 1)(w.data_ptr()) for w in self.w]
        ir, fs = self.inv_recv_ptr, self.fwd_send_ptr
        self._zargs = (vp(ir + 16 * self.inv_stride), vp(ir + 32 * self.inv_stride), vp(ir),
                       vp(fs), vp(fs + 16 * self.fwd_stride))
        self._redp = vp(self.red4.data_ptr())
        self._zrows = [self._zargs + (c * self.chunk_rows, self.chunk_rows) for c in range(self.nchunks)]
        d = _capi.XFwd()
        d.nf, d.mode = 2, 1
        d.src[0] = self.fwd_recv_ptr
        d.src[1] = self.fwd_recv_ptr + 16 * self.fwd_stride
        d.sym[0], d.sym[1] = _capi.SYM_FDX, _capi.SYM_FDZ
        d.coef[0] = d.coef[1] = -1.0
        d.lin = _capi.make_lin_terms([])
        d.integ.ab_order, d.integ.scheme = self.order, _capi.SCHEME_SI_LAP
        d.integ.alpha, d.integ.lcoef = self.alpha, self.coef
        self._xfwd = d
        self._hist_ptrs = [self.hist[k].data_ptr() for k in range(self.order)]

    def _tickers_due(self):
        """Will end_loop of this step fire the CFL or the tracker ticker (Ticker.py:14-25)?"""
        return self._cfl_counter < self.loop + 1 or self._trk_counter < self.loop + 1

    def step(self):
        # peer-store exchange (and the single-GPU stepper): nothing in a step depends on the host
        # -- the kernels of a step are ordered across ranks on the device -- so a step is replayed
        # as a CUDA graph: one launch instead of five, no Python or ctypes latency in between
        if self._graphs is not None:
            key = (self.cur, self.hidx, self.dt, self._tickers_due())
            g = self._graphs.get(key)
            if g is None and self.loop >= 2 * self.order:       # after the eager warm-up steps
                if len(self._graphs) > 4 * self.order:          # dt changed: drop the stale graphs
                    self._graphs.clear()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_step()
                self._graphs[key] = g
            if g is not None:
                g.replay()
                self._advance()
                return
        self._launch_step()
        self._advance()

    def _advance(self):
        self.cur = 1 - self.cur
        self.hidx = (self.hidx + 1) % self.order
        self.end_loop()

    def _launch_step(self):
        ctx = self.ctx
        w_in, w_out = self.w[self.cur], self.w[1 - self.cur]
        wp = w_in.data_ptr()
        # the four reductions are combined only on the steps whose tickers read them
        redp = self._redp if self._tickers_due() else None
        ctx.want_reductions = redp is not None         # (the z stage skips its partials otherwise)
        ctx.sync_reduction_mode()
        if self.mode == "dma":
            # one launch per field; the copies of field f (row block h -> rank h) run on the copy
            # engines while the x pass of field f+1 runs on the SMs
            for f in range(3):
                ctx.call("mlv_x_inverse", 1, self._srcs1[self.cur], self._ops1[f], self._dsts1[f])
                self._dma(0, f, self._ev[f])
            self._dma_join(self._ev[3])
            # z stage block by block; the two fields of a finished row block leave while the
            # next block is computed
            for c in range(self.nchunks):
                ctx.call("mlv_advect_z_rows", *self._zrows[c], redp if c == self.nchunks - 1 else None)
                for f in range(2):
                    self._dma(1, f, self._ev[c % 3], chunk=c if self.nchunks > 1 else None)
            self._dma_join(self._ev[4])
        elif self.world == 1 or self.p2p:
            # 1. inverse x pass on the local columns: q = w, ux, uz (psi shared inside the
            #    kernel); with peer memory every block lands in its consumer's buffer
            ctx.call("mlv_x_inverse", 3, self._srcs[self.cur], self._ops, self._dsts)
            # 2. producers and consumers are ordered across ranks by arrival counters in peer
            #    memory (no collective, no host synchronisation)
            if self.world > 1 and not self._flags:
                dist.all_reduce(self._sync, group=self.group)
            # 3. physical-space stage on the local rows
            ctx.call("mlv_advect_z", *self._zargs, redp)
            if self.world > 1 and not self._flags:
                dist.all_reduce(self._sync, group=self.group)
        else:
            # 1.+2. one launch and one all-to-all per field: the transpose of field f (row block
            #       h goes to rank h) travels while the x pass of field f+1 runs
            works = []
            for f in range(3):
                ctx.call("mlv_x_inverse", 1, self._srcs1[self.cur], self._ops1[f], self._dsts1[f])
                works.append(self._a2a(self.inv_recv, self.inv_send, f))
            for wk in works:
                wk.wait()
            # 3. physical-space stage on the local rows
            for c in range(self.nchunks):
                ctx.call("mlv_advect_z_rows", *self._zrows[c], redp if c == self.nchunks - 1 else None)
            # 4. transpose back: tile block h goes to rank h
            works = [self._a2a(self.fwd_recv, self.fwd_send, f) for f in range(2)]
            for wk in works:
                wk.wait()
        # 5. forward x pass + RHS + AB + theta-scheme on the local columns
        d = self._xfwd
        g = d.integ
        g.dt = self.dt
        g.q_in, g.q_out = wp, w_out.data_ptr()
        lv = [self._hist_ptrs[(self.hidx - k) % self.order] for k in range(self.order)]
        g.f0, g.fm1 = lv[0], lv[1]
        if self.order == 4:
            g.fm2, g.fm3 = lv[2], lv[3]
        ctx.call("mlv_x_forward", ctypes.byref(d))

    # ------------------------------------------------- tickers (Simulation.end_loop)
    def _global_reductions(self):
        r = self.red4.clone()
        if self.world > 1:
            mx, sm = r[:2].clone(), r[2:].clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=self.group)
            r = torch.cat([mx, sm])
        return _backend.to_host(r)

    def end_loop(self):
        self.loop += 1
        self.t += self.dt
        need_cfl = self._cfl_counter < self.loop
        need_trk = self._trk_counter < self.loop
        if not (need_cfl or need_trk):
            return
        mx, mz, sx, sz = self._global_reductions()
        if need_cfl:                                   # Integrator.set_dt (Integrator.py:35-44)
            with np.errstate(divide="ignore", invalid="ignore"):
                cfl_dt = min(np.float64(self.dx) / mx, np.float64(self.dz) / mz)
            if self.dt > cfl_dt or np.isnan(cfl_dt):
                raise Exception("CFL condition breached")
            while self.dt > self.cfl_cutoff * cfl_dt:
                self.dt = self.dt * 0.9
            self._cfl_counter += self.cfl_cadence
        if need_trk:                                   # calc_kinetic_energy (utility.py:42-59)
            self.ke_times.append(self.t)
            self.ke.append(0.5 * (sz + sx) / (self.nx * self.nz))
            self._extra_trackers()
            self._trk_counter += self.tracker_cadence

    def _extra_trackers(self):
        pass


class ShardedDoubleDiffusiveStepper(ShardedScalarStepper):
    """Slab-decomposed form of the double-diffusive loop (BASELINE config 4;
    examples/double_diffusive_convection.py:100-126): vorticity w, temperature tmp and
    composition xi advected by the velocity of the *old* vorticity, coupling terms
    Pr*(ddx xi - ddx tmp), -uz and -uz/R0 assembled in the epilogue of the forward x pass, the
    script's fix-up `tmp[:, 0] = 0; xi[:, 0] = 0`, kinetic-energy and Nusselt trackers.
    One exchange round per direction and step: 5 inverse fields, 6 forward fields."""
    N_INV, N_FWD = 5, 6
    MOVED_INV, MOVED_FWD = 5, 6

    def __init__(self, nx, nz, lx, lz, Pr, R0, tau, dt, **kw):
        self.Pr, self.R0, self.tau = float(Pr), float(R0), float(tau)
        super().__init__(nx, nz, lx, lz, Pr, dt, **kw)

    def _alloc_state(self):
        cplx = np.complex128
        mk = lambda: [_backend.zeros((self.rows, self.nml), cplx) for _ in range(2)]   # noqa: E731
        self.q = {"w": mk(), "tmp": mk(), "xi": mk()}
        self.h = {k: _backend.zeros((self.order, self.rows, self.nml), cplx) for k in self.q}
        self.w = self.q["w"]
        self.hist = self.h["w"]
        self.nu, self._phys = [], None

    def _prebuild(self):
        pass

    def load_spectral(self, w_full, tmp_full, xi_full):
        for k, full in (("w", w_full), ("tmp", tmp_full), ("xi", xi_full)):
            self._load_slab(self.q[k][self.cur], full)
            DeviceArray(self.h[k])[...] = 0.0
        self.hidx = 0

    def gather_spectral(self):
        return tuple(self._gather(self.q[k][self.cur]) for k in ("w", "tmp", "xi"))

    def step(self):
        c, n = self.cur, 1 - self.cur
        w, tmp, xi = (self.q[k][c] for k in ("w", "tmp", "xi"))
        OP = _capi
        # inverse slots: 0 w, 1 ux, 2 uz (all from the old vorticity), 3 tmp, 4 xi
        fields = [(w, OP.OP_IDENT, 0), (w, OP.OP_UX, 1), (w, OP.OP_UZ, 2),
                  (tmp, OP.OP_IDENT, 3), (xi, OP.OP_IDENT, 4)]
        red = self.red4 if self._tickers_due() else None
        jobs = [(1, 2, 0, 0, 1, red), (1, 2, 3, 2, 3, None), (1, 2, 4, 4, 5, None)]
        if self.mode == "dma" and os.environ.get("MLV_SPLIT_JOIN", "1") != "0":
            # joins by mark (see ShardedTearingStepper.step): every z-stage job starts as soon as
            # ITS scalar has landed, the later fields travel during the earlier jobs
            marks = self._inverse_round(fields, marks_at=(3, 4, 5))
            self._join_mark(marks[3])
            self._advect_round(jobs, before={1: lambda: self._join_mark(marks[4]),
                                             2: lambda: self._join_mark(marks[5])})
        else:
            self._inverse_round(fields)
            self._advect_round(jobs)
        Pr, R0, tau = self.Pr, self.R0, self.tau
        self._forward((0, 1), (-1.0, -1.0), [(Pr, OP.OP_DDX, xi), (-Pr, OP.OP_DDX, tmp)], Pr,
                      w, self.q["w"][n], self.h["w"], self.hidx)
        self._forward((2, 3), (-1.0, -1.0), [(-1.0, OP.OP_UZ, w)], 1.0,
                      tmp, self.q["tmp"][n], self.h["tmp"], self.hidx)
        self._forward((4, 5), (-1.0, -1.0), [(-1.0 / R0, OP.OP_UZ, w)], tau,
                      xi, self.q["xi"][n], self.h["xi"], self.hidx)
        if self.m_off == 0 and self.nm_local > 0:      # the rank that owns the m = 0 column
            DeviceArray(self.q["tmp"][n])[:, 0] = 0.0
            DeviceArray(self.q["xi"][n])[:, 0] = 0.0
        self.cur = n
        self.hidx = (self.hidx + 1) % self.order
        self.end_loop()

    def _extra_trackers(self):
        """Nusselt number 1 - mean(T_p uz_p) on this loop's (pre-update) fields
        (examples/double_diffusive_convection.py:20-23): inverse z pass of the local rows."""
        vp = ctypes.c_void_p
        if self._phys is None:
            self._phys = [_backend.zeros((self.nxl, self.nz), np.float64) for _ in range(2)]
        for slot, out in ((3, self._phys[0]), (2, self._phys[1])):
            self.ctx.call("mlv_z_inverse", vp(self._slot(0, True, slot)), vp(out.data_ptr()))
        local = sum_of_product(DeviceArray(self._phys[0]), DeviceArray(self._phys[1]))   # mlv_reduce
        sm = torch.tensor([local], dtype=torch.float64, device=_backend.device())
        if self.world > 1:
            dist.all_reduce(sm, group=self.group)
        self.nu.append(1.0 - float(_backend.to_host(sm)[0]) / (self.nx * self.nz))


class ShardedTearingStepper(ShardedScalarStepper):
    """Slab-decomposed form of the resistive-tearing (MHD) loop (BASELINE config 5;
    examples/resistive_tearing_instability.py:125-148): vorticity w and current j, velocity
    from w and magnetic field from j, and the reference's ordering -- the j equation advects
    the *updated* vorticity with the magnetic field.  Two exchange rounds per direction and
    step: 6 + 1 inverse fields, 6 + 2 forward fields."""
    N_INV, N_FWD = 7, 8
    MOVED_INV, MOVED_FWD = 7, 8

    def __init__(self, nx, nz, lx, lz, Re, S, dt, **kw):
        self.Re, self.S = float(Re), float(S)
        super().__init__(nx, nz, lx, lz, 1.0 / Re, dt, **kw)

    def _alloc_state(self):
        cplx = np.complex128
        mk = lambda: [_backend.zeros((self.rows, self.nml), cplx) for _ in range(2)]   # noqa: E731
        self.q = {"w": mk(), "j": mk()}
        self.h = {k: _backend.zeros((self.order, self.rows, self.nml), cplx) for k in self.q}
        self.w = self.q["w"]
        self.hist = self.h["w"]

    def _prebuild(self):
        pass

    def load_spectral(self, w_full, j_full):
        for k, full in (("w", w_full), ("j", j_full)):
            self._load_slab(self.q[k][self.cur], full)
            DeviceArray(self.h[k])[...] = 0.0
        self.hidx = 0

    def gather_spectral(self):
        return tuple(self._gather(self.q[k][self.cur]) for k in ("w", "j"))

    def step(self):
        c, n = self.cur, 1 - self.cur
        w, j = self.q["w"][c], self.q["j"][c]
        w_new, j_new = self.q["w"][n], self.q["j"][n]
        OP = _capi
        # inverse slots: 0 w, 1 ux, 2 uz, 3 j, 4 bx, 5 bz, 6 updated w
        fields = [(w, OP.OP_IDENT, 0), (w, OP.OP_UX, 1), (w, OP.OP_UZ, 2),
                  (j, OP.OP_IDENT, 3), (j, OP.OP_UX, 4), (j, OP.OP_UZ, 5)]
        # forward slots: (0,1) u.grad w, (2,3) b.grad j, (4,5) u.grad j, (6,7) b.grad w_new
        red = self.red4 if self._tickers_due() else None
        jobs = [(1, 2, 0, 0, 1, red), (4, 5, 3, 2, 3, None), (1, 2, 3, 4, 5, None)]
        if self.mode == "dma" and os.environ.get("MLV_SPLIT_JOIN", "1") != "0":
            # copy engines lag behind the x passes (443 GB/s vs 154 MB per field every ~0.25 ms at 8 GPUs):
            # the first z-stage job needs only the three fields derived from w -- start it as soon as
            # those have landed, the copies of the j fields travel meanwhile
            marks = self._inverse_round(fields, marks_at=(3, 6))
            self._join_mark(marks[3])
            self._advect_round(jobs, before={1: lambda: self._join_mark(marks[6])})
        else:
            self._inverse_round(fields)
            self._advect_round(jobs)
        self._forward((0, 1, 2, 3), (-1.0, -1.0, 1.0, 1.0), [], 1.0 / self.Re, w, w_new, self.h["w"], self.hidx)
        self._inverse_round([(w_new, OP.OP_IDENT, 6)])
        self._advect_round([(4, 5, 6, 6, 7, None)])
        self._forward((4, 5, 6, 7), (-1.0, -1.0, 1.0, 1.0), [], 1.0 / self.S, j, j_new, self.h["j"], self.hidx)
        self.cur = n
        self.hidx = (self.hidx + 1) % self.order
        self.end_loop()
