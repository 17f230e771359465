"""Row-slab sharding of the EASU+RCAS path across the GPUs of one box.

The reference has no multi-GPU path; this is new.  Every output pixel depends on a bounded input
neighbourhood, so the OUTPUT is cut into G contiguous row slabs, rank k owns input rows
[k*inH/G, (k+1)*inH/G) and needs a few more rows above/below (the EASU footprint of its output slab
extended by one output row each side, so that RCAS's +-1-row taps need no second exchange).
The only communication is that halo: point-to-point rows between neighbouring ranks
(torch.distributed batch_isend_irecv -> ncclSend/ncclRecv over NVLink); there is no collective and a
1-GPU run issues no communication at all.

SlabPlan is pure integer geometry (unit-tested on CPU); ShardedUpscaler moves the rows and launches
the kernels through the C ABI with image *windows* (row0/rows), so clamping at the true image
border and reading halo rows at slab borders are both handled by the same kernels.
"""
import math

import numpy as np

from . import api


def _cell(o, scale, offset):
    """floor(o*scale+offset) in the kernels' float arithmetic (mul and add rounded separately)."""
    return int(math.floor(np.float32(np.float32(np.float32(o) * np.float32(scale)) + np.float32(offset))))


class SlabPlan:
    def __init__(self, in_h, out_h, world, easu_con):
        self.in_h, self.out_h, self.world = int(in_h), int(out_h), int(world)
        self.scale = np.array([easu_con[1]], dtype=np.uint32).view(np.float32)[0]
        self.offset = np.array([easu_con[3]], dtype=np.uint32).view(np.float32)[0]

    def out_rows(self, rank):
        """Output rows [y0,y1) of this rank's slab."""
        return rank * self.out_h // self.world, (rank + 1) * self.out_h // self.world

    def easu_rows(self, rank):
        """EASU is run for the slab plus a one-row apron (what RCAS reads)."""
        y0, y1 = self.out_rows(rank)
        return max(y0 - 1, 0), min(y1 + 1, self.out_h)

    def owned_in_rows(self, rank):
        return rank * self.in_h // self.world, (rank + 1) * self.in_h // self.world

    def needed_in_rows(self, rank):
        """Input rows [r0,r1) the rank's EASU pass reads (clamped to the image)."""
        e0, e1 = self.easu_rows(rank)
        if e1 <= e0:
            return 0, 0
        lo = _cell(e0, self.scale, self.offset) - 1
        hi = _cell(e1 - 1, self.scale, self.offset) + 2
        lo = min(max(lo, 0), self.in_h - 1)
        hi = min(max(hi, 0), self.in_h - 1)
        return lo, hi + 1

    def transfers(self, rank):
        """(sends, recvs): lists of (peer, first_row, end_row) in logical input rows."""
        sends, recvs = [], []
        own0, own1 = self.owned_in_rows(rank)
        need0, need1 = self.needed_in_rows(rank)
        for peer in range(self.world):
            if peer == rank:
                continue
            p_own0, p_own1 = self.owned_in_rows(peer)
            p_need0, p_need1 = self.needed_in_rows(peer)
            a, b = max(own0, p_need0), min(own1, p_need1)      # my rows the peer needs
            if b > a:
                sends.append((peer, a, b))
            a, b = max(p_own0, need0), min(p_own1, need1)      # the peer's rows I need
            if b > a:
                recvs.append((peer, a, b))
        return sends, recvs

    def halo_bytes(self, rank, width, bytes_per_pixel):
        return sum((b - a) * width * bytes_per_pixel for _, a, b in self.transfers(rank)[1])


def exchange_halo(plan, rank, owned, window, dist=None):
    """Fill `window` (rows needed_in_rows(rank)) from `owned` (rows owned_in_rows(rank)) and the peers.

    `owned` may be a view INTO `window` (see ShardedUpscaler.owned): then nothing is copied locally and only the
    halo rows move.  Works on any backend: gloo with CPU tensors (tests) or nccl with CUDA tensors (production).
    """
    own0, own1 = plan.owned_in_rows(rank)
    need0, need1 = plan.needed_in_rows(rank)
    a, b = max(own0, need0), min(own1, need1)
    if b > a and owned[a - own0:a - own0 + 1].data_ptr() != window[a - need0:a - need0 + 1].data_ptr():
        window[a - need0:b - need0].copy_(owned[a - own0:b - own0])
    sends, recvs = plan.transfers(rank)
    if not sends and not recvs:
        return 0
    if dist is None:
        import torch.distributed as dist
    ops = []
    for peer, r0, r1 in sends:
        ops.append(dist.P2POp(dist.isend, owned[r0 - own0:r1 - own0], peer))
    for peer, r0, r1 in recvs:
        ops.append(dist.P2POp(dist.irecv, window[r0 - need0:r1 - need0], peer))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return len(ops)


def exchange_halo_many(plan, rank, frames, dist=None):
    """The halo exchange of SEVERAL frames in ONE batched group (one NCCL launch instead of one per frame: the
    per-frame cost of the exchange is host/launch latency, not bytes).  `frames` is a list of (owned, window) pairs as
    for exchange_halo; every rank must pass its frames in the same order.  Returns the number of point-to-point ops."""
    own0, own1 = plan.owned_in_rows(rank)
    need0, need1 = plan.needed_in_rows(rank)
    a, b = max(own0, need0), min(own1, need1)
    sends, recvs = plan.transfers(rank)
    if dist is None:
        import torch.distributed as dist
    ops = []
    for owned, window in frames:
        if b > a and owned[a - own0:a - own0 + 1].data_ptr() != window[a - need0:a - need0 + 1].data_ptr():
            window[a - need0:b - need0].copy_(owned[a - own0:b - own0])
        for peer, r0, r1 in sends:
            ops.append(dist.P2POp(dist.isend, owned[r0 - own0:r1 - own0], peer))
        for peer, r0, r1 in recvs:
            ops.append(dist.P2POp(dist.irecv, window[r0 - need0:r1 - need0], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return len(ops)


class ShardedUpscaler:
    """One instance per rank (one process per GPU).

    The rank's input slab lives INSIDE its halo window: write frames into `self.owned` (a view of the window's
    middle rows) and call upscale(); only the 2-3 halo rows per side travel, straight into the window's edge rows.
    capture() records halo exchange + EASU + RCAS into one CUDA graph so a frame costs one graph launch.
    """

    def __init__(self, in_w, in_h, out_w, out_h, world, rank, sharpness=0.25, dtype=None, device=None, flags=0):
        import torch
        self.rank, self.world = rank, world
        self.in_w, self.in_h, self.out_w, self.out_h = in_w, in_h, out_w, out_h
        self.econ = api.easu_con(in_w, in_h, in_w, in_h, out_w, out_h)
        self.rcon = api.rcas_con(sharpness)
        self.plan = SlabPlan(in_h, out_h, world, self.econ)
        self.flags = flags
        dtype = dtype or torch.float16
        device = device or torch.device("cuda", torch.cuda.current_device())
        n0, n1 = self.plan.needed_in_rows(rank)
        o0, o1 = self.plan.owned_in_rows(rank)
        e0, e1 = self.plan.easu_rows(rank)
        y0, y1 = self.plan.out_rows(rank)
        lo, hi = min(n0, o0), max(n1, o1)     # the window always holds the whole owned slab
        self._win0 = lo
        self.window = torch.zeros((hi - lo, in_w, 4), dtype=dtype, device=device)
        self.owned = self.window[o0 - lo:o1 - lo]
        self.tmp = torch.empty((e1 - e0, out_w, 4), dtype=dtype, device=device)
        self.out = torch.empty((y1 - y0, out_w, 4), dtype=dtype, device=device)
        self._graph = None
        self._prepared = None
        import os
        self.halo_mode = os.environ.get("FSR1_HALO_MODE", "a2a")   # "p2p": batch_isend_irecv; "a2a": one all_to_all

    def _build_exchange(self):
        """Halo transfers as views of the window, built once: rows to send / receive per peer."""
        import torch
        plan, rank, w0 = self.plan, self.rank, self._win0
        sends, recvs = plan.transfers(rank)
        self._sends = [(peer, self.window[a - w0:b - w0]) for peer, a, b in sends]
        self._recvs = [(peer, self.window[a - w0:b - w0]) for peer, a, b in recvs]
        empty = self.window[0:0]
        # the same transfers phrased as ONE all-to-all whose only non-empty entries are the neighbours: a single
        # NCCL group / work object per frame instead of one per send and receive (host launch cost, not bytes)
        self._a2a_in = [empty] * self.world
        self._a2a_out = [empty] * self.world
        multi = False
        for peer, t in self._sends:
            multi |= self._a2a_in[peer].numel() > 0
            self._a2a_in[peer] = t
        for peer, t in self._recvs:
            multi |= self._a2a_out[peer].numel() > 0
            self._a2a_out[peer] = t
        self._a2a_ok = not multi

    def _exchange(self):
        if not hasattr(self, "_sends"):
            self._build_exchange()
        if not self._sends and not self._recvs:
            return
        import torch.distributed as dist
        if self.halo_mode == "a2a" and self._a2a_ok:
            dist.all_to_all(self._a2a_out, self._a2a_in)
            return
        ops = [dist.P2POp(dist.isend, t, peer) for peer, t in self._sends]
        ops += [dist.P2POp(dist.irecv, t, peer) for peer, t in self._recvs]
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    @staticmethod
    def exchange_many(upscalers):
        """One batched NCCL group carrying the halos of several frames (each frame = one ShardedUpscaler of the same
        geometry, e.g. the slots of a ring): amortises the per-exchange launch cost when frames are processed in groups."""
        import torch.distributed as dist
        ops = []
        for u in upscalers:
            if not hasattr(u, "_sends"):
                u._build_exchange()
            ops += [dist.P2POp(dist.isend, t, peer) for peer, t in u._sends]
            ops += [dist.P2POp(dist.irecv, t, peer) for peer, t in u._recvs]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return len(ops)

    def _launch(self, stream=None):
        if self._prepared is None:
            plan, rank = self.plan, self.rank
            e0, _ = plan.easu_rows(rank)
            y0, y1 = plan.out_rows(rank)
            self._prepared = api.PreparedUpscale(
                api.image(self.window, height=self.in_h, row0=self._win0), api.image(self.tmp, height=self.out_h, row0=e0),
                api.image(self.out, height=self.out_h, row0=y0), self.econ, self.rcon, y0=y0, y1=y1, flags=self.flags)
        self._prepared.launch(stream)

    def upscale(self, owned_rows=None, stream=None):
        """Upscale the frame whose slab is in self.owned (or in `owned_rows`, which is then copied in)."""
        if owned_rows is not None and owned_rows.data_ptr() != self.owned.data_ptr():
            self.owned.copy_(owned_rows)
        if self._graph is not None and stream is None:
            self._graph.replay()
            return self.out
        self._exchange()
        self._launch(stream)
        return self.out

    def capture(self):
        """Record halo exchange + both kernels into a CUDA graph (NCCL point-to-point is capturable)."""
        import torch
        self.upscale()                      # warm up: NCCL channels, kernel attributes, TMA descriptors
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._exchange()
            self._launch(torch.cuda.current_stream())
        self._graph = g
        return self
