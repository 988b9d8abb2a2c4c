"""The jacobi3d iteration of the reference driver (bin/jacobi3d.cu:296-377) over DistributedDomain.

    interior kernel on a compute stream  ||  dd.exchange() on the library's high-priority streams
    exterior slabs (after the exchange)  ->  stream sync  ->  dd.swap()

All ctypes argument packs are built once per swap parity so the per-iteration host work is a
handful of foreign calls.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from ._lib import HaloPush, Pitched, check, i3, lib, stream_ptr
from .domain import DataHandle, DistributedDomain, Radius


def jacobi_radius() -> Radius:
    """faces only, radius 1 (bin/jacobi3d.cu:237-246)"""
    return Radius.face_edge_corner(1, 0, 0)


def scaled_size(x: int, y: int, z: int, num_subdoms: int) -> Tuple[int, int, int]:
    """Weak-scaling size rule of bin/jacobi3d.cu:189-199 (one node): multiply the prime factors of the
    subdomain count into the currently smallest dimension."""
    from .domain import prime_factors

    for pf in prime_factors(num_subdoms):
        if x <= y and x <= z:
            x *= pf
        elif y <= z:
            y *= pf
        else:
            z *= pf
    return x, y, z


class Jacobi3D:
    def __init__(self, dd: DistributedDomain, h: DataHandle, overlap: bool = True):
        import torch

        self.dd, self.h, self.overlap = dd, h, overlap
        self.creg = dd.get_compute_region()
        self.streams = [torch.cuda.Stream(device=d.gpu()) for d in dd.domains()]
        # the exterior slabs only depend on the exchange, not on the interior kernel: they run on their
        # own stream so that they (and their launch latency) hide behind the interior kernel's tail
        self.ext_streams = [torch.cuda.Stream(device=d.gpu()) for d in dd.domains()]
        interiors, exteriors = dd.get_interior(), dd.get_exterior()
        L = lib()
        self._fn = L.sb_jacobi3d
        self._fn_regions = L.sb_jacobi3d_regions
        # calls[parity][domain] = (interior args, [exterior args...], whole args)
        self._calls = []
        for parity in (0, 1):
            per_dom = []
            for di, d in enumerate(dd.domains()):
                src = d.pitched(h.id, "curr" if parity == 0 else "next")
                dst = d.pitched(h.id, "next" if parity == 0 else "curr")
                if parity == 1:
                    # pitched() reads the CURRENT curr_/next_ lists; parity 1 = after one swap
                    pass
                acc = i3(d.accessor_origin())
                clo, chi = i3(self.creg[0]), i3(self.creg[1])
                s = stream_ptr(self.streams[di])

                def pack(reg):
                    return (dst, src, d.elem_size(h.id), acc, i3(reg[0]), i3(reg[1]), clo, chi, s)

                ne = len(exteriors[di])
                elo = (C.c_int64 * (3 * max(ne, 1)))(*[v for r in exteriors[di] for v in r[0]])
                ehi = (C.c_int64 * (3 * max(ne, 1)))(*[v for r in exteriors[di] for v in r[1]])
                ext_pack = (dst, src, d.elem_size(h.id), acc, ne, elo, ehi, clo, chi, stream_ptr(self.ext_streams[di]))
                per_dom.append((pack(interiors[di]), ext_pack, pack(d.get_compute_region())))
            self._calls.append(per_dom)
        # like the reference driver (d.set_device() before every launch, bin/jacobi3d.cu:310), the CUDA device
        # must be current when launching on one of its streams; only matters with several GPUs per process
        self._devs = [d.gpu() for d in dd.domains()]
        self._multi_dev = len(set(self._devs)) > 1
        self._set_device = torch.cuda.set_device
        import os

        dbg = os.environ.get("SB_DEBUG_SKIP", "")
        self._debug_skip = {"both": ("ext", "xchg")}.get(dbg, (dbg,) if dbg else ())
        self._ghosts_current = False  # step_fused: do the ghost cells of curr hold the neighbours' current values?
        self._ev_ext = None  # step_async: exterior-done events of the previous iteration (one per subdomain)
        self._ev_int = None
        self.interior_cells = sum(int(np.prod([hi[a] - lo[a] for a in range(3)])) for lo, hi in interiors)
        self._parity0 = dd._parity

    def _args(self):
        return self._calls[(self.dd._parity - self._parity0) & 1]

    def launch_interior(self) -> None:
        for dev, a in zip(self._devs, self._args()):
            if self._multi_dev:
                self._set_device(dev)
            check(self._fn(*a[0]))

    def launch_exterior(self) -> None:
        """All exterior slabs of a subdomain in ONE launch (the reference issues up to six)."""
        for dev, a in zip(self._devs, self._args()):
            if self._multi_dev:
                self._set_device(dev)
            check(self._fn_regions(*a[1]))

    def launch_whole(self) -> None:
        for dev, a in zip(self._devs, self._args()):
            if self._multi_dev:
                self._set_device(dev)
            check(self._fn(*a[2]))

    def step(self) -> None:
        """One iteration, exactly the loop body of bin/jacobi3d.cu:296-368."""
        dd = self.dd
        self._ghosts_current = False
        if self._ev_ext is not None or getattr(self, "_ev_fused", None) is not None:
            self.synchronize()  # drain iterations queued by step_async / step_fused
            self._ev_fused = None
        if self.overlap:
            self.launch_interior()
            dd.exchange()
            self.launch_exterior()
        else:
            dd.exchange()
            self.launch_whole()
        self.synchronize()
        self._ev_ext = self._ev_int = None
        dd.swap()

    def step_async(self, timing=None) -> None:
        """The same iteration with every dependency expressed as a CUDA event instead of a host-side stream
        synchronisation, so consecutive iterations queue back to back on the device (the reference blocks the host
        in exchange() and after the exterior kernels, bin/jacobi3d.cu:337-365).  Per subdomain, iteration i
        (curr = A, next = B):

            interior i   reads A, writes B interior      after exterior i-1 (wrote A's boundary, read B near it)
            exchange i   reads A boundary, writes ghosts  after exterior i-1 of every sending subdomain
            exterior i   reads A + ghosts, writes B rim   after exchange i, and after interior i-1 (read B's rim)

        Remote ranks are ordered by the device-side ready/done flags of the exchange itself (dist.RemoteDomains):
        `ready` is signalled on the exchange stream, i.e. after this rank's exterior i-1.  Results are bitwise
        those of step().  `timing` = (event, event) recorded around subdomain 0's interior kernel."""
        import torch

        if not self.overlap:
            raise RuntimeError("step_async needs the overlapped (interior/exterior) schedule")
        dd = self.dd
        self._ghosts_current = False
        if getattr(self, "_ev_fused", None) is not None:
            self.synchronize()
            self._ev_fused = None
        skip = self._debug_skip  # timing diagnostics only (SB_DEBUG_SKIP=ext|xchg|both): results are wrong when set
        nd = len(self._devs)
        prev_ext, prev_int = self._ev_ext, self._ev_int
        xs = dd.exchange_streams()
        ev_int = []
        for di, (dev, a) in enumerate(zip(self._devs, self._args())):
            if self._multi_dev:
                self._set_device(dev)
            s = self.streams[di]
            if prev_ext is not None:
                s.wait_event(prev_ext[di])
            if timing is not None and di == 0:
                timing[0].record(s)
            check(self._fn(*a[0]))
            if timing is not None and di == 0:
                timing[1].record(s)
            e = torch.cuda.Event()
            e.record(s)
            ev_int.append(e)
        if prev_ext is not None:
            for x in xs:
                for e in prev_ext:
                    x.wait_event(e)
        if "xchg" not in skip:
            dd.exchange_async()
        ev_x = []
        for di, x in enumerate(xs):
            if self._multi_dev:
                self._set_device(self._devs[di])
            e = torch.cuda.Event()
            e.record(x)
            ev_x.append(e)
        ev_ext = []
        for di, (dev, a) in enumerate(zip(self._devs, self._args())):
            if self._multi_dev:
                self._set_device(dev)
            s = self.ext_streams[di]
            for e in ev_x:
                s.wait_event(e)
            if prev_int is not None:
                s.wait_event(prev_int[di])
            if "ext" not in skip:
                check(self._fn_regions(*a[1]))
            e = torch.cuda.Event()
            e.record(s)
            ev_ext.append(e)
        self._ev_ext, self._ev_int = ev_ext, ev_int
        dd.swap()

    # ------------------------------------------------------------------ fused schedule
    def _build_fused(self) -> None:
        """Argument packs of sb_jacobi3d_fused per swap parity: the whole compute region + the six face neighbours'
        output allocations (own memory, a peer GPU of this process, or another rank's IPC mapping)."""
        import os

        from .domain import get_neighbor

        dd, h = self.dd, self.h
        r = dd.radius_
        for d6 in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)):
            if r.dir(d6) != 1:
                raise RuntimeError("the fused jacobi schedule needs face radius 1 (bin/jacobi3d.cu:237-246)")
        if getattr(dd, "_use_nccl", False):
            raise RuntimeError("the fused jacobi schedule stores into peer memory: not available on the NCCL fallback")
        dirs = ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1))
        # An x face is one 8-byte cell per row.  Pushed cell by cell into another GPU it costs 17 us per iteration through
        # in-process peer access (fine) but 78 us through CUDA-IPC mappings; collecting the column in a dense array and
        # shipping it between kernels, or leaving it to the copy engine between kernels, costs 30-70 us of serialised
        # small launches (all measured on 2 x B200, profiles/README.md section 6).  So when an x face crosses RANKS the
        # queued schedule (x faces staged by the copy engine, hidden behind the interior kernel) is the faster one, and
        # this schedule is not used.  One decision for the whole job: every rank evaluates the same partition.
        part = dd.partition_
        if any(dd._owner[tuple(i)][0] != dd._owner[tuple(get_neighbor(i, dv, part.dim))][0] for i in part.indices() for dv in dirs[:2]):
            raise RuntimeError("x faces cross ranks: the queued schedule is faster than 8-byte stores through CUDA-IPC mappings")
        self._fused_calls = []
        self._fused_nbr_slots = []  # in-process neighbours of each local subdomain (stream dependencies)
        for rel in (0, 1):
            absolute = (self._parity0 + rel) & 1
            per_dom = []
            for di, d in enumerate(dd.domains()):
                idx = tuple(dd.domain_idx_[di])
                src = d.pitched(h.id, "curr" if rel == 0 else "next")
                dst = d.pitched(h.id, "next" if rel == 0 else "curr")
                push = HaloPush()
                slots = set()
                for k, dv in enumerate(dirs):
                    nidx = get_neighbor(idx, dv, dd.partition_.dim)
                    # the neighbour's NEXT buffer at this parity is its curr buffer of the other parity
                    pn, _ = dd._pitched_of(nidx, h.id, absolute ^ 1)
                    rank, slot = dd._owner[tuple(nidx)]
                    if rank == dd._world.rank:
                        raw = dd.domains_[slot].raw_size()
                        slots.add(slot)
                    else:
                        raw = dd._remote.raw_of(nidx)
                    push.nbr[k] = Pitched(pn.ptr, pn.pitch, pn.ysize)
                    push.nbr_zsize[k] = raw[2]
                    if os.environ.get("SB_DEBUG_NOPUSH"):  # timing diagnostics only: results are wrong
                        push.nbr[k] = Pitched(None, 0, 0)
                creg = d.get_compute_region()
                clo, chi = i3(self.creg[0]), i3(self.creg[1])
                pack = (dst, src, d.elem_size(h.id), i3(d.accessor_origin()), i3(creg[0]), i3(creg[1]), clo, chi, push, stream_ptr(self.streams[di]))
                per_dom.append(pack)
                if rel == 0:
                    self._fused_nbr_slots.append(sorted(slots))
            self._fused_calls.append(per_dom)
        self._fn_fused = lib().sb_jacobi3d_fused
        self._fused_epoch = 0
        self._ev_fused = None
        self._ghosts_current = False

    def step_fused(self, timing=None) -> None:
        """One iteration as ONE kernel per subdomain: the jacobi update of the whole compute region, with every boundary
        cell also stored into the ghost cell of the face neighbour that reads it next iteration (sb_jacobi3d_fused).
        The halo exchange of bin/jacobi3d.cu:337 is thereby part of the previous iteration's kernel; what remains
        between iterations is ordering -- CUDA events between the subdomains of this process, one device-side
        counter per neighbour rank (dist.RemoteDomains.signal_step / wait_step).  Bitwise the results of step().
        The first fused iteration after construction (or after step / step_async) runs a regular exchange()."""
        import ctypes as C

        import torch

        if not hasattr(self, "_fused_calls"):
            try:
                self._build_fused()
                self.fused_supported = all(d.size()[0] >= 16 for d in self.dd.domains())
            except RuntimeError:
                self._fused_calls, self.fused_supported = [], False
        if not self.fused_supported:  # see _build_fused: the queued schedule computes the same iteration
            return self.step_async(timing=timing)
        dd = self.dd
        if self._ev_ext is not None:
            self.synchronize()
            self._ev_ext = self._ev_int = None
        if not self._ghosts_current:
            self.synchronize()
            dd.exchange()  # ghost cells of curr, once; afterwards every iteration leaves them filled for the next
            if dd._remote is not None:
                import torch.distributed as td

                td.barrier()  # every rank's exchange has landed before anybody's first fused kernel
            self._ghosts_current = True
            self._ev_fused = None
        calls = self._fused_calls[(dd._parity - self._parity0) & 1]
        prev = self._ev_fused
        remote = dd._remote
        events = []
        for di, (dev, a) in enumerate(zip(self._devs, calls)):
            if self._multi_dev:
                self._set_device(dev)
            s = self.streams[di]
            if prev is not None:
                for sj in self._fused_nbr_slots[di]:
                    if sj != di:
                        s.wait_event(prev[sj])
            if remote is not None and self._fused_epoch > 0:
                remote.wait_step(self._fused_epoch, s)
            if timing is not None and di == 0:
                timing[0].record(s)
            check(self._fn_fused(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], C.byref(a[8]), a[9]))
            if timing is not None and di == 0:
                timing[1].record(s)
            e = torch.cuda.Event()
            e.record(s)
            events.append(e)
        self._fused_epoch += 1
        if remote is not None:
            # one counter per rank: signal after ALL local subdomains of this iteration are done
            s0 = self.streams[0]
            for e in events[1:]:
                s0.wait_event(e)
            remote.signal_step(self._fused_epoch, s0)
        self._ev_fused = events
        dd.swap()

    def close(self) -> None:
        """Drain the queued work (everything else is owned by DistributedDomain.close())."""
        self.synchronize()

    def synchronize(self) -> None:
        """Wait for the compute streams (bin/jacobi3d.cu:363-365)."""
        for s in self.streams:
            s.synchronize()
        if self.overlap:
            for s in self.ext_streams:
                s.synchronize()

    def init(self, value: float = 0.5) -> None:
        """init_kernel (bin/jacobi3d.cu:18-29) on curr; ghost cells are filled by the first exchange."""
        from .domain import fill

        for d in self.dd.domains():
            fill(d, self.h, d.get_compute_region(), value)
            check(lib().sb_device_sync(d.gpu()))
