"""One-process-per-GPU plumbing for DistributedDomain (torch.distributed = rendezvous only).

The reference reaches colocated ranks through CUDA IPC memory/event handles shipped over MPI
(include/stencil/tx_cuda.cuh:225-315, src/tx_ipc.cpp, src/tx_colocated.cu).  Here the handles travel
once, at realize(), through torch.distributed.all_gather_object; after that the data path is the
same fused kernel as in the single-process case, storing into IPC-mapped ghost cells over NVLink,
and completion is a pair of device-side flags per neighbour (ready / done) -- no host barrier and no
NCCL call inside exchange().

Protocol of exchange number e on rank A with neighbour ranks N(A):
    begin :  for B in N(A): ready[B][A] = e      (A's ghost cells may be overwritten for e)
             wait  ready[A][B] >= e for B in N(A)
    copy  :  fused halo write into every neighbour (one launch per local subdomain)
    finish:  for B in N(A): done[B][A] = e        (release, system scope, after the copy)
             wait  done[A][B] >= e for B in N(A)
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Tuple

from ._lib import Pitched, check, lib, stream_ptr

Vec = Tuple[int, int, int]


@dataclass
class World:
    rank: int
    size: int
    local_device: int


def world() -> World:
    try:
        import torch.distributed as td

        if td.is_available() and td.is_initialized():
            rank, size = td.get_rank(), td.get_world_size()
            local = int(os.environ.get("LOCAL_RANK", rank))
            return World(rank, size, local)
    except ImportError:
        pass
    return World(0, 1, 0)


def all_gather_object(obj):
    import torch.distributed as td

    out = [None] * td.get_world_size()
    td.all_gather_object(out, obj)
    return out


class RemoteDomains:
    """IPC mappings of every other rank's allocations + the flag mailboxes."""

    def __init__(self, dd, w: World):
        self.w = w
        self.dev = dd.domains_[0].gpu()
        self._opened: List[int] = []
        n = w.size
        # my mailbox: ready[0..n) then done[0..n)
        p = C.c_void_p()
        check(lib().sb_malloc(C.byref(p), 2 * n * 4, self.dev))
        check(lib().sb_memset(p, 0, 2 * n * 4, self.dev, None))
        check(lib().sb_device_sync(self.dev))
        self.flags = int(p.value)

        def export(ptr: int) -> bytes:
            buf = (C.c_char * 64)()
            check(lib().sb_ipc_export(C.c_void_p(ptr), buf))
            return bytes(buf)

        mine = {
            "rank": w.rank,
            "flags": export(self.flags),
            "domains": {
                dd.domain_idx_[i]: {
                    "raw": d.raw_size(),
                    "elem_sizes": list(d.elem_sizes_),
                    "curr": [export(p_) for p_ in d.curr_],
                    "next": [export(p_) for p_ in d.next_],
                }
                for i, d in enumerate(dd.domains_)
            },
        }
        everyone = all_gather_object(mine)

        def open_(handle: bytes) -> int:
            out = C.c_void_p()
            check(lib().sb_ipc_import(handle, self.dev, C.byref(out)))
            self._opened.append(int(out.value))
            return int(out.value)

        self.remote: Dict[Vec, dict] = {}
        self.peer_flags: Dict[int, int] = {}
        for info in everyone:
            r = info["rank"]
            if r == w.rank:
                continue
            self.peer_flags[r] = open_(info["flags"])
            for idx, dom in info["domains"].items():
                self.remote[tuple(idx)] = {
                    "raw": tuple(dom["raw"]),
                    "elem_sizes": dom["elem_sizes"],
                    "curr": [open_(h) for h in dom["curr"]],
                    "next": [open_(h) for h in dom["next"]],
                }
        # neighbour ranks: owners of any subdomain adjacent to one of mine (periodic, symmetric)
        from .domain import ALL_DIRS, get_neighbor

        nbrs = set()
        for idx in dd.domain_idx_:
            for d in ALL_DIRS:
                nbrs.add(dd._owner[get_neighbor(idx, d, dd.partition_.dim)][0])
        nbrs.discard(w.rank)
        self.nbrs = sorted(nbrs)
        k = len(self.nbrs)
        self._ready_remote = (C.c_void_p * max(k, 1))(*[self.peer_flags[b] + 4 * w.rank for b in self.nbrs])
        self._done_remote = (C.c_void_p * max(k, 1))(*[self.peer_flags[b] + 4 * (n + w.rank) for b in self.nbrs])
        # local slots to poll, gathered contiguously is not possible (slots are indexed by rank), so wait per slot list
        self._ready_local = [self.flags + 4 * b for b in self.nbrs]
        self._done_local = [self.flags + 4 * (n + b) for b in self.nbrs]

    def pitched(self, idx: Vec, q: int, parity: int) -> Tuple[Pitched, int]:
        dom = self.remote[tuple(idx)]
        base = dom["curr" if parity == 0 else "next"][q]
        return Pitched(base, dom["raw"][0] * dom["elem_sizes"][q], dom["raw"][1]), -1

    def _wait_all(self, slots: List[int], value: int, stream) -> None:
        # contiguous runs of slots are waited on by one kernel
        i = 0
        while i < len(slots):
            j = i
            while j + 1 < len(slots) and slots[j + 1] == slots[j] + 4:
                j += 1
            check(lib().sb_wait(C.c_void_p(slots[i]), j - i + 1, value, self.dev, stream_ptr(stream)))
            i = j + 1

    def begin(self, epoch: int, streams) -> None:
        if not self.nbrs:
            return
        s = streams[0]
        check(lib().sb_signal(self._ready_remote, len(self.nbrs), epoch, self.dev, stream_ptr(s)))
        self._wait_all(self._ready_local, epoch, s)
        # other local streams (several subdomains per rank) start after the handshake
        if len(streams) > 1:
            import torch

            ev = torch.cuda.Event()
            ev.record(s)
            for o in streams[1:]:
                o.wait_event(ev)

    def finish(self, epoch: int, streams) -> None:
        if not self.nbrs:
            return
        s = streams[0]
        if len(streams) > 1:
            import torch

            for o in streams[1:]:
                ev = torch.cuda.Event()
                ev.record(o)
                s.wait_event(ev)
        check(lib().sb_signal(self._done_remote, len(self.nbrs), epoch, self.dev, stream_ptr(s)))
        self._wait_all(self._done_local, epoch, s)

    def close(self) -> None:
        import torch.distributed as td

        # nobody may unmap while a peer could still be writing
        if td.is_initialized():
            td.barrier()
        for p in self._opened:
            lib().sb_ipc_close(C.c_void_p(p), self.dev)
        self._opened = []
        if td.is_initialized():
            td.barrier()
        if self.flags:
            lib().sb_free(C.c_void_p(self.flags), self.dev)
            self.flags = 0


def share_domains(dd, w: World) -> RemoteDomains:
    return RemoteDomains(dd, w)
