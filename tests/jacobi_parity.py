"""Bit-exact check of the jacobi iteration through DistributedDomain against the single-address-space CPU oracle.

Test infrastructure (it imports oracle/): used by tests/mp_exchange_check.py under torchrun, by tests/test_gpu_jacobi.py
and by bench.py AFTER its timed region (`parity_check` in the JSON line) -- the only way the driver's 1-GPU test box
ever sees the multi-rank path checked.  Mirrors the reference's exchange check of whole subdomains against a formula
evaluated in one address space (test/test_cuda_mpi_exchange.cu:193-245), applied to the jacobi loop of
bin/jacobi3d.cu:296-368 (which the reference does not test).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def oracle_jacobi(size, n_subdomains: int, dtype, iterations: int):
    """`iterations` of the reference loop on the whole periodic domain cut into n_subdomains (oracle partition =
    NodePartition); returns (Domains, {idx: array}) with the final curr arrays."""
    from oracle import c_oracle as co
    from oracle import geometry as g
    from oracle import np_oracle as no

    ro = g.Radius.face_edge_corner(1, 0, 0)
    od = no.Domains(tuple(size), ro, [dtype], n_subdomains=n_subdomains)
    nxt = {}
    for i in od.indices:
        od.arrays[i][0][...] = 0
        no.box(od.arrays[i][0], (1, 1, 1), od.sizes[i])[...] = 0.5
        nxt[i] = np.zeros_like(od.arrays[i][0])
    creg = ((0, 0, 0), tuple(size))
    for _ in range(iterations):
        for i in od.indices:
            lo = od.origins[i]
            hi = tuple(lo[a] + od.sizes[i][a] for a in range(3))
            co.jacobi_region(nxt[i], od.arrays[i][0], g.accessor_origin(lo, ro), *g.get_interior(lo, hi, ro), *creg)
        od.exchange()
        for i in od.indices:
            lo = od.origins[i]
            hi = tuple(lo[a] + od.sizes[i][a] for a in range(3))
            for elo, ehi in g.get_exterior(lo, hi, ro):
                co.jacobi_region(nxt[i], od.arrays[i][0], g.accessor_origin(lo, ro), elo, ehi, *creg)
            od.arrays[i][0], nxt[i] = nxt[i], od.arrays[i][0]
    return od


def check_jacobi_parity(size, gpus, dtype=np.float64, schedule=("fused",) * 6 + ("queued",) * 2, world: int = 1) -> dict:
    """Run len(schedule) iterations on a fresh DistributedDomain of global `size` over this process's `gpus`
    (world > 1: one rank per GPU, torch.distributed initialised) and compare every local subdomain's compute region
    bit for bit with the oracle.  Returns {"bit_exact", "ranks", "subdomains", "schedule", "mismatched_cells", ...};
    collective across ranks (all ranks get the global verdict)."""
    import stencil_b200 as sb
    from oracle import np_oracle as no
    from stencil_b200.jacobi import Jacobi3D, jacobi_radius

    dd = sb.DistributedDomain(*size)
    dd.set_gpus(list(gpus))
    dd.set_radius(jacobi_radius())
    h = dd.add_data(dtype)
    dd.realize()
    jac = Jacobi3D(dd, h)
    jac.init(0.5)
    ran = []
    for kind in schedule:
        if kind == "fused":
            jac.step_fused()
            ran.append("fused" if getattr(jac, "fused_supported", False) else "queued")
        elif kind == "queued":
            jac.step_async()
            ran.append("queued")
        else:
            jac.step()
            ran.append("host-sync")
    jac.synchronize()
    n_sub = len(dd.domains()) * world
    od = oracle_jacobi(size, n_sub, dtype, len(schedule))
    bad = 0
    for di, d in enumerate(dd.domains()):
        i = dd.domain_idx_[di]
        got = no.box(d.quantity_to_host(h.id), (1, 1, 1), od.sizes[i])
        want = no.box(od.arrays[i][0], (1, 1, 1), od.sizes[i])
        bad += int(np.count_nonzero(got != want))
    part = tuple(dd.partition_.dim)
    inkernel = getattr(jac, "_sync", None) is not None
    jac.close()
    dd.close()
    if world > 1:
        import torch
        import torch.distributed as td

        t = torch.tensor([bad], dtype=torch.int64, device="cuda")
        td.all_reduce(t)
        bad = int(t[0])
    # "fused" if every fused request ran fused
    kinds = sorted(set(ran))
    return {
        "bit_exact": bad == 0,
        "mismatched_cells": bad,
        "ranks": world,
        "subdomains": n_sub,
        "partition": list(part),
        "global_size": list(size),
        "dtype": np.dtype(dtype).name,
        "iterations": len(schedule),
        "schedule": "+".join(f"{ran.count(k)}x{k}" for k in kinds),
        "in_kernel_handshake": bool(inkernel),
        "oracle": "oracle/stencil_oracle.c in one address space (bin/jacobi3d.cu:296-368 restated)",
    }
