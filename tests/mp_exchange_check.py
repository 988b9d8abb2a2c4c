"""Run under torchrun (one rank per GPU): exchange parity of the one-process-per-GPU mode vs the oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mp_exchange_check.py
Set SB_FORCE_NCCL=1 to exercise the NCCL fallback instead of the CUDA-IPC direct write."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import stencil_b200 as sb  # noqa: E402
from gpu_util import oracle_radius  # noqa: E402
from oracle import np_oracle as no  # noqa: E402
from stencil_b200.jacobi import Jacobi3D, jacobi_radius  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = td.get_rank(), td.get_world_size()
    for size, radius, dtypes in [
        ((24, 20, 16), sb.Radius.constant(2), [np.float32, np.float64]),
        ((16, 18, 20), sb.Radius.face_edge_corner(3, 2, 1), [np.float64]),
        ((64, 64, 64), sb.Radius.constant(1), [np.float32] * 3),
    ]:
        dd = sb.DistributedDomain(*size)
        dd.set_gpus([local])
        dd.set_radius(radius)
        for dt in dtypes:
            dd.add_data(dt)
        dd.realize()
        ro = oracle_radius(radius)
        od = no.Domains(size, ro, dtypes, n_subdomains=world)
        for rep in range(3):
            fld = lambda q, x, y, z, rep=rep: no.hash_field(q, x, y, z) + rep  # noqa: E731
            od.fill(fld)
            for di, d in enumerate(dd.domains()):
                for q in range(len(dtypes)):
                    d.quantity_from_host(q, od.arrays[dd.domain_idx_[di]][q])
            td.barrier()
            dd.exchange()
            for di, d in enumerate(dd.domains()):
                for q in range(len(dtypes)):
                    want = od.expected_after_exchange(fld, dd.domain_idx_[di], q)
                    got = d.quantity_to_host(q)
                    assert np.array_equal(got, want), (rank, size, rep, q)
            dd.swap()
        dd.close()
    # jacobi through the multi-process path: 5 iterations, compare with a single-address-space oracle run
    from oracle import c_oracle as co
    from oracle import geometry as g

    n = 128  # 64 cells along x per rank: whole warp strips, so the fused schedule ships dense x columns between ranks
    dd = sb.DistributedDomain(n, n, n)
    dd.set_gpus([local])
    dd.set_radius(jacobi_radius())
    h = dd.add_data(np.float64)
    dd.realize()
    jac = Jacobi3D(dd, h)
    jac.init(0.5)
    # two host-synchronised iterations, then three queued back to back (events + device-side flags only)
    for _ in range(2):
        jac.step()
    fused = os.environ.get("SB_FORCE_NCCL") != "1"  # the fused schedule stores into peer memory
    for it in range(3):
        if fused and it != 1:
            jac.step_fused()  # update + halo push into the neighbour ranks' ghost cells, ordered by device-side counters
        else:
            jac.step_async()
    jac.synchronize()
    ro = g.Radius.face_edge_corner(1, 0, 0)
    od = no.Domains((n, n, n), ro, [np.float64], n_subdomains=world)
    nxt = {}
    for i in od.indices:
        od.arrays[i][0][...] = 0
        no.box(od.arrays[i][0], (1, 1, 1), od.sizes[i])[...] = 0.5
        nxt[i] = np.zeros_like(od.arrays[i][0])
    creg = ((0, 0, 0), (n, n, n))
    for _ in range(5):
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
    d = dd.domains()[0]
    i = dd.domain_idx_[0]
    got = d.quantity_to_host(0)
    assert np.array_equal(no.box(got, (1, 1, 1), od.sizes[i]), no.box(od.arrays[i][0], (1, 1, 1), od.sizes[i])), rank
    dd.close()
    td.barrier()
    if rank == 0:
        mode = "NCCL fallback" if os.environ.get("SB_FORCE_NCCL") == "1" else "CUDA-IPC direct write"
        print(f"mp_exchange_check OK: world={world} mode={mode}", flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
