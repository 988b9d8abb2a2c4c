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
    # jacobi through the multi-process path, compared with a single-address-space oracle run (tests/jacobi_parity.py).
    # 64-cell multiples along x per rank = whole warp strips, so the fused schedule ships dense x columns between ranks.
    # The global shapes make the partitioner cut x, then y, then z (and all of them with more ranks), so every face
    # direction crosses ranks in some case.
    from jacobi_parity import check_jacobi_parity

    fused = os.environ.get("SB_FORCE_NCCL") != "1"  # the fused schedule stores into peer memory
    mixed = ("host-sync", "host-sync", "fused", "queued", "fused", "fused", "fused", "queued", "fused")
    shapes = {
        2: [(256, 128, 128), (128, 192, 128), (64, 128, 256)],  # cut along x / y / z
        4: [(512, 128, 128), (256, 256, 128), (64, 256, 256)],  # (4,1,1), (2,2,1), (1,2,2)
        8: [(256, 256, 256), (1024, 128, 128)],  # (2,2,2), (8,1,1)
    }.get(world, [(128 * world, 128, 128)])
    for size in shapes:
        for dt in (np.float64, np.float32):
            res = check_jacobi_parity(size, [local], dt, mixed if fused else ("host-sync", "queued", "queued"), world)
            if rank == 0:
                print("jacobi parity", res, flush=True)
            assert res["bit_exact"], res
    td.barrier()
    if rank == 0:
        mode = "NCCL fallback" if os.environ.get("SB_FORCE_NCCL") == "1" else "CUDA-IPC direct write"
        print(f"mp_exchange_check OK: world={world} mode={mode}", flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
