"""world_size-2 and -4 `gloo` tests (CPU) of the one-process-per-GPU host logic: rank -> subdomain ownership,
the per-rank send plans, and that executing all ranks' plans reproduces the reference's exchange check
(test/test_exchange.cu:153-187).  Data is moved by the numpy oracle standing in for the GPU kernel; the
planning under test is the product's (stencil_b200.DistributedDomain.plan_messages over the C ABI)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, size, rname, dtypes, out_q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        td.init_process_group("gloo", rank=rank, world_size=world)
        import stencil_b200 as sb
        from gpu_util import oracle_radius
        from oracle import np_oracle as no
        from stencil_b200 import dist

        radius = {"c1": sb.Radius.constant(1), "c2": sb.Radius.constant(2), "f2e1": sb.Radius.face_edge_corner(2, 1, 0)}[rname]
        w = dist.world()
        assert (w.rank, w.size, w.local_device) == (rank, world, rank)
        dd = sb.DistributedDomain(*size)
        dd.set_gpus([0])  # one subdomain per rank
        dd.set_radius(radius)
        for dt in dtypes:
            dd.add_data(dt)
        dd.do_placement()  # no GPU needed: partition + ownership only
        msgs = dd.plan_messages()

        ro = oracle_radius(radius)
        odoms = no.Domains(size, ro, dtypes, n_subdomains=world)
        mine = [i for i in odoms.indices if dd._owner[i][0] == rank]
        assert len(mine) == 1 and all(m["src_idx"] == mine[0] for m in msgs)
        # the product's partition agrees with the oracle's
        assert dd.partition_.subdomain_size(mine[0]) == odoms.sizes[mine[0]]
        assert dd.partition_.subdomain_origin(mine[0]) == odoms.origins[mine[0]]
        odoms.fill(no.hash_field)
        # "send": pack my regions, ship them to everyone, receivers write what is addressed to them
        outbox = []
        for m in msgs:
            for q in range(len(dtypes)):
                outbox.append((m["dst_idx"], m["dst_rank"], q, m["dst_pos"], m["ext"], no.pack(odoms.arrays[mine[0]][q], m["src_pos"], m["ext"])))
        everyone = [None] * world
        td.all_gather_object(everyone, outbox)
        for box in everyone:
            for dst_idx, dst_rank, q, dst_pos, ext, payload in box:
                if dst_rank == rank:
                    assert tuple(dst_idx) == mine[0]
                    no.unpack(odoms.arrays[mine[0]][q], payload, dst_pos, ext)
        for q in range(len(dtypes)):
            want = odoms.expected_after_exchange(no.hash_field, mine[0], q)
            assert np.array_equal(odoms.arrays[mine[0]][q], want), (rank, q)
        # neighbour rank set is symmetric
        nb = sorted({m["dst_rank"] for m in msgs} - {rank})
        allnb = [None] * world
        td.all_gather_object(allnb, nb)
        for r, lst in enumerate(allnb):
            for o in lst:
                assert r in allnb[o]
        td.barrier()
        td.destroy_process_group()
        out_q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        out_q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.parametrize("world,size,rname", [(2, (10, 10, 10), "c1"), (2, (13, 9, 8), "c2"), (4, (12, 10, 14), "f2e1"), (4, (16, 16, 16), "c2")])
def test_multi_process_plans_over_gloo(world, size, rname):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, size, rname, [np.float32, np.float64], q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, "ok") for r in range(world)], results
