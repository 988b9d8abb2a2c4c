"""GPU parity of DistributedDomain.exchange()/swap(): the reference's own exchange check
(test/test_exchange.cu:37-220 -- whole allocation incl. ghosts equals the periodically wrapped field),
bit-exact, over the reference's radius list plus the cases its tests leave unpinned (SURVEY.md 8c:
exchange after swap, several quantities, radius 3, uneven partitions)."""
import numpy as np
import pytest

import stencil_b200 as sb
from oracle import geometry as g
from oracle import np_oracle as no
from gpu_util import oracle_radius

pytestmark = pytest.mark.gpu


def make_radius(name):
    if name.startswith("c"):
        return sb.Radius.constant(int(name[1:]))
    r = sb.Radius.constant(0)
    if name == "px2":
        r.set_dir((1, 0, 0), 2)
    elif name == "mx1":
        r.set_dir((-1, 0, 0), 1)
    elif name == "px2mx1":
        r.set_dir((1, 0, 0), 2)
        r.set_dir((-1, 0, 0), 1)
    elif name == "faces1":
        r = sb.Radius.face_edge_corner(1, 0, 0)
    elif name == "f2e1":
        r = sb.Radius.face_edge_corner(2, 1, 0)
    elif name == "f3e2c1":
        r = sb.Radius.face_edge_corner(3, 2, 1)
    else:
        raise KeyError(name)
    return r


def run_exchange_check(size, radius, gpus, dtypes, field, n_exchanges=1, swap_between=False):
    dd = sb.DistributedDomain(*size)
    dd.set_gpus(gpus)
    dd.set_radius(radius)
    handles = [dd.add_data(dt, f"d{i}") for i, dt in enumerate(dtypes)]
    dd.set_methods(sb.Method.Default)
    dd.realize()
    ro = oracle_radius(radius)
    odoms = no.Domains(size, ro, dtypes, n_subdomains=len(gpus))
    assert [d.size() for d in dd.domains()] == [odoms.sizes[i] for i in dd.domain_idx_]
    assert [d.origin() for d in dd.domains()] == [odoms.origins[i] for i in dd.domain_idx_]
    try:
        for rep in range(n_exchanges):
            fld = (lambda q, x, y, z, rep=rep: field(q, x, y, z) + 1000 * rep)
            odoms.fill(fld)
            for di, d in enumerate(dd.domains()):
                for q in range(len(dtypes)):
                    d.quantity_from_host(q, odoms.arrays[dd.domain_idx_[di]][q])
            dd.exchange()
            for di, d in enumerate(dd.domains()):
                idx = dd.domain_idx_[di]
                for q in range(len(dtypes)):
                    got = d.quantity_to_host(q)
                    want = odoms.expected_after_exchange(fld, idx, q)
                    assert np.array_equal(got, want), (rep, idx, q)
            if swap_between:
                dd.swap()
        moved = odoms.exchange()
        assert dd.exchange_bytes_for_method(sb.Method.Default) == moved
    finally:
        dd.close()


REF_RADII = ["c0", "c1", "c2", "px2", "mx1", "px2mx1"]  # test/test_exchange.cu:194-219


@pytest.mark.parametrize("rname", REF_RADII)
def test_exchange_reference_cases(rname):
    """DistributedDomain(10,10,10), set_gpus({0,0}) (two subdomains on GPU 0), one float quantity."""
    run_exchange_check((10, 10, 10), make_radius(rname), [0, 0], [np.float32], no.ripple_field)


@pytest.mark.parametrize("rname", ["c1", "c3", "faces1", "f2e1", "f3e2c1", "px2mx1"])
@pytest.mark.parametrize("ndom", [1, 4, 8])
def test_exchange_more_subdomains_and_quantities(rname, ndom):
    run_exchange_check((12, 10, 14), make_radius(rname), [0] * ndom, [np.float32, np.float64, np.int8], no.hash_field)


def test_exchange_after_swap_uses_the_new_curr():
    """Unpinned in the reference (no test calls exchange() after swap(); its captured pack graphs keep the
    old pointers).  Three exchange/swap rounds must each see the freshly written curr."""
    run_exchange_check((16, 12, 10), make_radius("c2"), [0, 0], [np.float64], no.ripple_field, n_exchanges=3, swap_between=True)


def test_exchange_uneven_partition_coordinate_field():
    """13x7x9 over 6 subdomains (uneven) with the coordinate-pack field of
    test/test_cuda_mpi_distributed_domain.cu:196-209."""
    run_exchange_check((13, 7, 9), make_radius("c2"), [0] * 6, [np.int32], no.coord_field)


def test_exchange_medium_128_r2_3q():
    """bench_exchange's shape (3 float quantities, uniform r=2) at 128^3 over 8 subdomains."""
    run_exchange_check((128, 128, 128), make_radius("c2"), [0] * 8, [np.float32] * 3, no.hash_field)


def test_interior_exterior_match_oracle():
    dd = sb.DistributedDomain(64, 48, 40)
    dd.set_gpus([0, 0, 0, 0])
    r = sb.Radius.face_edge_corner(2, 1, 0)
    dd.set_radius(r)
    dd.add_data(np.float32)
    dd.realize()
    try:
        ro = oracle_radius(r)
        for d, intr, ext in zip(dd.domains(), dd.get_interior(), dd.get_exterior()):
            lo, hi = d.get_compute_region()
            assert intr == g.get_interior(lo, hi, ro)
            assert ext == g.get_exterior(lo, hi, ro)
    finally:
        dd.close()


def test_multi_gpu_in_process_if_available():
    """1 process x N GPUs (the reference's 1 rank x N GPUs mode): peer-mapped direct writes."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    run_exchange_check((32, 24, 20), make_radius("c2"), list(range(n)), [np.float32, np.float64], no.hash_field, n_exchanges=2, swap_between=True)


@pytest.mark.parametrize("mode", ["ipc", "nccl"])
def test_one_process_per_gpu_if_available(mode):
    """torchrun with one rank per GPU: CUDA-IPC direct write (default) and the NCCL fallback."""
    import os
    import subprocess
    import sys

    import torch

    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SB_FORCE_NCCL="1" if mode == "nccl" else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if mode == "ipc" else "29534", os.path.join(root, "tests", "mp_exchange_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "mp_exchange_check OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
