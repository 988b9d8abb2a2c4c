"""Host-side logic of bench.py that decides WHAT is measured (no GPU): the weak-scaling shapes and how the partitioner
cuts them.  bin/jacobi3d.cu:189-199 is the size rule, partition.hpp:157-255 the partitioner (oracle/geometry.py)."""
import importlib.util
import os

import pytest

from oracle import geometry as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("n", [64, 128, 512])
@pytest.mark.parametrize("ngpu", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("grow", ["yz", "cube", "x"])
def test_every_gpu_gets_an_n_cubed_subdomain(bench, n, ngpu, grow):
    size = bench.grown_size(n, ngpu, grow)
    assert size[0] * size[1] * size[2] == ngpu * n**3
    part = g.NodePartition(size, g.Radius.constant(1), 1, ngpu)
    dim = part.dim()
    assert dim[0] * dim[1] * dim[2] == ngpu
    if grow == "yz":
        assert size[0] == n and dim[0] == 1  # x, the contiguous axis, is never grown and never cut
    sizes = {tuple(part.subdomain_size(i)) for i in part.all_indices()}
    if ngpu in (1, 2, 4, 8):
        assert sizes == {(n, n, n)}


def test_reference_order_is_the_size_rule_of_the_driver(bench):
    # bin/jacobi3d.cu:189-199 restated in oracle/geometry.py
    for ngpu in (1, 2, 4, 8):
        assert tuple(bench.grown_size(512, ngpu, "x")) == tuple(g.jacobi_scaled_size(512, 512, 512, ngpu))
    assert tuple(bench.grown_size(512, 8, "yz")) == (512, 1024, 2048)
    assert tuple(bench.grown_size(512, 4, "yz")) == (512, 1024, 1024)
    assert tuple(bench.grown_size(512, 2, "yz")) == (512, 512, 1024)
    assert tuple(bench.grown_size(512, 2, "cube")) == (512, 512, 1024)
    assert tuple(bench.grown_size(512, 8, "cube")) == (1024, 1024, 1024)
