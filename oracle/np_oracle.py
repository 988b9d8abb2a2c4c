"""numpy restatement of the data-movement kernels and the jacobi step -- CPU oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Arrays are indexed ``a[z, y, x]`` (x fastest), exactly the reference's unpitched
layout (src/local_domain.cu:187-203: pitch == xsize == width bytes).  A packed
buffer is the C-order ravel of the sub-box, i.e. ``zo*ey*ex + yo*ex + xo``
(src/pack_kernel.cu:35).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import geometry as g

Vec = Tuple[int, int, int]


def box(a: np.ndarray, pos: Vec, ext: Vec) -> np.ndarray:
    """View of the 3-D region [pos, pos+ext) of an allocation (x,y,z order of pos/ext)."""
    return a[pos[2] : pos[2] + ext[2], pos[1] : pos[1] + ext[1], pos[0] : pos[0] + ext[0]]


# --------------------------------------------------------------------------- pack / unpack / translate
def pack(a: np.ndarray, pos: Vec, ext: Vec) -> np.ndarray:
    """grid_pack / pack_kernel, src/pack_kernel.cu:3-59: strided -> contiguous, x fastest."""
    return np.ascontiguousarray(box(a, pos, ext)).reshape(-1)


def unpack(a: np.ndarray, buf: np.ndarray, pos: Vec, ext: Vec) -> None:
    """grid_unpack / unpack_kernel, src/pack_kernel.cu:61-108."""
    box(a, pos, ext)[...] = buf.reshape(ext[2], ext[1], ext[0])


def translate(dst: np.ndarray, dst_pos: Vec, src: np.ndarray, src_pos: Vec, ext: Vec) -> None:
    """translate / translate_grid, src/copy.cu:34-77: strided -> strided."""
    box(dst, dst_pos, ext)[...] = box(src, src_pos, ext)


def packer_pack(arrays: Sequence[np.ndarray], sz: Vec, radius: g.Radius, msgs) -> np.ndarray:
    """DevicePacker::pack over all quantities and messages -> one byte buffer in wire order
    (src/packer.cu:10-26, 109-148).  ``arrays[q]`` is quantity q's allocation."""
    es = [a.dtype.itemsize for a in arrays]
    total, entries = g.packer_layout(sz, radius, es, msgs)
    buf = np.zeros(total, dtype=np.uint8)
    for e in entries:
        data = pack(arrays[e["q"]], e["pos"], e["ext"])
        buf[e["offset"] : e["offset"] + e["nbytes"]] = data.view(np.uint8)
    return buf


def packer_unpack(arrays: Sequence[np.ndarray], buf: np.ndarray, sz: Vec, radius: g.Radius, msgs) -> None:
    """DeviceUnpacker::unpack (src/packer.cu:28-44, 210-253)."""
    es = [a.dtype.itemsize for a in arrays]
    total, entries = g.unpacker_layout(sz, radius, es, msgs)
    assert total == buf.size
    for e in entries:
        a = arrays[e["q"]]
        data = buf[e["offset"] : e["offset"] + e["nbytes"]].view(a.dtype)
        unpack(a, data, e["pos"], e["ext"])


# --------------------------------------------------------------------------- distributed exchange in one address space
class Domains:
    """N virtual subdomains of a periodic global box, each with Q quantities (curr only).

    Mirrors what DistributedDomain::realize builds (src/stencil.cu:241-268) without any GPU:
    ``arrays[idx][q]`` has shape raw_size[::-1].
    """

    def __init__(self, size: Vec, radius: g.Radius, dtypes: Sequence, n_subdomains: int = 1, partition: str = "node"):
        self.size = tuple(size)
        self.radius = radius
        self.dtypes = [np.dtype(d) for d in dtypes]
        if partition == "node":
            self.part = g.NodePartition(self.size, radius, 1, n_subdomains)
        else:
            self.part = g.RankPartition(self.size, n_subdomains)
        self.dim = self.part.dim()
        self.indices = self.part.all_indices()
        self.sizes: Dict[Vec, Vec] = {i: self.part.subdomain_size(i) for i in self.indices}
        self.origins: Dict[Vec, Vec] = {i: self.part.subdomain_origin(i) for i in self.indices}
        self.arrays: Dict[Vec, List[np.ndarray]] = {}
        for i in self.indices:
            raw = g.raw_size(self.sizes[i], radius)
            self.arrays[i] = [np.zeros(raw[::-1], dtype=dt) for dt in self.dtypes]

    def fill(self, fn) -> None:
        """Set every COMPUTE cell of every quantity to fn(q, X, Y, Z) (global coordinate arrays),
        halos to a poison value."""
        for i in self.indices:
            sz, org = self.sizes[i], self.origins[i]
            rm = tuple(self.radius.axis(ax, -1) for ax in range(3))
            zz, yy, xx = np.meshgrid(
                np.arange(org[2], org[2] + sz[2]),
                np.arange(org[1], org[1] + sz[1]),
                np.arange(org[0], org[0] + sz[0]),
                indexing="ij",
            )
            for q, a in enumerate(self.arrays[i]):
                a[...] = poison(a.dtype)
                box(a, rm, sz)[...] = fn(q, xx, yy, zz).astype(a.dtype)

    def exchange(self) -> int:
        """One halo exchange of all quantities (DistributedDomain::exchange, src/stencil.cu:1002-1186,
        with the plan of :327-412).  Returns the number of payload bytes moved."""
        moved = 0
        for m in g.plan_sends(self.dim, self.sizes, self.radius):
            for q in range(len(self.dtypes)):
                translate(self.arrays[m["dst_idx"]][q], m["dst_pos"], self.arrays[m["src_idx"]][q], m["src_pos"], m["ext"])
                moved += self.dtypes[q].itemsize * g.flatten(m["ext"])
        return moved

    def expected_after_exchange(self, fn, idx: Vec, q: int) -> np.ndarray:
        """What quantity q of subdomain idx must hold after an exchange if the compute cells were
        fn(q, x, y, z): every cell a message wrote equals fn at the periodically wrapped global
        coordinate (the check of test/test_exchange.cu:153-187); cells no message covers (e.g. corners
        when only faces have a radius) keep the poison."""
        sz, org = self.sizes[idx], self.origins[idx]
        raw = g.raw_size(sz, self.radius)
        ao = g.accessor_origin(org, self.radius)
        zz, yy, xx = np.meshgrid(
            np.arange(ao[2], ao[2] + raw[2]) % self.size[2],
            np.arange(ao[1], ao[1] + raw[1]) % self.size[1],
            np.arange(ao[0], ao[0] + raw[0]) % self.size[0],
            indexing="ij",
        )
        full = fn(q, xx, yy, zz).astype(self.dtypes[q])
        out = np.full(raw[::-1], poison(self.dtypes[q]), dtype=self.dtypes[q])
        rm = tuple(self.radius.axis(ax, -1) for ax in range(3))
        box(out, rm, sz)[...] = box(full, rm, sz)
        for d in g.ALL_DIRS:
            # we receive from direction -d what the neighbour sent in direction d
            if self.radius.dir(g.neg(d)) == 0:
                continue
            pos = g.halo_pos(g.neg(d), sz, self.radius, True)
            ext = g.halo_extent(g.neg(d), sz, self.radius)
            if g.flatten(ext):
                box(out, pos, ext)[...] = box(full, pos, ext)
        return out


def poison(dtype) -> object:
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        return dtype.type(-777.5)
    return dtype.type(0x5A if dtype.itemsize == 1 else 0x5A5A)


# --------------------------------------------------------------------------- synthetic fields
RIPPLE = (0.0, 0.25, 0.0, -0.25)


def ripple_field(q, x, y, z):
    """test/test_exchange.cu:16-28 -- v = x + r[x%4] + y + r[y%4] + z + r[z%4] (+ q so quantities differ)."""
    r = np.asarray(RIPPLE)
    return x + r[x % 4] + y + r[y % 4] + z + r[z % 4] + q


def coord_field(q, x, y, z):
    """test/test_cuda_mpi_distributed_domain.cu:196-209 -- x | y<<10 | z<<20 (as a number)."""
    return (x | (y << 10) | (z << 20)) + q


def _hash64(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64)
    with np.errstate(over="ignore"):
        v = (v ^ (v >> np.uint64(33))) * np.uint64(0xFF51AFD7ED558CCD)
        v = (v ^ (v >> np.uint64(33))) * np.uint64(0xC4CEB9FE1A85EC53)
        v = v ^ (v >> np.uint64(33))
    return v


def hash_field(q, x, y, z):
    """A bit-mixing fill in the spirit of astaroth's hash init (astaroth/astaroth.cu:84-110); used for
    multi-quantity bit-exact checks.  Returns small exactly-representable integers (< 2^23)."""
    h = _hash64(x + 1) ^ _hash64((y + 1) * 1000003) ^ _hash64((z + 1) * 998244353) ^ _hash64(np.uint64(q + 7))
    return (h & np.uint64(0x7FFFFF)).astype(np.int64)


# --------------------------------------------------------------------------- jacobi
def sphere_masks(xx, yy, zz, creg_lo: Vec, creg_hi: Vec):
    """Hot / cold sphere membership, bin/jacobi3d.cu:31-33, 46-63.

    centre hot  = (lo.x + (hi.x-lo.x)/3,   (lo.y+hi.y)/2, (lo.z+hi.z)/2)
    centre cold = (lo.x + (hi.x-lo.x)*2/3, ...), radius = (hi.x-lo.x)/10 (int division),
    member iff int64(sqrtf(float(d2))) <= radius.  The cold test only applies where hot is false.
    """
    ex = creg_hi[0] - creg_lo[0]
    hot = (creg_lo[0] + ex // 3, (creg_lo[1] + creg_hi[1]) // 2, (creg_lo[2] + creg_hi[2]) // 2)
    cold = (creg_lo[0] + ex * 2 // 3, hot[1], hot[2])
    rad = ex // 10

    def member(c):
        d2 = (xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2
        return np.sqrt(d2.astype(np.float32)).astype(np.int64) <= rad

    h = member(hot)
    c = member(cold) & ~h
    return h, c


def jacobi_region(dst: np.ndarray, src: np.ndarray, acc_origin: Vec, reg_lo: Vec, reg_hi: Vec, creg_lo: Vec, creg_hi: Vec) -> None:
    """stencil_kernel over region [reg_lo, reg_hi) (global coords), bin/jacobi3d.cu:40-85.

    ``src``/``dst`` are whole allocations; ``acc_origin`` is the global coordinate of element
    [0,0,0] (Accessor origin).  Sum order is ((((((0+px)+mx)+py)+my)+pz)+mz)/6 in the array's own
    precision; IEEE division (the reference's --use_fast_math approximate divide is NOT reproduced:
    documented in DESIGN.md).
    """
    if any(reg_hi[ax] <= reg_lo[ax] for ax in range(3)):
        return
    T = src.dtype.type
    lo = tuple(reg_lo[ax] - acc_origin[ax] for ax in range(3))
    ext = tuple(reg_hi[ax] - reg_lo[ax] for ax in range(3))

    def sh(dx, dy, dz):
        return box(src, (lo[0] + dx, lo[1] + dy, lo[2] + dz), ext)

    val = T(0) + sh(1, 0, 0)
    val = val + sh(-1, 0, 0)
    val = val + sh(0, 1, 0)
    val = val + sh(0, -1, 0)
    val = val + sh(0, 0, 1)
    val = val + sh(0, 0, -1)
    val = val / T(6)
    zz, yy, xx = np.meshgrid(
        np.arange(reg_lo[2], reg_hi[2]), np.arange(reg_lo[1], reg_hi[1]), np.arange(reg_lo[0], reg_hi[0]), indexing="ij"
    )
    h, c = sphere_masks(xx, yy, zz, creg_lo, creg_hi)
    val = np.where(h, T(1), np.where(c, T(0), val)).astype(src.dtype)
    box(dst, lo, ext)[...] = val


def residual_l2(a: np.ndarray, b: np.ndarray, pos: Vec, ext: Vec) -> float:
    """||a - b||_2 over a box, accumulated in float64 (new requirement)."""
    d = box(a, pos, ext).astype(np.float64) - box(b, pos, ext).astype(np.float64)
    return float(np.sqrt(np.sum(d * d)))
