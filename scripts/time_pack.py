"""Time pack / unpack / self-exchange plans (CUDA events). usage: time_pack.py [n] [r] [dtype] [nq]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stencil_b200 as sb
from stencil_b200._lib import Pitched, check, lib
import ctypes as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
r = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dtype = np.dtype(sys.argv[3]) if len(sys.argv) > 3 else np.dtype("float32")
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 1
es = dtype.itemsize
radius = sb.Radius.constant(r)
d = sb.LocalDomain((n, n, n), (0, 0, 0), 0)
d.set_radius(radius)
for _ in range(nq):
    d.add_data(dtype)
d.realize()
s = torch.cuda.Stream()
tag = os.environ.get("SB_LIB_PATH", "default").split("/")[-1]


def timeit(plan, reps=30):
    for _ in range(3):
        plan.launch(s)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps):
        plan.launch(s)
    b.record(s)
    s.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for dirv in [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 1, 1)]:
    ext = d.halo_extent(tuple(-c for c in dirv))
    nbytes = ext[0] * ext[1] * ext[2] * es * nq
    buf = C.c_void_p()
    check(lib().sb_malloc(C.byref(buf), nbytes, 0))
    pk, up = [], []
    off = 0
    for q in range(nq):
        dense = Pitched(buf.value + off, ext[0] * es, ext[1])
        pk.append(sb.box_copy(dense, (0, 0, 0), d.curr_data(q), d.halo_pos(dirv, False), ext, es))
        up.append(sb.box_copy(d.curr_data(q), d.halo_pos(tuple(-c for c in dirv), True), dense, (0, 0, 0), ext, es))
        off += ext[0] * ext[1] * ext[2] * es
    tr = [sb.box_copy(d.curr_data(q), d.halo_pos(tuple(-c for c in dirv), True), d.curr_data(q), d.halo_pos(dirv, False), ext, es) for q in range(nq)]
    tp, tu, tt = timeit(sb.CopyPlan(0, pk)), timeit(sb.CopyPlan(0, up)), timeit(sb.CopyPlan(0, tr))
    print(f"[{tag}] n={n} r={r} {dtype} q={nq} dir={dirv} {nbytes/1e6:.2f} MB: pack {tp:.1f} us  unpack {tu:.1f} us  direct {tt:.1f} us ({nbytes/tt/1e3:.0f} GB/s payload)", flush=True)
