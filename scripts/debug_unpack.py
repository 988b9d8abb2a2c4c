import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import DevArray
from stencil_b200._lib import check, i3, lib
from oracle import np_oracle as no
for dtype in (np.float32, np.int16):
    raw, pos, ext = (70, 70, 70), (67, 3, 3), (3, 64, 64)
    rng = np.random.default_rng(7)
    a = rng.integers(0, 250, size=raw).astype(dtype)
    packed = no.pack(a, pos, ext)
    dst = DevArray(np.full(raw, 3, dtype=dtype))
    dbuf = DevArray(packed.reshape(1, 1, -1))
    check(lib().sb_unpack(dst.pitched(), C.c_void_p(dbuf.ptr), i3(pos), i3(ext), a.dtype.itemsize, None))
    want = np.full(raw, 3, dtype=dtype); no.unpack(want, packed, pos, ext)
    got = dst.get()
    bad = np.argwhere(got != want)
    print(dtype.__name__, "mismatches", len(bad), bad[:10].tolist(), [ (int(got[tuple(b)]), int(want[tuple(b)])) for b in bad[:10]])
