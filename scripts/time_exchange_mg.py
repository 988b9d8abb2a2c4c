"""NVLink-side measurement of the fused exchange: 1 process x N GPUs, time dd.exchange() and report per-GPU
bytes that cross NVLink per second.  usage: time_exchange_mg.py [n per gpu] [radius] [nq] [dtype]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stencil_b200 as sb
from stencil_b200.jacobi import scaled_size

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
r = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dtype = np.dtype(sys.argv[4]) if len(sys.argv) > 4 else np.dtype("float32")
ng = torch.cuda.device_count()
X, Y, Z = scaled_size(n, n, n, ng)
dd = sb.DistributedDomain(X, Y, Z)
dd.set_gpus(list(range(ng)))
dd.set_radius(sb.Radius.constant(r))
for _ in range(nq):
    dd.add_data(dtype)
dd.realize()
for _ in range(5):
    dd.exchange()
reps = 30
ts = []
for _ in range(reps):
    for d in dd.domains():
        torch.cuda.synchronize(d.gpu())
    t0 = time.perf_counter()
    dd.exchange()
    ts.append(time.perf_counter() - t0)
ts = np.array(ts)
total = dd.exchange_bytes_for_method(sb.Method.Default)
peer = dd._bytes_peer
print(
    f"gpus={ng} dim={dd.partition_.dim} global={X}x{Y}x{Z} r={r} q={nq} {dtype}: exchange median {np.median(ts)*1e6:.1f} us min {ts.min()*1e6:.1f} us; "
    f"payload {total/1e6:.1f} MB total, {peer/1e6:.1f} MB over NVLink ({peer/ng/1e6:.1f} MB sent per GPU) -> "
    f"{peer/ng/np.median(ts)/1e9:.1f} GB/s per GPU per direction (staged {dd._staged_bytes/1e6:.1f} MB)",
    flush=True,
)
dd.close()
