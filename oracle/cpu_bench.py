#!/usr/bin/env python
"""CPU arm of bench.py: the oracle port of the jacobi3d loop (bin/jacobi3d.cu:296-368) timed on the host cores.

TEST / BASELINE INFRASTRUCTURE (it is the oracle): run as its OWN PROCESS by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference`, so that the OpenMP runtime starts with a known environment whatever the parent process
imported before (torch ships its own OpenMP runtime; torchrun exports OMP_NUM_THREADS=1):

    threads      = physical cores this process may run on: distinct (package, core) pairs of os.sched_getaffinity(0),
                   capped by the cgroup CPU quota (cpu.max)
    OMP_PROC_BIND=close, OMP_PLACES=cores   one thread pinned per core, no migration
    first touch  = both arrays are written once over their WHOLE extent with the same static collapse(2) schedule the
                   compute loops use, so every page lives on the NUMA node of the thread that streams it

The reference has no runnable CPU path: this restatement is the CPU baseline (`kind: "port"`).
Prints one JSON object.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cgroup_cpu_quota():
    """CPUs allowed by the cgroup quota (None = unlimited / unknown)."""
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:  # cgroup v2
            quota, period = f.read().split()[:2]
            if quota != "max":
                return max(1, int(int(quota) / int(period)))
            return None
    except (OSError, ValueError):
        pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0 and p > 0:
            return max(1, q // p)
    except (OSError, ValueError):
        pass
    return None


def host_topology() -> dict:
    """Affinity, physical cores inside it, NUMA nodes they span, the thread count to use."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cpus = list(range(os.cpu_count() or 1))
    cores, nodes = set(), set()
    for c in cpus:
        base = f"/sys/devices/system/cpu/cpu{c}"
        try:
            pkg = open(f"{base}/topology/physical_package_id").read().strip()
            core = open(f"{base}/topology/core_id").read().strip()
            cores.add((pkg, core))
        except OSError:
            cores.add(("?", str(c)))
        try:
            for e in os.listdir(base):
                if e.startswith("node") and e[4:].isdigit():
                    nodes.add(int(e[4:]))
        except OSError:
            pass
    quota = cgroup_cpu_quota()
    threads = len(cores)
    if quota is not None:
        threads = max(1, min(threads, quota))
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"affinity_cpus": len(cpus), "physical_cores": len(cores), "cgroup_cpu_quota": quota, "numa_nodes": sorted(nodes), "threads": threads, "cpu": model}


def omp_env(threads: int) -> dict:
    return {"OMP_NUM_THREADS": str(threads), "OMP_PROC_BIND": "close", "OMP_PLACES": "cores", "OMP_DYNAMIC": "false"}


class CpuJacobi:
    """The reference loop on one periodic n^3 subdomain: interior -> 6-face periodic exchange -> exterior slabs -> swap."""

    def __init__(self, n: int, dtype):
        import numpy as np

        from oracle import c_oracle as co
        from oracle import geometry as g

        self.co, self.g, self.n = co, g, n
        self.r = g.Radius.face_edge_corner(1, 0, 0)
        raw = g.raw_size((n, n, n), self.r)
        self.cur = np.empty(raw[::-1], dtype=dtype)
        self.nxt = np.empty(raw[::-1], dtype=dtype)
        for a in (self.cur, self.nxt):  # first touch, whole allocation, the compute loops' static schedule
            co.fill(a, (0, 0, 0), raw, 0.0)
        co.fill(self.cur, (1, 1, 1), (n, n, n), 0.5)
        self.interior = g.get_interior((0, 0, 0), (n, n, n), self.r)
        self.exterior = g.get_exterior((0, 0, 0), (n, n, n), self.r)
        self.plan = g.plan_sends((1, 1, 1), {(0, 0, 0): (n, n, n)}, self.r)
        self.copies = [co.make_copies([(a, m["dst_pos"], a, m["src_pos"], m["ext"]) for m in self.plan]) for a in (self.cur, self.nxt)]
        self.par = 0

    def step(self):
        co, n = self.co, self.n
        cur, nxt = (self.cur, self.nxt) if self.par == 0 else (self.nxt, self.cur)
        creg = ((0, 0, 0), (n, n, n))
        co.jacobi_region(nxt, cur, (-1, -1, -1), *self.interior, *creg)
        co.translate_many(*self.copies[self.par])
        for lo, hi in self.exterior:
            co.jacobi_region(nxt, cur, (-1, -1, -1), lo, hi, *creg)
        self.par ^= 1


def run(n: int, dtype_name: str, steps: int, warmup: int, budget_s: float) -> dict:
    import numpy as np

    sys.path.insert(0, ROOT)
    from oracle import c_oracle as co

    topo = host_topology()
    want = int(os.environ.get("OMP_NUM_THREADS", topo["threads"]))
    co.set_num_threads(want)
    dtype = np.float64 if dtype_name == "f64" else np.float32
    cj = CpuJacobi(n, dtype)
    for _ in range(max(warmup, 1)):
        cj.step()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        cj.step()
        times.append(time.perf_counter() - t0)
        if budget_s > 0 and time.perf_counter() - t_begin > budget_s:
            break
    mean = float(np.mean(times))
    return {
        "value": n**3 / mean,
        "ms_per_step": mean * 1e3,
        "min_ms_per_step": float(np.min(times)) * 1e3,
        "steps": len(times),
        "warmup": max(warmup, 1),
        "cores": co.num_threads(),
        "host": topo,
        "omp": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OMP_PROC_BIND", "OMP_PLACES")},
    }


def run_in_subprocess(n: int, dtype_name: str, steps: int, warmup: int, budget_s: float = 0.0) -> dict:
    """What bench.py calls: this file as a fresh process with the pinned OpenMP environment."""
    import subprocess

    topo = host_topology()
    env = dict(os.environ)
    env.update(omp_env(topo["threads"]))
    cmd = [sys.executable, os.path.abspath(__file__), "--n", str(n), "--dtype", dtype_name, "--steps", str(steps), "--warmup", str(warmup), "--budget", str(budget_s)]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1200)
    if out.returncode != 0:
        raise RuntimeError("oracle/cpu_bench.py failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=512)
    p.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=1)
    p.add_argument("--budget", type=float, default=0.0, help="stop after this many seconds of timed steps (0 = run all)")
    a = p.parse_args()
    print(json.dumps(run(a.n, a.dtype, a.steps, a.warmup, a.budget)), flush=True)


if __name__ == "__main__":
    main()
