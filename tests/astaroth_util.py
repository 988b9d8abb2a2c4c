"""Shared pieces of the astaroth parity tests (CPU and GPU sides)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "astaroth_solve_ref.npz")
DS = 0.04908738521


def make_fields(shape_xyz, seed: int, dtype=np.float64, smooth: float = 1.0, noise: float = 0.25):
    """16 arrays (in[8], out[8]) of raw size shape_xyz (x, y, z): seeded noise + a smooth component, all O(1)."""
    mx, my, mz = shape_xyz
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1.0, 1.0, size=(16, mz, my, mx))
    Z, Y, X = np.meshgrid(np.arange(mz) * DS, np.arange(my) * DS, np.arange(mx) * DS, indexing="ij")
    for f in range(16):
        a[f] = noise * a[f] + smooth * np.sin((1 + f % 3) * X + 0.3 * f) * np.cos((1 + f % 2) * Y) * np.sin(2 * Z + f)
    return [np.ascontiguousarray(a[f].astype(dtype)) for f in range(16)]


def tolerance(dtype) -> float:
    """Relative to the largest magnitude of the compared field.  FP64: the CUDA kernels contract a*b+c into FMAs and
    fold exp(lnT0 + arg), the oracle (-ffp-contract=off) does not -- differences are a few ulp of terms up to 1e4 times
    larger than the result.  FP32: same, at float epsilon."""
    return 2e-11 if np.dtype(dtype) == np.float64 else 2e-4
