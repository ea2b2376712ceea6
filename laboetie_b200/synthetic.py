"""Geometry builders for the benchmark configurations (host side, numpy, slab-aware).

The reference builds `nature` in its Fortran driver (module_geometry.f90) and that
driver stays; these helpers exist so that bench.py and the tests can build the
BASELINE configurations -- including single z-slabs of lattices far too large to
hold on one host -- without the oracle.  Labels 1/2/3 follow
module_geometry.f90:158-166 (slit), :253-277 (cylinder: solid iff
|r-(l+1)/2| >= (lx-1)/2) and :206-245 (BCC: solid iff the distance to a cube
corner or to the centre is <= (lx-1)*sqrt(3)/4), evaluated tie-free in integer
arithmetic; tests/test_synthetic.py checks them against the literal oracle.

Every builder returns int8 (nz, ly, lx) for global planes k0 .. k0+nz-1, taken
periodically (so k0=-1, nz=nzl+2 yields a slab with its two halo planes).
"""
import numpy as np


def _planes(k0, nz, lz):
    return (np.arange(k0, k0 + nz) % lz).astype(np.int64)


def slit(lx, ly, lz, k0=0, nz=None):
    nz = lz if nz is None else nz
    k = _planes(k0, nz, lz)
    nat = np.zeros((nz, ly, lx), np.int8)
    nat[(k == 0) | (k == lz - 1)] = 1
    return nat


def cylinder(lx, ly, lz, k0=0, nz=None):
    if lx != ly or lx < 3:
        raise ValueError("cylinder needs lx == ly >= 3")
    nz = lz if nz is None else nz
    i = np.arange(1, lx + 1, dtype=np.int64)
    d2 = (2 * i[None, :] - (lx + 1)) ** 2 + (2 * i[:, None] - (ly + 1)) ** 2   # 4 |r - o|^2
    plane = (d2 >= (lx - 1) ** 2).astype(np.int8)
    return np.broadcast_to(plane, (nz, ly, lx)).copy()


def bcc(lx, ly, lz, k0=0, nz=None):
    if not (lx == ly == lz):
        raise ValueError("bcc needs a cubic cell")
    nz = lz if nz is None else nz
    k = _planes(k0, nz, lz) + 1
    i = np.arange(1, lx + 1, dtype=np.int64)
    thr = 3 * (lx - 1) ** 2                      # 16 d^2 <= 3 (lx-1)^2, with (2d)^2 in integers: 4 (2d)^2 <= thr
    solid = np.zeros((nz, ly, lx), bool)
    ax = lambda a: (2 * i - a) ** 2              # noqa: E731
    az = lambda a: (2 * k - a) ** 2              # noqa: E731
    for a in (2, 2 * lx):
        for b in (2, 2 * ly):
            for c in (2, 2 * lz):
                solid |= 4 * (ax(a)[None, None, :] + ax(b)[None, :, None] + az(c)[:, None, None]) <= thr
    solid |= 4 * (ax(lx + 1)[None, None, :] + ax(ly + 1)[None, :, None] + az(lz + 1)[:, None, None]) <= thr
    return solid.astype(np.int8)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def bernoulli(lx, ly, lz, p_solid=0.25, seed=12345, k0=0, nz=None):
    """i.i.d. hash noise: node (i,j,k) is solid iff splitmix64(seed, linear index) < p * 2^64."""
    nz = lz if nz is None else nz
    k = _planes(k0, nz, lz)
    with np.errstate(over="ignore"):
        idx = (k[:, None, None] * ly + np.arange(ly, dtype=np.int64)[None, :, None]) * lx + np.arange(lx, dtype=np.int64)[None, None, :]
        h = _splitmix64(idx.astype(np.uint64) + np.uint64(seed) * np.uint64(0x632BE59BD9B4E019))
    thr = np.uint64(min(int(p_solid * 2.0 ** 64), 2 ** 64 - 1))
    return (h < thr).astype(np.int8)


def porous_spheres(lx, ly, lz, porosity=0.6, radius=8, seed=12345, k0=0, nz=None):
    """Overlapping solid spheres at counter-based random centres (periodic), Boolean model:
    the number of spheres is fixed so that the expected porosity is `porosity`
    (exp(-n V_s / V) = porosity), which keeps every rank's view of the global geometry identical."""
    nz = lz if nz is None else nz
    vol = float(lx) * ly * lz
    # lattice volume of one sphere (number of integer points with d^2 <= r^2)
    a = np.arange(-radius, radius + 1)
    ball = (a[:, None, None] ** 2 + a[None, :, None] ** 2 + a[None, None, :] ** 2) <= radius * radius
    nsph = int(round(-np.log(porosity) * vol / ball.sum()))
    with np.errstate(over="ignore"):
        ctr = np.arange(3 * nsph, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x632BE59BD9B4E019)
        h = _splitmix64(ctr).reshape(nsph, 3)
    cx = (h[:, 0] % np.uint64(lx)).astype(np.int64)
    cy = (h[:, 1] % np.uint64(ly)).astype(np.int64)
    cz = (h[:, 2] % np.uint64(lz)).astype(np.int64)
    k = _planes(k0, nz, lz)
    nat = np.zeros((nz, ly, lx), np.int8)
    dz, dy, dx = np.nonzero(ball)
    dz, dy, dx = dz - radius, dy - radius, dx - radius
    # global plane -> local slot(s); a global plane can appear more than once when nz > lz
    slot_tables = []
    remaining = list(enumerate(k))
    while remaining:
        tab = np.full(lz, -1, np.int64)
        rest = []
        for li, kg in remaining:
            if tab[kg] < 0:
                tab[kg] = li
            else:
                rest.append((li, kg))
        slot_tables.append(tab)
        remaining = rest
    present = np.zeros(lz, bool)
    present[k] = True
    # keep only spheres that reach one of our planes
    reach = np.zeros(nsph, bool)
    for o in range(-radius, radius + 1):
        reach |= present[(cz + o) % lz]
    ids = np.nonzero(reach)[0]
    for c0 in range(0, len(ids), 2048):
        sel = ids[c0:c0 + 2048]
        gz = ((cz[sel, None] + dz[None, :]) % lz).ravel()
        gy = ((cy[sel, None] + dy[None, :]) % ly).ravel()
        gx = ((cx[sel, None] + dx[None, :]) % lx).ravel()
        for tab in slot_tables:
            sl = tab[gz]
            ok = sl >= 0
            nat[sl[ok], gy[ok], gx[ok]] = 1
    return nat


WORKLOADS = {
    # name: (builder, lx, ly, lz per GPU, f_ext, description)
    "cfg2": (slit, 64, 64, 256, (1e-6, 0.0, 0.0), "geometryLabel=1 slit 64x64x256"),
    "slitL": (slit, 512, 512, 256, (1e-6, 0.0, 0.0), "geometryLabel=1 slit 512x512x256 (all-fluid steady-state probe)"),
    "cfg3": (bcc, 256, 256, 256, (1e-6, 0.0, 0.0), "geometryLabel=3 BCC spheres 256^3"),
    "cfg5w": (porous_spheres, 1024, 1024, 128, (1e-6, 0.0, 0.0),
              "synthetic random porous 1024x1024x(128 per GPU), spheres r=8, porosity~0.6, splitmix64 seed 12345"),
    "cfg5b": (bernoulli, 1024, 1024, 128, (1e-6, 0.0, 0.0),
              "synthetic Bernoulli(p_solid=0.25) hash noise 1024x1024x(128 per GPU), seed 12345"),
    # strong-scaling configurations (fixed lattice, cut into z-slabs)
    "cfg4": (cylinder, 512, 512, 1024, (0.0, 0.0, 1e-6), "geometryLabel=2 cylinder along z 512x512x1024"),
    "cfg5s": (porous_spheres, 1024, 1024, 1024, (1e-6, 0.0, 0.0),
              "synthetic random porous 1024^3 (strong scaling), spheres r=8, porosity~0.6, splitmix64 seed 12345"),
}
# workloads whose z extent grows with the number of GPUs (weak scaling); all others are strong
WEAK = {"cfg5w", "cfg5b"}
