"""Second, independently written restatement of the laboetie hot path (numpy).

TEST INFRASTRUCTURE ONLY.  Its one job is to cross-check the C++ oracle: the two
were written separately (this one whole-array with np.roll, the C++ one with
explicit index loops and the reference's il/jl/kl tables) and must agree bit for
bit on the per-node quantities.  It also carries the closed-form pull rule
(SURVEY 8a "A2 o A3") that the CUDA kernels implement, so that rule is checked
against the literal swap-then-shift sequence on the CPU as well.

PARITY UNPINNED (no reference goldens exist; see laboetie_oracle.cpp).

Reference lines restated: module_lbmodel.f90:66-86,122-162;
module_collision.f90:77-108; equilibration.f90:204-300,339-343;
drop_tracers.f90:97-105; module_moment_propagation.f90:96-137,207-267.
"""
import numpy as np

C = np.array([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
              [1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0], [1, 0, 1], [-1, 0, 1], [1, 0, -1],
              [-1, 0, -1], [0, 1, 1], [0, -1, 1], [0, 1, -1], [0, -1, -1]], dtype=np.int64)
INV = np.array([int(np.where((C == -C[l]).all(1))[0][0]) for l in range(19)])
_one, _three = np.float64(1.0), np.float64(3.0)
CSQ = _one / _three
W = np.array([_one / _three] + [_one / np.float64(18.0)] * 6 + [_one / np.float64(36.0)] * 12)
A0 = W.copy()
A1 = W / CSQ
A2 = W / (2 * (CSQ * CSQ))
EPS = np.finfo(np.float64).eps


def at_plus(a, c):
    """a evaluated at r + c (periodic); arrays are (lz, ly, lx)."""
    return np.roll(a, shift=(-c[2], -c[1], -c[0]), axis=(0, 1, 2))


def collide(n, rho, jx, jy, jz, fx, fy, fz, fluid, tau):
    ux = np.zeros_like(rho); uy = np.zeros_like(rho); uz = np.zeros_like(rho)
    ux[fluid] = jx[fluid] / rho[fluid]
    uy[fluid] = jy[fluid] / rho[fluid]
    uz[fluid] = jz[fluid] / rho[fluid]
    out = n.copy()
    with np.errstate(all="ignore"):
        for l in range(19):
            cx, cy, cz = (np.float64(v) for v in C[l])
            neq = A0[l] * rho + A1[l] * (cx * jx + cy * jy + cz * jz) + A2[l] * (
                jx * ux * (cx * cx - CSQ) + jx * uy * cx * cy + jx * uz * cx * cz
                + jy * ux * cy * cx + jy * uy * (cy * cy - CSQ) + jy * uz * cy * cz
                + jz * ux * cz * cx + jz * uy * cz * cy + jz * uz * (cz * cz - CSQ))
            new = (1.0 - 1.0 / tau) * n[l] + (1.0 / tau) * neq + (1.0 - 1.0 / (2.0 * tau)) * (
                A1[l] * ((cx - ux) * fx + (cy - uy) * fy + (cz - uz) * fz)
                + 2.0 * A2[l] * (cx * ux + cy * uy + cz * uz) * (cx * fx + cy * fy + cz * fz))
            out[l][fluid] = new[fluid]
    return out


def bounce_back(n, nature):
    n = n.copy()
    for l in range(0, 19, 2):
        nat_p = at_plus(nature, C[l])
        differs = nature != nat_p                      # at r: nature(r) != nature(r+c_l)
        a = n[l].copy()                                # n(r, l)
        b = at_plus(n[INV[l]], C[l])                   # n(r+c_l, inv l), seen from r
        n[l] = np.where(differs, b, a)
        # write n_loc back at r+c_l: roll the update to the neighbour's frame
        upd = np.roll(np.where(differs, a, b), shift=(C[l][2], C[l][1], C[l][0]), axis=(0, 1, 2))
        n[INV[l]] = upd
    return n


def stream(n):
    return np.stack([np.roll(n[l], shift=(C[l][2], C[l][1], C[l][0]), axis=(0, 1, 2)) for l in range(19)])


def pull_closed_form(nstar, nature):
    """n+(r,l) = n*(r-c_l,l) if r-c_l fluid else n*(r,inv l); solid nodes keep 0."""
    fluid = nature == 0
    out = np.zeros_like(nstar)
    for l in range(19):
        src_fluid = at_plus(fluid, -C[l])
        pulled = at_plus(nstar[l], -C[l])
        out[l] = np.where(fluid, np.where(src_fluid, pulled, nstar[INV[l]]), 0.0)
    return out


def moments(n, fx, fy, fz):
    rho = np.zeros_like(n[0])
    for l in range(19):
        rho = rho + n[l]
    jx, jy, jz = fx / 2.0, fy / 2.0, fz / 2.0
    for l in range(19):
        jx = jx + n[l] * np.float64(C[l][0])
        jy = jy + n[l] * np.float64(C[l][1])
        jz = jz + n[l] * np.float64(C[l][2])
    return rho, jx, jy, jz


def lb_step(n, rho, jx, jy, jz, fx, fy, fz, nature, tau, use_pull=False):
    fluid = nature == 0
    ns = collide(n, rho, jx, jy, jz, fx, fy, fz, fluid, tau)
    n2 = pull_closed_form(ns, nature) if use_pull else stream(bounce_back(ns, nature))
    rho2, jx2, jy2, jz2 = moments(n2, fx, fy, fz)
    l2err = max(np.abs(jx2 - jx).max(), np.abs(jy2 - jy).max(), np.abs(jz2 - jz).max())
    return n2, rho2, jx2, jy2, jz2, l2err, bool((n2 < 0).any())


def tracer_population(nature, rho, jx, jy, jz, f_ext):
    """(19, lz, ly, lx); drop_tracers.f90:97-105 with q=0, elec_slope=0."""
    fluid = nature == 0
    out = np.zeros((19,) + rho.shape)
    j = (jx, jy, jz)
    for l in range(19):
        s = np.zeros_like(rho)
        for d in range(3):
            t = np.where(fluid, j[d] + f_ext[d] - rho * 0.0 * 1.0 * 0.0, j[d])
            s = s + np.float64(C[l][d]) * t
        out[l] = A0[l] * rho + A1[l] * s
    return out


def _scatt(n, rho, w, lam, fermi):
    with np.errstate(all="ignore"):
        return n / rho - w + lam * w * fermi


def mp_init(nature, interfacial, ntr, rho, Db, ka, kd):
    fluid = nature == 0
    K = 0.0 if abs(kd) <= EPS else ka / kd
    lam = 4.0 * Db / (_one / _three)
    Pstat = float(fluid.sum()) + K * float((fluid & (interfacial != 0)).sum())
    bw = 1.0 / Pstat
    P = np.zeros(rho.shape + (3,))
    for l in range(1, 19):
        li = INV[l]
        ok = fluid & at_plus(fluid, C[l])
        spp = _scatt(at_plus(ntr[li], C[l]), at_plus(rho, C[l]), A0[li], lam, 1.0 - 0.5)
        for d in range(3):
            P[..., d] = np.where(ok, P[..., d] + 1.0 * spp * np.float64(C[li][d]) * bw, P[..., d])
    return P, abs(K) > EPS


def mp_propagate(nature, interfacial, ntr, rho, Db, ka, kd, ads, P, Pads):
    fluid = nature == 0
    lam = 4.0 * Db / (_one / _three)
    frac = np.ones_like(rho)
    ustar = np.zeros(rho.shape + (3,))
    acc = np.zeros(rho.shape + (3,))
    for l in range(1, 19):
        li = INV[l]
        ok = fluid & at_plus(fluid, C[l])
        sp = _scatt(ntr[l], rho, A0[l], lam, 0.5)
        frac = np.where(ok, frac - sp, frac)
        spp = _scatt(at_plus(ntr[li], C[l]), at_plus(rho, C[l]), A0[li], lam, 1.0 - 0.5)
        for d in range(3):
            ustar[..., d] = np.where(ok, ustar[..., d] + sp * np.float64(C[l][d]), ustar[..., d])
            acc[..., d] = np.where(ok, acc[..., d] + at_plus(P[..., d], C[l]) * spp, acc[..., d])
    vacf = np.array([(P[..., d] * ustar[..., d])[fluid].sum() for d in range(3)])
    sel = fluid & (interfacial != 0) if ads else np.zeros_like(fluid)
    frac2 = np.where(sel, frac - ka, frac)
    Pn = np.zeros_like(P)
    An = np.zeros_like(P)
    for d in range(3):
        plain = acc[..., d] + frac * P[..., d]
        adsb = acc[..., d] + frac2 * P[..., d] + Pads[..., d] * kd
        Pn[..., d] = np.where(fluid, np.where(sel, adsb, plain), 0.0)
        An[..., d] = np.where(sel, Pads[..., d] * (1.0 - kd) + P[..., d] * ka, 0.0)
    err = bool((frac2[fluid] < EPS).any())
    return Pn, An, vacf, err
