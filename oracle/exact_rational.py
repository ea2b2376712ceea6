"""Exact-rational known-answer model of the hot path on tiny lattices.

TEST INFRASTRUCTURE ONLY (imported by tests/ alone; never by laboetie_b200).

Purpose: pin the two floating-point restatements (oracle/laboetie_oracle.cpp and
oracle/numpy_restatement.py) and the CUDA kernels a third, independent way.  The reference cannot be
built here (no Fortran compiler) and ships no golden vectors, so nothing checks that the fp64
restatements evaluate the right FORMULAS -- two restatements by the same author would share a
misreading.  This module evaluates the reference's expressions in exact rational arithmetic
(fractions.Fraction) straight from the Fortran text, with none of the restatements' code:
direction table, weights, inverse directions, periodic wrap, swap-then-shift bounce-back and
streaming, moments, tracer populations, scattering probabilities, moment propagation with
adsorption / desorption.  Inputs are fp64 values taken exactly; the model-constant PARAMETERs
(1/3, 1/18, 1/36, a1 = w/csq, a2 = w/(2 csq**2), kBT) are the fp64-rounded values the compiler
folds (module_lbmodel.f90:122-136), taken exactly.  The exact result differs from any fp64
evaluation of the same formula by a few ulps of the largest term; a wrong sign, coefficient,
direction or inverse differs by many orders of magnitude more.

Array conventions as in oracle/oracle.py: nature (lz, ly, lx) int8 with 0 = fluid, 1 = solid;
n (19, lz, ly, lx); P / Pads (lz, ly, lx, 3).  Everything is kept in dicts keyed by (i, j, k).
"""
from fractions import Fraction as Fr

# module_lbmodel.f90:65-85 (D3Q19), 0-based here, lmin..lmax = 0..18
C = [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1),
     (1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0), (1, 0, 1), (-1, 0, 1), (1, 0, -1), (-1, 0, -1),
     (0, 1, 1), (0, -1, 1), (0, 1, -1), (0, -1, -1)]
NV = 19
# determine_velocity_inverse, module_lbmodel.f90:160-166
INV = [next(li for li in range(NV) if all(C[l][d] == -C[li][d] for d in range(3))) for l in range(NV)]

# init_weight_factors, module_lbmodel.f90:122-136: fp64 PARAMETER arithmetic, then exact
_csq = 1.0 / 3.0
_w = [1.0 / 3.0, 1.0 / 18.0, 1.0 / 36.0]
_kind = [0] + [1] * 6 + [2] * 12           # module_lbmodel.f90:141-143
A0 = [Fr(_w[_kind[l]]) for l in range(NV)]
A1 = [Fr(_w[_kind[l]] / _csq) for l in range(NV)]
A2 = [Fr(_w[_kind[l]] / (2 * _csq ** 2)) for l in range(NV)]
CSQ = Fr(_csq)
KBT = Fr(1.0 / 3.0)                       # system kBT = 1/3 (module_system / init_simu)


def nodes(shape):
    lz, ly, lx = shape
    return [(i, j, k) for k in range(lz) for j in range(ly) for i in range(lx)]


def nb(r, l, shape):
    """pbc(i + c_l) in every direction (module_geometry pbc)."""
    lz, ly, lx = shape
    return ((r[0] + C[l][0]) % lx, (r[1] + C[l][1]) % ly, (r[2] + C[l][2]) % lz)


def to_exact(arr, shape, comps=None):
    """numpy (…, lz, ly, lx) -> dict[(i,j,k)] of Fraction (or list over the leading axis)."""
    out = {}
    for (i, j, k) in nodes(shape):
        if comps is None:
            out[(i, j, k)] = Fr(float(arr[k, j, i]))
        else:
            out[(i, j, k)] = [Fr(float(arr[c, k, j, i])) for c in range(comps)]
    return out


def interfacial(nature):
    """supercell_definition.f90:115-147: a node with a neighbour of the other nature."""
    shape = nature.shape
    out = {}
    for r in nodes(shape):
        i, j, k = r
        out[r] = any(nature[k, j, i] != nature[q[2], q[1], q[0]] for q in (nb(r, l, shape) for l in range(1, NV)))
    return out


def lb_step(nature, n, rho, j, F, tau):
    """One body of the equilibration time loop, exact (equilibration.f90:190-300).

    n: dict r -> [19]; rho: dict r -> Fr; j, F: dict r -> [3]; tau: float.  Returns (n, rho, j) of the new step.
    """
    shape = nature.shape
    tau = Fr(float(tau))
    fluid = lambda r: nature[r[2], r[1], r[0]] == 0   # noqa: E731
    n = {r: list(v) for r, v in n.items()}
    # ---- collide, module_collision.f90:77-108 (second-order branch), fluid nodes only
    for r in nodes(shape):
        if not fluid(r):
            continue
        jx, jy, jz = j[r]
        fx, fy, fz = F[r]
        d = rho[r]
        ux, uy, uz = jx / d, jy / d, jz / d
        for l in range(NV):
            cx, cy, cz = C[l]
            neq = (A0[l] * d + A1[l] * (cx * jx + cy * jy + cz * jz)
                   + A2[l] * (jx * ux * (cx ** 2 - CSQ) + jx * uy * cx * cy + jx * uz * cx * cz
                              + jy * ux * cy * cx + jy * uy * (cy ** 2 - CSQ) + jy * uz * cy * cz
                              + jz * ux * cz * cx + jz * uy * cz * cy + jz * uz * (cz ** 2 - CSQ)))
            n[r][l] = ((1 - 1 / tau) * n[r][l] + (1 / tau) * neq
                       + (1 - 1 / (2 * tau)) * (A1[l] * ((cx - ux) * fx + (cy - uy) * fy + (cz - uz) * fz)
                                                + 2 * A2[l] * (cx * ux + cy * uy + cz * uz) * (cx * fx + cy * fy + cz * fz)))
    # ---- bounce back, equilibration.f90:204-222: l = lmin, lmin+2, ... (1-based odd l = 0-based even l)
    for l in range(0, NV, 2):
        for r in nodes(shape):
            p = nb(r, l, shape)
            if nature[r[2], r[1], r[0]] != nature[p[2], p[1], p[0]]:
                n[r][l], n[p][INV[l]] = n[p][INV[l]], n[r][l]
    # ---- propagation, equilibration.f90:227-243
    new = {r: [None] * NV for r in n}
    for l in range(NV):
        for r in nodes(shape):
            new[nb(r, l, shape)][l] = n[r][l]
    n = new
    # ---- density = SUM(n,4) (:254); j = f/2 + sum n c (:287-294)
    rho2 = {r: sum(n[r]) for r in n}
    j2 = {r: [F[r][d] / 2 + sum(n[r][l] * C[l][d] for l in range(NV)) for d in range(3)] for r in n}
    return n, rho2, j2


def tracer_populations(nature, rho, j, f_ext):
    """drop_tracers.f90:97-105 for a neutral tracer (tr%q = 0)."""
    shape = nature.shape
    out = {}
    for r in nodes(shape):
        fl = nature[r[2], r[1], r[0]] == 0
        t = [j[r][d] + (Fr(float(f_ext[d])) if fl else 0) for d in range(3)]
        out[r] = [A0[l] * rho[r] + A1[l] * sum(C[l][d] * t[d] for d in range(3)) for l in range(NV)]
    return out


def scattprop(n, rho, w, lam, fermi):
    """calc_scattprop, module_moment_propagation.f90:341-346."""
    return n / rho - w + lam * w * fermi


def mp_init(nature, itf, ntr, rho, Db, ka, kd):
    """module_moment_propagation.f90:30-137 (neutral tracer).  Returns dict with P, Pads, vacf0, lam, ads, K."""
    shape = nature.shape
    eps = Fr(2.0 ** -52)
    Db, ka, kd = Fr(float(Db)), Fr(float(ka)), Fr(float(kd))
    K = Fr(0) if abs(kd) <= eps else ka / kd
    ads = abs(K) > eps
    lam = 4 * Db / KBT                                    # calc_lambda :333-338
    fluid = lambda r: nature[r[2], r[1], r[0]] == 0        # noqa: E731
    nf = sum(1 for r in nodes(shape) if fluid(r))
    nif = sum(1 for r in nodes(shape) if fluid(r) and itf[r])
    Pstat = nf + K * nif                                   # :100-101
    bw = 1 / Pstat
    half = Fr(1, 2)                                        # fermi = 1/(1+1)
    vacf0 = [Fr(0)] * 3
    P = {r: [Fr(0)] * 3 for r in nodes(shape)}
    for r in nodes(shape):
        if not fluid(r):
            continue
        for l in range(1, NV):
            p = nb(r, l, shape)
            if not fluid(p):
                continue
            sp = scattprop(ntr[r][l], rho[r], A0[l], lam, half)
            vacf0 = [vacf0[d] + bw * sp * C[l][d] ** 2 for d in range(3)]
            li = INV[l]
            spp = scattprop(ntr[p][li], rho[p], A0[li], lam, 1 - half)
            P[r] = [P[r][d] + spp * C[li][d] * bw for d in range(3)]
    return dict(P=P, Pads={r: [Fr(0)] * 3 for r in nodes(shape)}, vacf0=vacf0, lam=lam, ads=ads, ka=ka, kd=kd)


def mp_propagate(nature, itf, ntr, rho, st):
    """One PROPAGATE call, module_moment_propagation.f90:207-267.  Updates st['P'], st['Pads']; returns
    (vacf of this step, smallest remaining fraction)."""
    shape = nature.shape
    fluid = lambda r: nature[r[2], r[1], r[0]] == 0        # noqa: E731
    lam, ka, kd, ads = st["lam"], st["ka"], st["kd"], st["ads"]
    half = Fr(1, 2)
    Pn, An = st["P"], st["Pads"]
    Pnext = {r: [Fr(0)] * 3 for r in Pn}
    Anext = {r: [Fr(0)] * 3 for r in Pn}
    vacf = [Fr(0)] * 3
    min_frac = None
    for r in nodes(shape):
        if not fluid(r):
            continue
        u = [Fr(0)] * 3
        frac = Fr(1)
        acc = list(Pnext[r])
        for l in range(1, NV):
            p = nb(r, l, shape)
            if not fluid(p):
                continue
            sp = scattprop(ntr[r][l], rho[r], A0[l], lam, half)
            frac -= sp
            u = [u[d] + sp * C[l][d] for d in range(3)]
            li = INV[l]
            spp = scattprop(ntr[p][li], rho[p], A0[li], lam, 1 - half)
            acc = [acc[d] + Pn[p][d] * spp for d in range(3)]
        vacf = [vacf[d] + Pn[r][d] * u[d] for d in range(3)]
        if (not itf[r] and ads) or not ads:
            Pnext[r] = [acc[d] + frac * Pn[r][d] for d in range(3)]
        else:
            frac -= ka
            Pnext[r] = [acc[d] + frac * Pn[r][d] + An[r][d] * kd for d in range(3)]
            Anext[r] = [An[r][d] * (1 - kd) + Pn[r][d] * ka for d in range(3)]
        min_frac = frac if min_frac is None or frac < min_frac else min_frac
    st["P"], st["Pads"] = Pnext, Anext
    return vacf, min_frac
