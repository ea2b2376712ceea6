"""Exact-rational known-answer tests: the third, independent pin of the oracle and of the CUDA path.

oracle/exact_rational.py evaluates the reference's formulas (module_collision.f90:77-108,
equilibration.f90:204-300, drop_tracers.f90:97-105, module_moment_propagation.f90:96-137,207-267,
341-346) with fractions.Fraction on tiny lattices, sharing no code with the fp64 restatements.  The
fp64 results (C++ oracle on CPU; CUDA kernels through the C ABI on the GPU) must sit within a few
ulps of the largest term of each expression of the exact value.  A mis-association costs a few ulps
and passes (bit-exactness against the Fortran binary itself cannot be tested without a Fortran
compiler); a wrong sign / coefficient / direction / inverse / neighbour is off by >= 1e-6 relative and fails.
"""
from fractions import Fraction as Fr

import numpy as np
import pytest

from oracle import exact_rational as X
from oracle import oracle as O

EPS = np.finfo(np.float64).eps
TAUS = [1.0, 0.8]


def lattice():
    """4 x 3 x 3 with two solid nodes: fluid / solid links in x, y, z and diagonal directions, periodic wraps."""
    nat = np.zeros((3, 3, 4), np.int8)
    nat[1, 1, 2] = 1
    nat[0, 2, 0] = 1
    return nat


def lb_state(nat, seed):
    """A far-from-equilibrium state: random populations, density / momentum that do NOT derive from them
    (collide takes them as separate inputs), velocities ~0.05, a force field ~1e-2 that differs per node."""
    rng = np.random.default_rng(seed)
    fl = nat == 0
    n = np.zeros((19,) + nat.shape)
    w = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12)
    n[:, fl] = w[:, None] * rng.uniform(0.7, 1.3, size=(19, int(fl.sum())))   # positive after any collision below
    rho = np.zeros(nat.shape)
    rho[fl] = rng.uniform(0.8, 1.3, size=int(fl.sum()))
    j = [np.zeros(nat.shape) for _ in range(3)]
    F = [np.zeros(nat.shape) for _ in range(3)]
    for d in range(3):
        j[d][fl] = rng.uniform(-0.06, 0.06, size=int(fl.sum()))
        F[d][fl] = rng.uniform(-0.01, 0.01, size=int(fl.sum()))
    return n, rho, j, F


def exact_lb(nat, n, rho, j, F, tau):
    shape = nat.shape
    en = X.to_exact(n, shape, comps=19)
    er = X.to_exact(rho, shape)
    ej = {r: [Fr(float(j[d][r[2], r[1], r[0]])) for d in range(3)] for r in X.nodes(shape)}
    eF = {r: [Fr(float(F[d][r[2], r[1], r[0]])) for d in range(3)] for r in X.nodes(shape)}
    return X.lb_step(nat, en, er, ej, eF, tau)


def check_lb(nat, got_n, got_rho, got_j, ex, what):
    en, er, ej = ex
    # every term of the collision / moment expressions is below ~rho; a handful of operations each
    tol = 16 * EPS * 1.5
    worst = 0.0
    for r in X.nodes(nat.shape):
        i, jj, k = r
        if nat[k, jj, i] != 0:
            assert (got_n[:, k, jj, i] == 0).all(), f"{what}: solid node {r} holds a population"
            continue
        for l in range(19):
            worst = max(worst, abs(float(Fr(float(got_n[l, k, jj, i])) - en[r][l])))
        worst = max(worst, abs(float(Fr(float(got_rho[k, jj, i])) - er[r])))
        for d in range(3):
            worst = max(worst, abs(float(Fr(float(got_j[d][k, jj, i])) - ej[r][d])))
    assert worst <= tol, f"{what}: max |fp64 - exact| = {worst:.3e} > {tol:.3e}"
    return worst


def mp_inputs(nat, seed):
    rng = np.random.default_rng(seed)
    fl = nat == 0
    rho = np.zeros(nat.shape)
    rho[fl] = rng.uniform(0.9, 1.1, size=int(fl.sum()))
    j = [np.zeros(nat.shape) for _ in range(3)]
    for d in range(3):
        j[d][fl] = rng.uniform(-0.03, 0.03, size=int(fl.sum()))
    f_ext = [3e-3, -2e-3, 1e-3]
    return rho, j, f_ext


MP_PARAMS = [(0.01, 0.1, 0.01), (0.02, 0.0, 0.0), (0.01, 0.05, 0.2)]
MP_STEPS = 4


def exact_mp(nat, rho, j, f_ext, Db, ka, kd):
    shape = nat.shape
    itf = X.interfacial(nat)
    er = X.to_exact(rho, shape)
    ej = {r: [Fr(float(j[d][r[2], r[1], r[0]])) for d in range(3)] for r in X.nodes(shape)}
    ntr = X.tracer_populations(nat, er, ej, f_ext)
    st = X.mp_init(nat, itf, ntr, er, Db, ka, kd)
    out = dict(itf=itf, vacf0=st["vacf0"], P0={r: list(v) for r, v in st["P"].items()}, vacf=[], ads=st["ads"])
    for _ in range(MP_STEPS):
        v, mf = X.mp_propagate(nat, itf, ntr, er, st)
        assert mf > 0
        out["vacf"].append(v)
    out["P"], out["Pads"] = st["P"], st["Pads"]
    return out


def check_mp(nat, ex, vacf0, vacf, P, Pads, what):
    nf = int((nat == 0).sum())
    scale = 1.0 / nf                       # boltz_weight: every P component is a sum of <= 18 terms ~ q * bw
    tol = 256 * EPS * scale
    for d in range(3):
        assert abs(float(Fr(float(vacf0[d])) - ex["vacf0"][d])) <= 256 * EPS, f"{what}: vacf(0)[{d}]"
        for s in range(MP_STEPS):
            # vacf(t) = sum over nodes of P*u_star: terms ~ scale * 0.1 each, nf of them
            assert abs(float(Fr(float(vacf[s][d])) - ex["vacf"][s][d])) <= 256 * EPS * 0.1, f"{what}: vacf({s + 1})[{d}]"
    worst = 0.0
    for r in X.nodes(nat.shape):
        i, j, k = r
        for d in range(3):
            worst = max(worst, abs(float(Fr(float(P[k, j, i, d])) - ex["P"][r][d])))
            worst = max(worst, abs(float(Fr(float(Pads[k, j, i, d])) - ex["Pads"][r][d])))
    assert worst <= tol, f"{what}: max |fp64 - exact| over P, Pads = {worst:.3e} > {tol:.3e}"


# --------------------------------------------------------------------------- the model itself
def test_exact_model_tables():
    # module_lbmodel.f90:122-136 worked by hand: a1(rest) = 1, a2 = 3/2, 1/4, 1/8 up to the rounding of 1/3
    assert [X.INV[l] for l in (1, 3, 5, 7, 8, 11, 12, 15, 16)] == [2, 4, 6, 10, 9, 14, 13, 18, 17]
    assert abs(float(sum(X.A0)) - 1.0) < 1e-15
    assert float(X.A1[0]) == 1.0 and float(X.A2[0]) == 1.5 and float(X.A2[1]) == 0.25 and float(X.A2[7]) == 0.125


def test_exact_model_conserves_mass_and_momentum():
    """T1(a): without a force, one step conserves total mass and -- on an all-fluid lattice -- total momentum,
    when density and momentum are the populations' own moments."""
    nat = np.zeros((3, 3, 4), np.int8)
    n, _, _, _ = lb_state(nat, 3)
    rho = n.sum(0)
    c = np.array(X.C)
    j = [np.tensordot(c[:, d].astype(float), n, axes=(0, 0)) for d in range(3)]
    F = [np.zeros(nat.shape) for _ in range(3)]
    en, er, ej = exact_lb(nat, n, rho, j, F, 0.8)
    # rho, j are fp64 sums of n and the weights are fp64-rounded thirds: conservation holds to rounding
    assert abs(float(sum(er.values()) - sum(Fr(float(v)) for v in rho.ravel()))) < 1e-14
    for d in range(3):
        assert abs(float(sum(v[d] for v in ej.values()) - sum(Fr(float(v)) for v in j[d].ravel()))) < 1e-14


# --------------------------------------------------------------------------- oracle (CPU)
@pytest.mark.parametrize("tau", TAUS)
def test_oracle_lb_step_matches_exact(tau):
    nat = lattice()
    n, rho, j, F = lb_state(nat, 11)
    ex = exact_lb(nat, n, rho, j, F, tau)
    st = O.LBState(nat, 1.0, tau)
    st.n[...] = n
    st.rho[...] = rho
    st.jx[...], st.jy[...], st.jz[...] = j
    st.fx[...], st.fy[...], st.fz[...] = F
    rc, _ = st.step()
    assert rc == 0
    check_lb(nat, st.n, st.rho, [st.jx, st.jy, st.jz], ex, f"oracle tau={tau}")


def test_numpy_restatement_lb_step_matches_exact():
    from oracle import numpy_restatement as R
    nat = lattice()
    n, rho, j, F = lb_state(nat, 12)
    ex = exact_lb(nat, n, rho, j, F, 0.8)
    n2, rho2, jx, jy, jz, _, neg = R.lb_step(n.copy(), rho.copy(), *[a.copy() for a in j], *F, nat, 0.8, use_pull=True)
    assert not neg
    check_lb(nat, n2, rho2, [jx, jy, jz], ex, "numpy restatement")


@pytest.mark.parametrize("Db,ka,kd", MP_PARAMS)
def test_oracle_mp_matches_exact(Db, ka, kd):
    nat = lattice()
    rho, j, f_ext = mp_inputs(nat, 21)
    ex = exact_mp(nat, rho, j, f_ext, Db, ka, kd)
    itf = O.detect_interfacial(nat)
    for r in X.nodes(nat.shape):
        assert bool(itf[r[2], r[1], r[0]]) == ex["itf"][r]
    mp = O.MPState(nat, itf, rho, *j, f_ext, Db, ka, kd)
    assert bool(mp.ads) == ex["ads"]
    vacf = [mp.propagate()[1] for _ in range(MP_STEPS)]
    check_mp(nat, ex, mp.vacf0, vacf, mp.P[0], mp.Pads[0], f"oracle Db={Db} ka={ka} kd={kd}")


def test_kat_is_sensitive():
    """The tolerance separates rounding from formula errors: dropping the factor 2 of the 2*a2 force term
    (module_collision.f90:103), or using c_l instead of c_inv(l) in the propagated quantity, moves results by
    many orders of magnitude more than the tolerance."""
    nat = lattice()
    n, rho, j, F = lb_state(nat, 11)
    ex = exact_lb(nat, n, rho, j, F, 0.8)
    saved = list(X.A2)
    try:
        X.A2[:] = [a / 2 for a in saved]
        bad = exact_lb(nat, n, rho, j, F, 0.8)
    finally:
        X.A2[:] = saved
    diff = max(abs(float(bad[0][r][l] - ex[0][r][l])) for r in ex[0] for l in range(19))
    assert diff > 1e6 * 64 * EPS


# --------------------------------------------------------------------------- CUDA (through the C ABI)
@pytest.mark.gpu
@pytest.mark.parametrize("in_place", [False, True])
@pytest.mark.parametrize("tau", TAUS)
def test_cuda_lb_step_matches_exact(tau, in_place):
    import laboetie_b200 as lb
    nat = lattice()
    n, rho, j, F = lb_state(nat, 11)
    ex = exact_lb(nat, n, rho, j, F, tau)
    with lb.LaboetieGPU(nat) as sim:
        if in_place:
            sim.lb_set_in_place(True)
        sim.lb_upload(n, rho, *j)
        sim.lb_set_force_field(*F)
        done, conv, _ = sim.lb_step(1, tau=tau, check_every=1, target_error=-1.0)
        assert done == 1
        got_n = sim.lb_populations()
        got_rho, jx, jy, jz = sim.lb_moments()
        profs = [sim.lb_profiles(a) for a in range(3)]
        flux = sim.lb_total_flux()
        probe = sim.lb_probe(1, 2, 0)
    check_lb(nat, got_n, got_rho, [jx, jy, jz], ex, f"cuda tau={tau} in_place={in_place}")
    # A7: plane sums (equilibration.f90:161-172), total flux (:260) and the probe (:187) against exact sums of the
    # exact moments
    en, er, ej = ex
    lz, ly, lx = nat.shape
    tol = 64 * EPS * 1.5 * nat.size
    for axis, n_ax in ((0, lx), (1, ly), (2, lz)):
        for p in range(n_ax):
            sel = [r for r in X.nodes(nat.shape) if r[axis] == p and nat[r[2], r[1], r[0]] == 0]
            for d in range(3):
                assert abs(float(Fr(float(profs[axis][p, d])) - sum(ej[r][d] for r in sel))) <= tol
            mean = sum(er[r] for r in sel) / max(len(sel), 1)
            assert abs(float(Fr(float(profs[axis][p, 3])) - mean)) <= tol
    fl = [r for r in X.nodes(nat.shape) if nat[r[2], r[1], r[0]] == 0]
    for d in range(3):
        assert abs(float(Fr(float(flux[d])) - sum(ej[r][d] for r in fl))) <= tol
        assert abs(float(Fr(float(probe[d])) - ej[(1, 2, 0)][d])) <= tol
    assert abs(float(Fr(float(probe[3])) - er[(1, 2, 0)])) <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("nbt", ["0", "1"])
@pytest.mark.parametrize("Db,ka,kd", MP_PARAMS)
def test_cuda_mp_matches_exact(Db, ka, kd, nbt, monkeypatch):
    import laboetie_b200 as lb
    monkeypatch.setenv("LBG_MP_NBT", nbt)
    nat = lattice()
    rho, j, f_ext = mp_inputs(nat, 21)
    ex = exact_mp(nat, rho, j, f_ext, Db, ka, kd)
    with lb.LaboetieGPU(nat) as sim:
        itf = sim.interfacial()
        for r in X.nodes(nat.shape):
            if nat[r[2], r[1], r[0]] == 0:
                assert bool(itf[r[2], r[1], r[0]]) == ex["itf"][r]
        v0 = sim.mp_init_from_moments(rho, *j, Db, ka, kd, f_ext)
        done, conv, vacf = sim.mp_step(MP_STEPS)
        assert done == MP_STEPS
        P, A = sim.mp_download()
    check_mp(nat, ex, v0, vacf, P, A, f"cuda Db={Db} ka={ka} kd={kd} nbt={nbt}")


# --------------------------------------------------------------------------- the reference's own fixture geometry
def _tuto_state():
    """BASELINE config 1 (tuto geom.in_chromat_1disks-dia10-1x50x50_v1, a reference input fixture): the flow state
    after 6 oracle steps (force switched on after step 3), taken exactly as the start of one exact step."""
    import os
    from tests.util import GOLDEN
    nat = O.read_geom_in(os.path.join(GOLDEN, "geom.in_chromat_1disks-dia10-1x50x50_v1"), 1, 50, 50)
    f = [0.0, 1e-5, 0.0]
    st = O.LBState(nat, 1.0, 1.0)
    for _ in range(3):
        st.step()
    st.set_force_uniform(f)
    for _ in range(3):
        st.step()
    start = dict(n=st.n.copy(), rho=st.rho.copy(), j=[st.jx.copy(), st.jy.copy(), st.jz.copy()],
                 F=[st.fx.copy(), st.fy.copy(), st.fz.copy()])
    return nat, f, st, start


_TUTO_CACHE = {}


def _tuto_exact():
    if not _TUTO_CACHE:
        nat, f, st, start = _tuto_state()
        ex_lb = exact_lb(nat, start["n"], start["rho"], start["j"], start["F"], 1.0)
        st.step()          # the oracle's own step 7: rho, j of it feed Phase B below (exact inputs = these fp64 values)
        rho, j = st.rho.copy(), [st.jx.copy(), st.jy.copy(), st.jz.copy()]
        global MP_STEPS
        saved, MP_STEPS = MP_STEPS, 2
        try:
            ex_mp = exact_mp(nat, rho, j, f, 0.01, 0.1, 0.01)
        finally:
            MP_STEPS = saved
        _TUTO_CACHE.update(nat=nat, f=f, start=start, ex_lb=ex_lb, st=st, rho=rho, j=j, ex_mp=ex_mp)
    return _TUTO_CACHE


def _check_mp2(nat, ex, vacf0, vacf, P, Pads, what):
    global MP_STEPS
    saved, MP_STEPS = MP_STEPS, 2
    try:
        check_mp(nat, ex, vacf0, vacf, P, Pads, what)
    finally:
        MP_STEPS = saved


def test_oracle_on_the_tuto_fixture_matches_exact():
    c = _tuto_exact()
    nat, st = c["nat"], c["st"]
    check_lb(nat, st.n, st.rho, [st.jx, st.jy, st.jz], c["ex_lb"], "oracle, tuto fixture")
    itf = O.detect_interfacial(nat)
    mp = O.MPState(nat, itf, c["rho"], *c["j"], c["f"], 0.01, 0.1, 0.01)
    vacf = [mp.propagate()[1] for _ in range(2)]
    _check_mp2(nat, c["ex_mp"], mp.vacf0, vacf, mp.P[0], mp.Pads[0], "oracle, tuto fixture")


@pytest.mark.gpu
def test_cuda_on_the_tuto_fixture_matches_exact():
    import laboetie_b200 as lb
    c = _tuto_exact()
    nat, s0 = c["nat"], c["start"]
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_upload(s0["n"], s0["rho"], *s0["j"])
        sim.lb_set_force_field(*s0["F"])
        done, _, _ = sim.lb_step(1, tau=1.0, check_every=1, target_error=-1.0)
        assert done == 1
        got_n = sim.lb_populations()
        got_rho, jx, jy, jz = sim.lb_moments()
        check_lb(nat, got_n, got_rho, [jx, jy, jz], c["ex_lb"], "cuda, tuto fixture")
        v0 = sim.mp_init_from_moments(c["rho"], *c["j"], 0.01, 0.1, 0.01, c["f"])
        done, _, vacf = sim.mp_step(2)
        P, A = sim.mp_download()
    _check_mp2(nat, c["ex_mp"], v0, vacf, P, A, "cuda, tuto fixture")
