"""CPU tests that pin the oracle as far as it can be pinned.

The reference ships no tests, goldens or expected outputs, so the
oracle is "parity unpinned".  What holds it in place:
  * two independently written restatements (C++ loops vs numpy whole-array)
    agree bit for bit;
  * the closed-form pull rule equals the literal swap-then-shift sequence;
  * the analytic invariants T1 (a)-(f) of SURVEY 4;
  * the reference's own input fixtures (stock lb.in, geom.in, tuto, geom.pbm)
    copied as small files under tests/golden/ parse to the documented counts.
"""
import os

import numpy as np
import pytest

from oracle import numpy_restatement as R
from oracle import oracle as O
from tests.util import GOLDEN, random_nature, read_geom_in_py

CASES = [(6, 5, 7, 0.3, 1), (1, 1, 22, 0.1, 2), (4, 4, 4, 0.5, 3), (2, 3, 2, 0.3, 4), (1, 12, 9, 0.2, 5), (7, 1, 3, 0.25, 6)]


def test_lbm_table():
    c, a0, a1, a2, inv = O.lbm_table()
    assert np.array_equal(c, R.C) and np.array_equal(inv, R.INV)
    assert np.array_equal(a0, R.A0) and np.array_equal(a1, R.A1) and np.array_equal(a2, R.A2)
    # module_lbmodel.f90:122-136 worked by hand
    assert a1[0] == 1.0 and a2[0] == 1.5 and a2[1] == 0.25 and a2[7] == 0.125
    assert [int(inv[l]) + 1 for l in (1, 3, 5, 7, 8, 11, 12, 15, 16)] == [3, 5, 7, 11, 10, 15, 14, 19, 18]
    assert abs(a0.sum() - 1.0) < 1e-15


@pytest.mark.parametrize("lx,ly,lz,p,seed", CASES)
def test_two_restatements_agree_lb(lx, ly, lz, p, seed):
    tau = 0.8
    nat = random_nature(lx, ly, lz, p, seed)
    st = O.LBState(nat, 1.0, tau)
    a = dict(n=st.n.copy(), rho=st.rho.copy(), j=[np.zeros_like(st.rho) for _ in range(3)])
    b = dict(n=st.n.copy(), rho=st.rho.copy(), j=[np.zeros_like(st.rho) for _ in range(3)])
    f = [np.zeros_like(st.rho) for _ in range(3)]
    mass0 = st.n.sum()
    for t in range(1, 11):
        if t == 4:
            fv = [1e-3, -2e-3, 5e-4]
            st.set_force_uniform(fv)
            for d in range(3):
                f[d][nat == 0] = fv[d]
        rc, err = st.step()
        for s, pull in ((a, False), (b, True)):
            n, rho, jx, jy, jz, e, neg = R.lb_step(s["n"], s["rho"], *s["j"], *f, nat, tau, use_pull=pull)
            s.update(n=n, rho=rho, j=[jx, jy, jz])
            assert np.array_equal(n, st.n)
            assert np.array_equal(rho, st.rho)
            assert np.array_equal(jx, st.jx) and np.array_equal(jy, st.jy) and np.array_equal(jz, st.jz)
            assert e == err and int(neg) == rc
        assert (st.n[:, nat == 1] == 0).all()          # T1(b): solid populations stay exactly 0
        assert abs(st.n.sum() - mass0) < 1e-12 * mass0  # T1(a): mass conserved


@pytest.mark.parametrize("lx,ly,lz,p,seed", CASES[:4])
def test_two_restatements_agree_mp(lx, ly, lz, p, seed):
    nat = random_nature(lx, ly, lz, p, seed)
    itf = O.detect_interfacial(nat)
    st = O.LBState(nat)
    f = [1e-4, 2e-4, -1e-4]
    for _ in range(3):
        st.step()
    st.set_force_uniform(f)
    for _ in range(15):
        st.step()
    Db, ka, kd = 0.01, 0.1, 0.01
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, f, Db, ka, kd)
    ntr = R.tracer_population(nat, st.rho, st.jx, st.jy, st.jz, f)
    assert np.array_equal(np.moveaxis(ntr, 0, -1), mp.ntr)
    P, ads = R.mp_init(nat, itf, ntr, st.rho, Db, ka, kd)
    assert ads == bool(mp.ads)
    assert np.array_equal(P, mp.P[0])
    Pads = np.zeros_like(P)
    tot0 = (mp.P[0] + mp.Pads[0]).sum(axis=(0, 1, 2))
    for it in range(1, 25):
        rc, v, conv = mp.propagate()
        P, Pads, vacf, err = R.mp_propagate(nat, itf, ntr, st.rho, Db, ka, kd, ads, P, Pads)
        assert rc == int(err) == 0
        assert np.array_equal(P, mp.P[0]) and np.array_equal(Pads, mp.Pads[0])
        assert np.allclose(vacf, v, rtol=1e-12, atol=1e-20)
        assert (mp.P[1] == 0).all() and (mp.Pads[1] == 0).all()
        # T1(d): sum_r (P + Pads) per component is constant in time
        tot = (mp.P[0] + mp.Pads[0]).sum(axis=(0, 1, 2))
        assert np.allclose(tot, tot0, rtol=0, atol=1e-15 * max(1.0, np.abs(mp.P[0]).sum()))


def test_stock_input_exits_at_step_4():
    """T1(f): lb.in as shipped (1x1x102 slit, f_ext = 0) leaves the time loop at t=4."""
    nat = O.geometry(1, 1, 1, 102)
    assert nat.sum() == 2 and nat[0, 0, 0] == 1 and nat[-1, 0, 0] == 1
    r = O.equilibration(nat, [0.0, 0.0, 0.0])
    assert r["rc"] == 0 and r["t_exit"] == 4 and r["t_fext"] == 4
    assert (r["l2err"] == 0).all()


def test_bulk_vacf_known_answer():
    """T1(c): no solid, fluid at rest: vacf(0) = 2*Db per component and vacf(t>=1) = 0."""
    nat = O.geometry(-1, 4, 5, 3)
    itf = O.detect_interfacial(nat)
    assert not itf.any()
    st = O.LBState(nat)
    Db = 0.0123
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, [0, 0, 0], Db, 0.0, 0.0)
    assert np.allclose(mp.vacf0, 2 * Db, rtol=1e-13, atol=0)
    for _ in range(4):
        rc, v, conv = mp.propagate()
        assert rc == 0 and np.abs(v).max() < 1e-17


def test_slit_poiseuille_profile():
    """T1(e): j_x(z) ~ f/(2 nu) (z-1.5)(lz-0.5-z), nu=(tau-1/2)/3, to ~1e-3 (halfway bounce-back)."""
    lz, fx, tau = 22, 1e-6, 1.0
    nat = O.geometry(1, 1, 1, lz)
    r = O.equilibration(nat, [fx, 0, 0], tau=tau, target_error=1e-16, max_steps=20000)
    z = np.arange(1, lz + 1, dtype=float)
    nu = (tau - 0.5) / 3.0
    ana = fx / (2 * nu) * (z - 1.5) * (lz - 0.5 - z)
    jx = r["jx"][:, 0, 0]
    assert jx[0] == 0 and jx[-1] == 0
    assert np.abs(jx[1:-1] - ana[1:-1]).max() / ana.max() < 2e-3


def test_fixtures_parse_to_documented_counts():
    """The reference's own input files (copied verbatim as data fixtures)."""
    g = os.path.join(GOLDEN, "geom.in_chromat_1disks-dia10-1x50x50_v1")
    nat = O.read_geom_in(g, 1, 50, 50)
    assert nat.sum() == 79 and np.array_equal(nat, read_geom_in_py(g, 1, 50, 50))
    zs = np.where(nat)[0] + 1
    assert zs.min() == 8 and zs.max() == 17
    nat = O.read_geom_in(os.path.join(GOLDEN, "geom.in"), 9, 8, 9)
    assert nat.sum() == 190 or nat.sum() == 191
    nat = O.read_pbm(os.path.join(GOLDEN, "geom.pbm"), 1, 91, 25)
    assert nat.shape == (25, 91, 1) and 0 < nat.sum() < nat.size


def test_geometry_builders_exact_thresholds():
    """Cylinder / BCC thresholds against tie-free integer arithmetic."""
    for lx in (5, 8, 11, 26, 51):
        nat = O.geometry(2, lx, lx, 3)
        i = np.arange(1, lx + 1)
        # 4*((i-o)^2+(j-o)^2) >= (lx-1)^2 with o=(lx+1)/2 in half-units
        d2 = (2 * i[None, :] - (lx + 1)) ** 2 + (2 * i[:, None] - (lx + 1)) ** 2
        assert np.array_equal(nat[0] == 1, d2 >= (lx - 1) ** 2), lx
    for lx in (6, 9, 16):
        nat = O.geometry(3, lx, lx, lx)
        i = np.arange(1, lx + 1)
        K, J, I = np.meshgrid(i, i, i, indexing="ij")
        solid = np.zeros(nat.shape, bool)
        pts = [(a, b, c) for a in (2, 2 * lx) for b in (2, 2 * lx) for c in (2, 2 * lx)] + [(lx + 1,) * 3]
        for (a, b, c) in pts:  # doubled coordinates; 16 d^2 <= 3 (lx-1)^2  <=>  4*(2d)^2 <= 3 (lx-1)^2
            dd = (2 * I - a) ** 2 + (2 * J - b) ** 2 + (2 * K - c) ** 2
            solid |= 4 * dd <= 3 * (lx - 1) ** 2
        assert np.array_equal(nat == 1, solid), lx


def test_interfacial_flags():
    nat = random_nature(5, 4, 6, 0.3, 11)
    itf = O.detect_interfacial(nat)
    ref = np.zeros_like(nat)
    for l in range(1, 19):
        ref |= (nat != R.at_plus(nat, R.C[l])).astype(np.int8)
    assert np.array_equal(itf, ref)


def test_profiles_and_total_flux():
    rng = np.random.default_rng(3)
    shp = (5, 4, 3)
    rho, jx, jy, jz = (rng.random(shp) for _ in range(4))
    rho[rho < 0.2] = 0
    for axis, ax in ((0, (0, 1)), (1, (0, 2)), (2, (1, 2))):
        p = O.profiles(rho, jx, jy, jz, axis)
        assert np.allclose(p[:, 0], jx.sum(axis=ax), rtol=1e-14)
        assert np.allclose(p[:, 2], jz.sum(axis=ax), rtol=1e-14)
        cnt = np.maximum((rho > np.finfo(float).eps).sum(axis=ax), 1)
        assert np.allclose(p[:, 3], rho.sum(axis=ax) / cnt, rtol=1e-14)
    assert np.allclose(O.total_flux(jx, jy, jz), [jx.sum(), jy.sum(), jz.sum()], rtol=1e-14)
