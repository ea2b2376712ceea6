"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerance: north_star asks for 1e-12 relative in fp64 (the allowed difference
being summation order of cross-node sums).  Per-node quantities (populations,
density, momentum, P, Pads, interfacial flags, l2err, exit step) are expected
to be *bit-identical* and are tested with array_equal; cross-node sums (vacf,
profiles, total flux) are tested at RTOL = 1e-12 relative to the field scale.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import GOLDEN, RTOL, random_nature, rel_err, sum_rtol

pytestmark = pytest.mark.gpu


def _gpu():
    import laboetie_b200 as lb
    return lb


GEOMS = [
    ("rand6x5x7", lambda: random_nature(6, 5, 7, 0.3, 1)),
    ("line1x1x22", lambda: random_nature(1, 1, 22, 0.1, 2)),
    ("rand4x4x4", lambda: random_nature(4, 4, 4, 0.5, 3)),
    ("rand2x3x2", lambda: random_nature(2, 3, 2, 0.3, 4)),
    ("rand33x7x5", lambda: random_nature(33, 7, 5, 0.25, 7)),
    ("rand70x3x4", lambda: random_nature(70, 3, 4, 0.2, 8)),
    ("plane1x12x9", lambda: random_nature(1, 12, 9, 0.2, 5)),
    ("row7x1x3", lambda: random_nature(7, 1, 3, 0.25, 6)),
    ("slit8x8x16", lambda: O.geometry(1, 8, 8, 16)),
    ("cyl11x11x4", lambda: O.geometry(2, 11, 11, 4)),
    ("bcc12", lambda: O.geometry(3, 12, 12, 12)),
    ("bulk5x4x3", lambda: O.geometry(-1, 5, 4, 3)),
    ("onefluid3x3x3", lambda: _one_fluid()),
    ("solidplane6x5x8", lambda: _solid_plane()),
    ("rand40x9x6", lambda: random_nature(40, 9, 6, 0.6, 9)),
]


def _one_fluid():
    nat = np.ones((3, 3, 3), np.int8)
    nat[1, 1, 1] = 0
    return nat


def _solid_plane():
    nat = random_nature(6, 5, 8, 0.2, 12)
    nat[3] = 1          # a whole z-plane of solid
    nat[:, 2, :] = 1    # and a whole y-row slab
    return nat


@pytest.mark.parametrize("name,mk", GEOMS, ids=[g[0] for g in GEOMS])
def test_interfacial_flags_and_counts(name, mk):
    lb = _gpu()
    nat = mk()
    with lb.LaboetieGPU(nat) as sim:
        itf = sim.interfacial()
        ref = O.detect_interfacial(nat)
        assert np.array_equal(itf, ref)
        nf, nif = sim.counts()
        assert nf == int((nat == 0).sum()) and nif == int(((nat == 0) & (ref == 1)).sum())


@pytest.mark.parametrize("tau", [1.0, 0.8])
@pytest.mark.parametrize("name,mk", GEOMS, ids=[g[0] for g in GEOMS])
def test_lb_steps_bit_exact(name, mk, tau):
    """check_every=1, force switched on after step 3 (as the driver does), populations read back."""
    lb = _gpu()
    nat = mk()
    f = [1e-3, -2e-3, 5e-4]
    st = O.LBState(nat, 1.0, tau)
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        assert np.array_equal(sim.lb_populations(), st.n)
        ref_hist = []
        for _ in range(3):
            ref_hist.append(st.step()[1])
        done, conv, h = sim.lb_step(3, tau=tau, check_every=1, target_error=-1.0)
        assert done == 3 and not conv
        assert np.array_equal(h, np.array(ref_hist))
        assert np.array_equal(sim.lb_populations(), st.n)
        st.set_force_uniform(f)
        sim.lb_set_force_uniform(f)
        ref_hist = [st.step()[1] for _ in range(9)]
        done, conv, h = sim.lb_step(9, tau=tau, check_every=1, target_error=-1.0)
        assert done == 9 and sim.t == 12
        assert np.array_equal(h, np.array(ref_hist))
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(rho, st.rho) and np.array_equal(jx, st.jx)
        assert np.array_equal(jy, st.jy) and np.array_equal(jz, st.jz)
        n = sim.lb_populations()
        assert rel_err(n, st.n) <= RTOL
        assert np.array_equal(n, st.n)
        assert (n[:, nat == 1] == 0).all()
        # a further step after the read-back (exercises the redo path)
        e = st.step()[1]
        done, conv, h = sim.lb_step(1, tau=tau, check_every=1, target_error=-1.0)
        assert h[0] == e and np.array_equal(sim.lb_populations(), st.n)


def test_check_every_variants_agree():
    """check_every = 0 / 4 / 1 give the same populations; l2err values on checked steps coincide."""
    lb = _gpu()
    nat = random_nature(9, 6, 5, 0.25, 21)
    outs = []
    for ce in (1, 4, 0):
        with lb.LaboetieGPU(nat) as sim:
            sim.lb_init(1.0)
            sim.lb_set_force_uniform([1e-4, 0, 2e-4])
            done, conv, h = sim.lb_step(17, tau=0.9, check_every=ce, target_error=-1.0)
            assert done == 17 and not conv
            outs.append((sim.lb_populations(), h))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][0], outs[2][0])
    h1, h4, h0 = outs[0][1], outs[1][1], outs[2][1]
    assert np.isnan(h0).all()
    for i in range(17):
        if (i + 1) % 4 == 0:
            assert h4[i] == h1[i]
        else:
            assert np.isnan(h4[i])


def test_force_field_and_upload_paths():
    lb = _gpu()
    nat = random_nature(7, 6, 5, 0.3, 31)
    rng = np.random.default_rng(5)
    fl = nat == 0
    fx, fy, fz = (np.where(fl, rng.normal(0, 1e-4, nat.shape), 0.0) for _ in range(3))
    st = O.LBState(nat, 1.0, 0.7)
    for _ in range(4):
        st.step()
    with lb.LaboetieGPU(nat) as sim:
        # restart from a host state, then switch to a per-node force field (compensate_f_ext style)
        sim.lb_upload(st.n, st.rho, st.jx, st.jy, st.jz)
        st.fx[...], st.fy[...], st.fz[...] = fx, fy, fz
        sim.lb_set_force_field(fx, fy, fz)
        ref = [st.step()[1] for _ in range(6)]
        done, conv, h = sim.lb_step(6, tau=0.7, check_every=1, target_error=-1.0)
        assert np.array_equal(h, np.array(ref))
        assert np.array_equal(sim.lb_populations(), st.n)
        # back to a uniform force: j(t) keeps the field's f/2, the collision takes the new one
        st.fx[...], st.fy[...], st.fz[...] = 0, 0, 0
        st.set_force_uniform([2e-4, 0, 0])
        sim.lb_set_force_uniform([2e-4, 0, 0])
        ref = [st.step()[1] for _ in range(3)]
        done, conv, h = sim.lb_step(3, tau=0.7, check_every=1, target_error=-1.0)
        assert np.array_equal(h, np.array(ref))
        assert np.array_equal(sim.lb_populations(), st.n)
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(jx, st.jx) and np.array_equal(rho, st.rho)


@pytest.mark.parametrize("label,shape", [(-1, (7, 5, 9)), (1, (5, 7, 9))])
def test_compensate_f_ext_workflow(label, shape):
    """compensate_f_ext = T (equilibration.f90:185-188,388-487): particle force field +
    background, central-node probe every step, same exit steps as the oracle-driven loop."""
    lb = _gpu()
    from laboetie_b200 import driver
    nat = O.geometry(label, *shape)
    f = [1e-4, 0.0, 3e-4]
    tau, target, pd = 0.9, 1e-8, 3
    # oracle-driven reference loop
    st = O.LBState(nat, 1.0, tau)
    px, py, pz = shape[0] // 2 + 1, shape[1] // 2 + 1, shape[2] // 2 + 1
    hist, probe, without, t, tf = [], [], False, 0, 0
    while True:
        if without:
            probe.append((t + 1 - tf, st.jx[pz - 1, py - 1, px - 1], st.jy[pz - 1, py - 1, px - 1], st.jz[pz - 1, py - 1, px - 1]))
        rc, e = st.step()
        assert rc == 0
        t += 1
        hist.append(e)
        if e <= target and t > 2:
            if not without:
                without, tf = True, t + 1
                st.fx, st.fy, st.fz, nl = O.compensate_force(nat, f, pd=pd, geometry_label=label)
                assert nl == 19
            else:
                break
        assert t < 20000
    with lb.LaboetieGPU(nat) as sim:
        r = driver.equilibration_compensated(sim, nat, f, tau=tau, target_error=target, particle_diameter=pd,
                                             geometry_label=label)
        assert (r["t_exit"], r["t_fext"]) == (t, tf)
        assert np.array_equal(r["l2err"], np.array(hist))
        assert np.array_equal(r["v_centralnode"], np.array(probe))
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(jx, st.jx) and np.array_equal(jz, st.jz) and np.array_equal(rho, st.rho)


def test_equilibration_exit_steps_match_reference_semantics():
    """Stock lb.in leaves the loop at t=4; a forced slit reproduces exit step and l2err history."""
    lb = _gpu()
    from laboetie_b200 import driver
    nat = O.geometry(1, 1, 1, 102)
    with lb.LaboetieGPU(nat) as sim:
        r = driver.equilibration(sim, [0, 0, 0])
        assert r["rc"] == 0 and r["t_exit"] == 4 and r["t_fext"] == 4 and (r["l2err"] == 0).all()
    nat = O.geometry(1, 3, 2, 14)
    ref = O.equilibration(nat, [1e-5, 0, 0], tau=1.0, target_error=1e-10)
    with lb.LaboetieGPU(nat) as sim:
        r = driver.equilibration(sim, [1e-5, 0, 0], tau=1.0, target_error=1e-10, chunk=97)
        assert (r["t_exit"], r["t_fext"]) == (ref["t_exit"], ref["t_fext"])
        assert np.array_equal(r["l2err"], ref["l2err"])
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(jx, ref["jx"]) and np.array_equal(rho, ref["rho"])
        assert np.array_equal(sim.lb_populations(), ref["n"])


@pytest.mark.parametrize("in_place", [False, True], ids=["two-lattice", "in-place"])
def test_porous_equilibration_to_convergence_and_tracers(in_place):
    """A whole run of the reference's two phases on a porous lattice (config-5 geometry, 30 720 nodes): about 1000
    LB steps through both convergence events of the state machine (force switched on at the first one), then 300
    propagate steps with adsorption.  Exit steps, the complete l2err history, the final populations and moments,
    P and Pads bit for bit; the vacf rows to summation order."""
    lb = _gpu()
    from laboetie_b200 import driver, synthetic as S
    nat = S.porous_spheres(48, 40, 16, radius=5)
    f = [1e-5, 0.0, 2e-5]
    ref = O.equilibration(nat, f, tau=1.0, target_error=1e-9)
    assert ref["rc"] == 0 and ref["t_exit"] > 500 and ref["t_fext"] == 4
    itf = O.detect_interfacial(nat)
    mp = O.MPState(nat, itf, ref["rho"], ref["jx"], ref["jy"], ref["jz"], f, 0.01, 0.1, 0.01)
    ref_v = np.array([mp.propagate()[1] for _ in range(300)])
    with lb.LaboetieGPU(nat) as sim:
        if in_place:
            sim.lb_set_in_place(True)
        r = driver.equilibration(sim, f, tau=1.0, target_error=1e-9, chunk=233)
        assert (r["rc"], r["t_exit"], r["t_fext"]) == (0, ref["t_exit"], ref["t_fext"])
        assert np.array_equal(r["l2err"], ref["l2err"])
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(rho, ref["rho"]) and np.array_equal(jx, ref["jx"])
        assert np.array_equal(jy, ref["jy"]) and np.array_equal(jz, ref["jz"])
        assert np.array_equal(sim.lb_populations(), ref["n"])
        d = driver.drop_tracers(sim, f, 0.01, 0.1, 0.01, max_steps=300, chunk=128)
        assert d["steps"] == 300 and not d["converged"]
        tol = sum_rtol(18 * nat.size)
        assert rel_err(d["vacf"][0], mp.vacf0) <= tol
        assert (np.abs(d["vacf"][1:] - ref_v) <= tol * np.abs(mp.vacf0).max()).all()
        P, A = sim.mp_download()
        assert np.array_equal(P, mp.P[0]) and np.array_equal(A, mp.Pads[0])


def test_tuto_config1_against_oracle_derived_golden():
    """BASELINE config 1: tuto 1x50x50 one-disk geometry, flow equilibration then moment propagation, against
    tests/golden/tuto_cfg1_oracle.npz -- ORACLE output (tests/golden/make_golden.py), a regression pin, not a reference pin."""
    lb = _gpu()
    from laboetie_b200 import driver
    import os
    nat = O.read_geom_in(os.path.join(GOLDEN, "geom.in_chromat_1disks-dia10-1x50x50_v1"), 1, 50, 50)
    f = [0.0, 1e-5, 0.0]
    g = np.load(os.path.join(GOLDEN, "tuto_cfg1_oracle.npz"))
    with lb.LaboetieGPU(nat) as sim:
        r = driver.equilibration(sim, f, tau=1.0, target_error=1e-10)
        assert r["t_exit"] == int(g["t_exit"]) and r["t_fext"] == int(g["t_fext"])
        assert np.array_equal(r["l2err"], g["l2err"])
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(rho, g["rho"]) and np.array_equal(jy, g["jy"]) and np.array_equal(jz, g["jz"])
        prof = sim.lb_profiles(2)
        assert rel_err(prof, g["prof_z"]) <= RTOL
        d = driver.drop_tracers(sim, f, 0.01, 0.1, 0.01, max_steps=int(g["mp_steps"]))
        P, A = sim.mp_download()
        assert np.array_equal(P, g["P"]) and np.array_equal(A, g["Pads"])
        scale = np.abs(g["vacf"]).max(axis=0)
        assert (np.abs(d["vacf"] - g["vacf"]) <= RTOL * scale).all()


MP_GEOMS = GEOMS[:9] + [GEOMS[14]]


@pytest.mark.parametrize("nbt", ["0", "1"], ids=["rank-lookups", "neighbour-table"])
@pytest.mark.parametrize("ka,kd", [(0.1, 0.01), (0.0, 0.0), (0.05, 0.0)])
@pytest.mark.parametrize("name,mk", MP_GEOMS, ids=[g[0] for g in MP_GEOMS])
def test_moment_propagation_bit_exact(name, mk, ka, kd, nbt, monkeypatch):
    """Both ways the propagate kernel finds a node's neighbours -- 18 rank lookups, or the static
    neighbour table with its periodic-x-seam fallback (the library picks by lattice shape; forced
    here) -- must give the oracle's P, Pads bit for bit."""
    lb = _gpu()
    monkeypatch.setenv("LBG_MP_NBT", nbt)
    nat = mk()
    itf = O.detect_interfacial(nat)
    f = [1e-4, 2e-4, -1e-4]
    st = O.LBState(nat)
    for _ in range(3):
        st.step()
    st.set_force_uniform(f)
    for _ in range(12):
        st.step()
    Db = 0.01
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, f, Db, ka, kd)
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_step(3, check_every=0)
        sim.lb_set_force_uniform(f)
        sim.lb_step(12, check_every=0)
        v0 = sim.mp_init(Db, ka, kd, f)
        assert rel_err(v0, mp.vacf0) <= RTOL
        P, A = sim.mp_download()
        assert np.array_equal(P, mp.P[0]) and (A == 0).all()
        ref_v = []
        for _ in range(21):
            rc, v, conv = mp.propagate()
            assert rc == 0
            ref_v.append(v)
        done, conv, v = sim.mp_step(21)
        assert done == 21
        P, A = sim.mp_download()
        assert np.array_equal(P, mp.P[0])
        assert np.array_equal(A, mp.Pads[0])
        ref_v = np.array(ref_v)
        scale = np.maximum(np.abs(ref_v).max(axis=0), 1e-300)
        assert (np.abs(v - ref_v) <= RTOL * scale).all()


@pytest.mark.parametrize("rows", ["2", "3", "5"])
def test_moment_propagation_strip_order(rows, monkeypatch):
    """The L2-friendly strip order of the propagate kernel (forced on a small lattice) changes nothing per node."""
    lb = _gpu()
    monkeypatch.setenv("LBG_MP_STRIP_ROWS", rows)
    nat = random_nature(9, 11, 7, 0.3, 51)
    itf = O.detect_interfacial(nat)
    f = [1e-4, 2e-4, -1e-4]
    st = O.LBState(nat)
    st.set_force_uniform(f)
    for _ in range(10):
        st.step()
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, f, 0.01, 0.1, 0.01)
    ref_v = np.array([mp.propagate()[1] for _ in range(15)])
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f)
        sim.lb_step(10, check_every=0)
        sim.mp_init(0.01, 0.1, 0.01, f)
        done, conv, v = sim.mp_step(15)
        P, A = sim.mp_download()
        assert np.array_equal(P, mp.P[0]) and np.array_equal(A, mp.Pads[0])
        assert (np.abs(v - ref_v) <= RTOL * np.abs(mp.vacf0).max()).all()


@pytest.mark.parametrize("tpc", ["0", "2"], ids=["static-schedule", "dynamic-schedule"])
@pytest.mark.parametrize("rows", ["2", "5"])
def test_lb_strip_order(rows, tpc, monkeypatch):
    """The strip order of the Phase-A step kernel (on by itself only on large wall-rich lattices; forced here, with
    both tile schedules) changes nothing: l2err history, populations and moments bit for bit."""
    lb = _gpu()
    monkeypatch.setenv("LBG_LB_STRIP_ROWS", rows)
    monkeypatch.setenv("LBG_LB_PIPE", "0")
    monkeypatch.setenv("LBG_LB_TPC", tpc)
    nat = random_nature(37, 11, 7, 0.3, 52)
    f = [1e-4, 2e-4, -1e-4]
    st = O.LBState(nat, 1.0, 0.8)
    st.set_force_uniform(f)
    ref = [st.step()[1] for _ in range(9)]
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f)
        done, conv, h = sim.lb_step(9, tau=0.8, check_every=1, target_error=-1.0)
        assert done == 9 and np.array_equal(h, np.array(ref))
        assert np.array_equal(sim.lb_populations(), st.n)
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(rho, st.rho) and np.array_equal(jx, st.jx) and np.array_equal(jz, st.jz)


def test_phase_b_from_host_moments():
    """lbg_mp_init_from_moments: Phase B started from the driver's density / momentum arrays (what
    drop_tracers.f90:63-105 reads from node%solventdensity/solventflux) is bit-identical to Phase B started
    from the resident Lattice-Boltzmann state, and to the oracle; no LB state is needed on that handle."""
    lb = _gpu()
    nat = random_nature(34, 6, 7, 0.3, 77)
    itf = O.detect_interfacial(nat)
    f = [2e-4, -1e-4, 3e-4]
    st = O.LBState(nat, 1.0, 0.9)
    st.set_force_uniform(f)
    for _ in range(9):
        st.step()
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, f, 0.02, 0.08, 0.03)
    ref_v = np.array([mp.propagate()[1] for _ in range(12)])
    with lb.LaboetieGPU(nat) as a:
        a.lb_init(1.0)
        a.lb_set_force_uniform(f)
        a.lb_step(9, tau=0.9, check_every=0)
        rho, jx, jy, jz = a.lb_moments()
        assert np.array_equal(rho, st.rho) and np.array_equal(jx, st.jx)
        v0a = a.mp_init(0.02, 0.08, 0.03, f)
        _, _, va = a.mp_step(12)
        Pa, Aa = a.mp_download()
    for nbt in ("0", "1"):
        os.environ["LBG_MP_NBT"] = nbt
        try:
            with lb.LaboetieGPU(nat) as b:
                v0b = b.mp_init_from_moments(rho, jx, jy, jz, 0.02, 0.08, 0.03, f)
                _, _, vb = b.mp_step(12)
                Pb, Ab = b.mp_download()
                with pytest.raises(lb.LbgError):      # no Lattice-Boltzmann state on this handle
                    b.lb_step(1)
        finally:
            del os.environ["LBG_MP_NBT"]
        assert np.array_equal(v0a, v0b) and np.array_equal(va, vb)
        assert np.array_equal(Pa, Pb) and np.array_equal(Aa, Ab)
        assert np.array_equal(Pb, mp.P[0]) and np.array_equal(Ab, mp.Pads[0])
        assert (np.abs(vb - ref_v) <= RTOL * np.abs(mp.vacf0).max()).all()


def test_mp_convergence_step_matches():
    """Bulk fluid at rest: vacf(t>=1) = 0, so propagate reports convergence at it=3 (it>2, :284)."""
    lb = _gpu()
    from laboetie_b200 import driver
    nat = O.geometry(-1, 4, 5, 3)
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        d = driver.drop_tracers(sim, [0, 0, 0], 0.0123, 0.0, 0.0, max_steps=-1, chunk=50)
        assert d["converged"] and d["steps"] == 3
        assert np.allclose(d["vacf"][0], 2 * 0.0123, rtol=1e-13)
        assert np.abs(d["vacf"][1:]).max() < 1e-17


def test_profiles_flux_probe():
    lb = _gpu()
    nat = random_nature(9, 7, 6, 0.3, 41)
    st = O.LBState(nat, 1.0, 0.9)
    st.set_force_uniform([1e-3, 2e-3, -1e-3])
    for _ in range(8):
        st.step()
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_set_force_uniform([1e-3, 2e-3, -1e-3])
        sim.lb_step(8, tau=0.9, check_every=0)
        for axis in range(3):
            ref = O.profiles(st.rho, st.jx, st.jy, st.jz, axis)
            got = sim.lb_profiles(axis)
            for c in range(4):
                assert rel_err(got[:, c], ref[:, c]) <= RTOL
        assert rel_err(sim.lb_total_flux(), O.total_flux(st.jx, st.jy, st.jz)) <= RTOL
        i, j, k = 3, 2, 4
        assert np.array_equal(sim.lb_probe(i, j, k), [st.jx[k, j, i], st.jy[k, j, i], st.jz[k, j, i], st.rho[k, j, i]])


def test_error_codes_mirror_reference_stops():
    lb = _gpu()
    nat = O.geometry(1, 4, 4, 8)
    with pytest.raises(lb.LbgError) as e:
        lb.LaboetieGPU(np.ones((3, 3, 3), np.int8))
    assert e.value.status == 6                      # all solid
    with lb.LaboetieGPU(nat) as sim:
        with pytest.raises(lb.LbgError) as e:
            sim.lb_step(1)
        assert e.value.status == 8                  # no lb_init yet
        sim.lb_init(1.0)
        with pytest.raises(lb.LbgError) as e:
            sim.lb_step(1, tau=0.4)
        assert e.value.status == 3                  # relaxation_time < 0.5
        with pytest.raises(lb.LbgError) as e:
            sim.mp_init(0.0, 0.1, 0.01, [0, 0, 0])
        assert e.value.status == 4                  # tracer_Db invalid
        with pytest.raises(lb.LbgError) as e:
            sim.mp_init(0.01, -0.1, 0.01, [0, 0, 0])
        assert e.value.status == 5
        # a huge force drives populations negative: the reference ERROR STOPs at that step
        st = O.LBState(nat)
        st.set_force_uniform([0.9, 0, 0])
        t_neg = None
        for t in range(1, 20):
            rc, _ = st.step()
            if rc:
                t_neg = t
                break
        assert t_neg is not None
        sim.lb_set_force_uniform([0.9, 0, 0])
        with pytest.raises(lb.LbgError) as e:
            sim.lb_step(50, check_every=1, target_error=-1.0)
        assert e.value.status == 1 and sim.last_steps_done == t_neg
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        st = O.LBState(nat)
        mp = O.MPState(nat, O.detect_interfacial(nat), st.rho, st.jx, st.jy, st.jz, [0, 0, 0], 0.01, 0.99, 0.5)
        assert mp.propagate()[0] == 1               # the oracle hits 'restpart is negative' too
        sim.mp_init(0.01, 0.99, 0.5, [0, 0, 0])     # ka so large that the remaining fraction goes negative
        with pytest.raises(lb.LbgError) as e:
            sim.mp_step(1)
        assert e.value.status == 2
        with pytest.raises(lb.LbgError) as e:
            sim.lb_step(1)
        assert e.value.status == 8                  # populations were released by mp_init


def test_medium_lattice_against_oracle():
    """64x64x32 slit (config 2 cross-section): 20 LB + 20 MP steps against the oracle."""
    lb = _gpu()
    nat = O.geometry(1, 64, 64, 32)
    itf = O.detect_interfacial(nat)
    f = [1e-6, 0, 0]
    st = O.LBState(nat)
    st.set_force_uniform(f)
    ref = [st.step()[1] for _ in range(20)]
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, f, 0.01, 0.1, 0.01)
    ref_v = np.array([mp.propagate()[1] for _ in range(20)])
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f)
        done, conv, h = sim.lb_step(20, check_every=1, target_error=-1.0)
        assert np.array_equal(h, np.array(ref))
        assert np.array_equal(sim.lb_populations(), st.n)
        v0 = sim.mp_init(0.01, 0.1, 0.01, f)
        tol = sum_rtol(18 * nat.size)          # 2.4e6 terms in one accumulator on the CPU side
        assert rel_err(v0, mp.vacf0) <= tol
        done, conv, v = sim.mp_step(20)
        P, A = sim.mp_download()
        assert np.array_equal(P, mp.P[0]) and np.array_equal(A, mp.Pads[0])
        # vacf_y, vacf_z cancel to rounding noise in this geometry: compare on the scale of vacf(0)
        assert (np.abs(v - ref_v) <= tol * np.abs(mp.vacf0).max()).all()


def test_full_size_config2_properties():
    """BASELINE config 2 at full size (64x64x256 slit): size-independent properties."""
    lb = _gpu()
    nat = O.geometry(1, 64, 64, 256)
    f = [1e-6, 0, 0]
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f)
        sim.lb_step(200, check_every=0)
        n = sim.lb_populations()
        rho, jx, jy, jz = sim.lb_moments()
        nf = int((nat == 0).sum())
        assert (n[:, nat == 1] == 0).all()                      # solid populations stay 0
        assert n.min() >= 0
        assert abs(n.sum() - nf) <= 1e-12 * nf                  # mass conservation
        # translation invariance in x and y of the slit: every (x,y) column is identical
        assert np.array_equal(jx, np.broadcast_to(jx[:, :1, :1], jx.shape))
        assert np.abs(jy).max() < 1e-13 and np.abs(jz).max() < 1e-13   # cancel up to rounding of the l-ordered sum
        # momentum balance: d/dt sum(jx) -> 0 as the Poiseuille profile builds; sign and symmetry in z
        prof = sim.lb_profiles(2)[:, 0]
        assert (prof[1:-1] > 0).all() and prof[0] == 0 and prof[-1] == 0
        assert np.allclose(prof, prof[::-1], rtol=1e-12, atol=0)
        v0 = sim.mp_init(0.01, 0.1, 0.01, f)
        P0, A0 = sim.mp_download()
        tot0 = (P0 + A0).sum(axis=(0, 1, 2))
        done, conv, v = sim.mp_step(100)
        P, A = sim.mp_download()
        tot = (P + A).sum(axis=(0, 1, 2))
        assert np.allclose(tot, tot0, rtol=0, atol=1e-11 * np.abs(P0).sum())   # sum(P + Pads) conserved up to rounding over 100 steps
        assert (P[nat == 1] == 0).all() and (A[itf_not(nat)] == 0).all()
        assert np.isfinite(v).all()


def test_full_size_config3_properties():
    """BASELINE config 3 at full size (256^3 BCC spheres, adsorbing tracer): size-independent properties.
    At this size the propagate kernel runs on its neighbour-table path by default."""
    lb = _gpu()
    L = 256
    nat = O.geometry(3, L, L, L)
    f = [1e-6, 0, 0]
    nf = int((nat == 0).sum())
    with lb.LaboetieGPU(nat) as sim:
        assert sim.counts()[0] == nf
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f)
        done, conv, h = sim.lb_step(60, check_every=1, target_error=-1.0)
        assert done == 60 and np.isfinite(h).all() and (h >= 0).all()
        rho, jx, jy, jz = sim.lb_moments()
        assert (rho[nat == 1] == 0).all() and (jx[nat == 1] == 0).all()
        assert abs(rho.sum() - nf) <= 1e-12 * nf                         # mass conservation
        # the BCC cell is symmetric under y <-> z, and so is the forcing along x
        tol = 1e-10 * np.abs(jx).max()   # the l-ordered sums treat y and z in a different order: rounding only
        assert np.allclose(jx, jx.transpose(1, 0, 2), rtol=1e-9, atol=tol)
        assert np.allclose(jy, jz.transpose(1, 0, 2), rtol=1e-9, atol=tol)
        assert jx.sum() > 0
        # mirror planes of the cell: total transverse flux vanishes up to rounding
        assert abs(jy.sum()) <= 1e-9 * jx.sum() and abs(jz.sum()) <= 1e-9 * jx.sum()
        v0 = sim.mp_init(0.01, 0.1, 0.01, f)
        P0, A0 = sim.mp_download()
        assert (A0 == 0).all() and (P0[nat == 1] == 0).all()
        tot0 = P0.sum(axis=(0, 1, 2))
        done, conv, v = sim.mp_step(40)
        assert done == 40 and np.isfinite(v).all()
        P, A = sim.mp_download()
        tot = (P + A).sum(axis=(0, 1, 2))
        assert np.allclose(tot, tot0, rtol=0, atol=1e-11 * np.abs(P0).sum())   # sum(P + Pads) is conserved
        assert (P[nat == 1] == 0).all() and (A[itf_not(nat)] == 0).all()
        assert (A[~itf_not(nat)] != 0).any()
        # vacf decays from vacf(0) > 0 along every axis (diffusive tracer)
        assert (v0 > 0).all() and (np.abs(v[-1]) < v0).all()


def test_full_size_config5_properties(monkeypatch):
    """BASELINE config 5 at the size bench.py measures (1024x1024x128 overlapping-sphere porous medium, 80 M fluid
    nodes), on the code paths the library picks there by itself (dynamic tile schedule, strip order, neighbour
    table, compact adsorbed storage): size-independent properties, and agreement bit for bit between independent
    kernels -- two-lattice vs in-place Phase A, neighbour table vs rank lookups in Phase B."""
    lb = _gpu()
    from laboetie_b200 import synthetic as S
    for v in ("LBG_MP_NBT", "LBG_LB_MINB", "LBG_LB_PIPE", "LBG_LB_TPC", "LBG_LB_STRIP_ROWS", "LBG_MP_STRIP_ROWS"):
        monkeypatch.delenv(v, raising=False)
    builder, lx, ly, lz, f, _ = S.WORKLOADS["cfg5w"]
    nat = builder(lx, ly, lz)
    fluid = nat == 0
    nf = int(fluid.sum())
    f = [1e-5, 0.0, 2e-5]
    steps = 7          # odd: the in-place buffer ends in its swapped layout
    with lb.LaboetieGPU(nat) as sim:
        assert sim.counts()[0] == nf
        sim.lb_init(1.0)
        d1, _, h1 = sim.lb_step(3, tau=0.9, check_every=1, target_error=-1.0)
        sim.lb_set_force_uniform(f)
        d2, _, h2 = sim.lb_step(steps - 3, tau=0.9, check_every=1, target_error=-1.0)
        assert d1 + d2 == steps and sim.info("lb_variant") == 13      # plain kernel, dynamic schedule, 3 CTAs per SM
        rho, jx, jy, jz = sim.lb_moments()
        assert (rho[~fluid] == 0).all() and (jz[~fluid] == 0).all()
        assert abs(rho.sum() - nf) <= 1e-12 * nf                         # mass conservation
        assert jx.sum() > 0 and jz.sum() > 0
        Db, ka, kd = 0.01, 0.1, 0.01
        v0 = sim.mp_init(Db, ka, kd, f)
        assert sim.info("mp_neighbour_table") == 1
        P0, _ = sim.mp_download(want_ads=False)
        tot0, scale = P0.sum(axis=(0, 1, 2)), np.abs(P0).sum()
        del P0
        done, _, v = sim.mp_step(5)
        assert done == 5 and np.isfinite(v).all() and (v0 > 0).all()
        P, A = sim.mp_download()
        assert (P[~fluid] == 0).all() and (A[~fluid] == 0).all() and (A != 0).any()
        assert np.allclose((P.sum(axis=(0, 1, 2)) + A.sum(axis=(0, 1, 2))), tot0, rtol=0, atol=1e-11 * scale)   # sum(P + Pads) conserved
    # the same Phase A in place (AA kernels), the same Phase B through rank lookups
    monkeypatch.setenv("LBG_MP_NBT", "0")
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_set_in_place(True)
        sim.lb_init(1.0)
        _, _, g1 = sim.lb_step(3, tau=0.9, check_every=1, target_error=-1.0)
        sim.lb_set_force_uniform(f)
        _, _, g2 = sim.lb_step(steps - 3, tau=0.9, check_every=1, target_error=-1.0)
        assert np.array_equal(g1, h1) and np.array_equal(g2, h2)
        for a, b in zip(sim.lb_moments(), (rho, jx, jy, jz)):
            assert np.array_equal(a, b)
        del rho, jx, jy, jz
        w0 = sim.mp_init(Db, ka, kd, f)
        assert sim.info("mp_neighbour_table") == 0
        _, _, w = sim.mp_step(5)
        P2, A2 = sim.mp_download()
        assert np.array_equal(P2, P) and np.array_equal(A2, A)
        assert np.allclose(w0, v0, rtol=1e-9, atol=0) and np.allclose(w, v, rtol=0, atol=1e-9 * np.abs(v0).max())


def itf_not(nat):
    itf = O.detect_interfacial(nat)
    return ~((itf == 1) & (nat == 0))


# --------------------------------------------------------------------------- in-place (AA) mode
AA_GEOMS = [g for g in GEOMS if g[0] in ("rand6x5x7", "line1x1x22", "rand2x3x2", "rand33x7x5", "plane1x12x9", "slit8x8x16",
                                         "onefluid3x3x3", "solidplane6x5x8", "rand40x9x6")]


@pytest.mark.parametrize("tau", [1.0, 0.8])
@pytest.mark.parametrize("name,mk", AA_GEOMS, ids=[g[0] for g in AA_GEOMS])
def test_in_place_lb_steps_bit_exact(name, mk, tau):
    """AA pattern (one population buffer): same l2err history, populations and moments as the oracle,
    read back in both buffer layouts (after odd and even step counts), across a force switch."""
    lb = _gpu()
    nat = mk()
    f = [1e-3, -2e-3, 5e-4]
    st = O.LBState(nat, 1.0, tau)
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_set_in_place(True)
        sim.lb_init(1.0)
        assert np.array_equal(sim.lb_populations(), st.n)
        ref = [st.step()[1] for _ in range(3)]
        done, conv, h = sim.lb_step(3, tau=tau, check_every=1, target_error=-1.0)
        assert done == 3 and np.array_equal(h, np.array(ref))
        assert np.array_equal(sim.lb_populations(), st.n)          # swapped layout (odd number of steps)
        st.set_force_uniform(f)
        sim.lb_set_force_uniform(f)
        ref = [st.step()[1] for _ in range(9)]
        done, conv, h = sim.lb_step(9, tau=tau, check_every=1, target_error=-1.0)
        assert done == 9 and sim.t == 12 and np.array_equal(h, np.array(ref))
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(rho, st.rho) and np.array_equal(jx, st.jx)
        assert np.array_equal(jy, st.jy) and np.array_equal(jz, st.jz)
        assert np.array_equal(sim.lb_populations(), st.n)          # normal layout (even number of steps)
        # unchecked steps, then a checked one
        for _ in range(4):
            st.step()
        sim.lb_step(4, tau=tau, check_every=0)
        e = st.step()[1]
        done, conv, h = sim.lb_step(1, tau=tau, check_every=1, target_error=-1.0)
        assert h[0] == e and np.array_equal(sim.lb_populations(), st.n)


def test_in_place_equilibration_and_tracers():
    """The whole driver flow in AA mode: exit steps, final moments, then Phase B from the resident state."""
    lb = _gpu()
    from laboetie_b200 import driver
    nat = O.geometry(1, 3, 2, 14)
    ref = O.equilibration(nat, [1e-5, 0, 0], tau=1.0, target_error=1e-10)
    itf = O.detect_interfacial(nat)
    mp = O.MPState(nat, itf, ref["rho"], ref["jx"], ref["jy"], ref["jz"], [1e-5, 0, 0], 0.01, 0.1, 0.01)
    ref_v = np.array([mp.propagate()[1] for _ in range(12)])
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_set_in_place(True)
        r = driver.equilibration(sim, [1e-5, 0, 0], tau=1.0, target_error=1e-10, chunk=97)
        assert (r["t_exit"], r["t_fext"]) == (ref["t_exit"], ref["t_fext"])
        assert np.array_equal(r["l2err"], ref["l2err"])
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(jx, ref["jx"]) and np.array_equal(rho, ref["rho"])
        assert np.array_equal(sim.lb_populations(), ref["n"])
        d = driver.drop_tracers(sim, [1e-5, 0, 0], 0.01, 0.1, 0.01, max_steps=12)
        P, A = sim.mp_download()
        assert np.array_equal(P, mp.P[0]) and np.array_equal(A, mp.Pads[0])
        assert (np.abs(d["vacf"][1:] - ref_v) <= RTOL * np.abs(mp.vacf0).max()).all()


def test_in_place_force_field_upload_and_negative_guard():
    lb = _gpu()
    nat = random_nature(7, 6, 5, 0.3, 31)
    rng = np.random.default_rng(5)
    fl = nat == 0
    fx, fy, fz = (np.where(fl, rng.normal(0, 1e-4, nat.shape), 0.0) for _ in range(3))
    st = O.LBState(nat, 1.0, 0.7)
    for _ in range(4):
        st.step()
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_set_in_place(True)
        sim.lb_upload(st.n, st.rho, st.jx, st.jy, st.jz)
        st.fx[...], st.fy[...], st.fz[...] = fx, fy, fz
        sim.lb_set_force_field(fx, fy, fz)
        ref = [st.step()[1] for _ in range(5)]
        done, conv, h = sim.lb_step(5, tau=0.7, check_every=1, target_error=-1.0)
        assert np.array_equal(h, np.array(ref)) and np.array_equal(sim.lb_populations(), st.n)
        st.fx[...], st.fy[...], st.fz[...] = 0, 0, 0
        st.set_force_uniform([2e-4, 0, 0])
        sim.lb_set_force_uniform([2e-4, 0, 0])
        ref = [st.step()[1] for _ in range(4)]
        done, conv, h = sim.lb_step(4, tau=0.7, check_every=1, target_error=-1.0)
        assert np.array_equal(h, np.array(ref)) and np.array_equal(sim.lb_populations(), st.n)
    nat = O.geometry(1, 4, 4, 8)
    st = O.LBState(nat)
    st.set_force_uniform([0.9, 0, 0])
    t_neg = next(t for t in range(1, 20) if st.step()[0])
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_set_in_place(True)
        sim.lb_init(1.0)
        sim.lb_set_force_uniform([0.9, 0, 0])
        with pytest.raises(lb.LbgError) as e:
            sim.lb_step(50, check_every=1, target_error=-1.0)
        assert e.value.status == 1 and sim.last_steps_done == t_neg


@pytest.mark.parametrize("label,shape", [(-1, (5, 4, 3)), (1, (4, 3, 9)), (2, (11, 11, 4)), (2, (26, 26, 3)), (3, (9, 9, 9)),
                                          (3, (16, 16, 16)), (3, (33, 33, 33))])
def test_device_side_geometry_builders(label, shape):
    """geometryLabel -1/1/2/3 built on the device equal the reference's builders (oracle),
    whole lattice and as a z-slab; flow on the device-built lattice equals flow on the host-built one."""
    lb = _gpu()
    nat = O.geometry(label, *shape)
    with lb.LaboetieGPU(label=label, shape=shape) as sim:
        assert np.array_equal(sim.nature(), nat)
        assert np.array_equal(sim.interfacial(), O.detect_interfacial(nat))
        sim.lb_init(1.0)
        sim.lb_set_force_uniform([1e-4, 0, 2e-4])
        sim.lb_step(5, check_every=0)
        n_dev = sim.lb_populations()
    st = O.LBState(nat)
    st.set_force_uniform([1e-4, 0, 2e-4])
    for _ in range(5):
        st.step()
    assert np.array_equal(n_dev, st.n)
    lz = shape[2]
    if lz >= 4:
        k0, nzl = 1, lz - 2
        with lb.LaboetieGPU(label=label, shape=shape, k0=k0, nzl=nzl) as slab:
            assert np.array_equal(slab.nature(), nat[k0:k0 + nzl])
            assert np.array_equal(slab.interfacial(), O.detect_interfacial(nat)[k0:k0 + nzl])


# --------------------------------------------------------------------------- BASELINE-shaped lattices
def _bench_geoms():
    from laboetie_b200 import synthetic as S
    return [
        # (id, builder, force, neighbour table expected by the library's own heuristic)
        ("porous128x96x24", lambda: S.porous_spheres(128, 96, 24, radius=6), [1e-6, 0.0, 0.0], 1),   # BASELINE config 5
        ("bcc64", lambda: S.bcc(64, 64, 64), [1e-6, 0.0, 0.0], 1),                                   # config 3
        ("cylinder65x65x16", lambda: S.cylinder(65, 65, 16), [0.0, 0.0, 1e-6], 1),                   # config 4
        ("bernoulli96x64x16", lambda: S.bernoulli(96, 64, 16), [1e-6, 0.0, 0.0], 1),                 # cfg5b
        ("slit64x64x24", lambda: S.slit(64, 64, 24), [1e-6, 0.0, 0.0], 1),                           # config 2 (phi < 0.95)
    ]


BENCH_GEOMS = _bench_geoms()


@pytest.mark.parametrize("in_place", [False, True], ids=["two-lattice", "in-place"])
@pytest.mark.parametrize("name,mk,f,want_nbt", BENCH_GEOMS, ids=[g[0] for g in BENCH_GEOMS])
def test_benchmark_shaped_lattices_against_oracle(name, mk, f, want_nbt, in_place, monkeypatch):
    """The geometries bench.py measures (overlapping-sphere porous medium, BCC, cylinder, Bernoulli noise, slit),
    at sizes the oracle runs in seconds, on the code paths the LIBRARY picks by itself (no LBG_MP_NBT / LBG_LB_*
    overrides: the default neighbour-table heuristic, with periodic-x-seam nodes present): 12 LB steps with the
    per-step check across the force switch, then 12 propagate steps with adsorption, everything per node
    bit for bit, l2err history bit for bit, vacf to summation order."""
    lb = _gpu()
    for v in ("LBG_MP_NBT", "LBG_LB_MINB", "LBG_LB_PIPE", "LBG_LB_TPC", "LBG_MP_TPC", "LBG_MP_STRIP_ROWS"):
        monkeypatch.delenv(v, raising=False)
    nat = mk()
    itf = O.detect_interfacial(nat)
    tau, Db, ka, kd = 0.9, 0.01, 0.1, 0.01
    st = O.LBState(nat, 1.0, tau)
    ref_h = [st.step()[1] for _ in range(4)]
    st.set_force_uniform(f)
    ref_h += [st.step()[1] for _ in range(8)]
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, f, Db, ka, kd)
    ref_v = np.array([mp.propagate()[1] for _ in range(12)])
    with lb.LaboetieGPU(nat) as sim:
        if in_place:
            sim.lb_set_in_place(True)
        assert np.array_equal(sim.interfacial(), itf)
        sim.lb_init(1.0)
        d1, _, h1 = sim.lb_step(4, tau=tau, check_every=1, target_error=-1.0)
        sim.lb_set_force_uniform(f)
        d2, _, h2 = sim.lb_step(8, tau=tau, check_every=1, target_error=-1.0)
        assert d1 == 4 and d2 == 8
        assert np.array_equal(np.concatenate([h1, h2]), np.array(ref_h))
        rho, jx, jy, jz = sim.lb_moments()
        assert np.array_equal(rho, st.rho) and np.array_equal(jx, st.jx)
        assert np.array_equal(jy, st.jy) and np.array_equal(jz, st.jz)
        assert np.array_equal(sim.lb_populations(), st.n)
        v0 = sim.mp_init(Db, ka, kd, f)
        assert sim.info("mp_neighbour_table") == want_nbt
        tol = sum_rtol(18 * nat.size)
        assert rel_err(v0, mp.vacf0) <= tol
        done, conv, v = sim.mp_step(12)
        assert done == 12 and not conv
        P, A = sim.mp_download()
        assert np.array_equal(P, mp.P[0])
        assert np.array_equal(A, mp.Pads[0])
        assert (np.abs(v - ref_v) <= tol * np.abs(mp.vacf0).max()).all()


def test_plane_slices_of_the_moments():
    """lbg_lb_slice (the 2-D field outputs of equilibration.f90:526-548) equals the same plane of the full read-back."""
    lb = _gpu()
    nat = random_nature(13, 7, 9, 0.3, 71)
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_set_force_uniform([1e-4, -1e-4, 2e-4])
        sim.lb_step(6, tau=0.9, check_every=1, target_error=-1.0)
        full = sim.lb_moments()
        for axis, idx in ((0, 0), (0, 12), (1, 3), (2, 8), (2, 0)):
            got = sim.lb_slice(axis, idx)
            for a, b in zip(got, full):
                ref = b[:, :, idx] if axis == 0 else (b[:, idx, :] if axis == 1 else b[idx])
                assert np.array_equal(a, ref)
        with pytest.raises(lb.LbgError):
            sim.lb_slice(0, 13)


def test_asynchronous_moments_read_back():
    """lbg_lb_download_moments_async: Phase B runs while density / momentum cross PCIe; same arrays as the blocking call."""
    lb = _gpu()
    nat = random_nature(21, 9, 8, 0.3, 61)
    f = [1e-4, 0, 2e-4]
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f)
        sim.lb_step(7, tau=0.9, check_every=1, target_error=-1.0)
        ref = sim.lb_moments()
        out = [np.full(nat.shape, np.nan) for _ in range(4)]
        sim.lb_moments_async(out)
        v0 = sim.mp_init(0.01, 0.1, 0.01, f)
        done, _, v = sim.mp_step(5)
        P, A = sim.mp_download()          # shares the staging buffer: must wait for the queued read-back itself
        sim.wait_transfers()
        for a, b in zip(out, ref):
            assert np.array_equal(a, b)
    with lb.LaboetieGPU(nat) as sim:     # same Phase B without the overlapped read-back
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f)
        sim.lb_step(7, tau=0.9, check_every=1, target_error=-1.0)
        assert np.array_equal(sim.mp_init(0.01, 0.1, 0.01, f), v0)
        _, _, v2 = sim.mp_step(5)
        P2, A2 = sim.mp_download()
        assert np.array_equal(v, v2) and np.array_equal(P, P2) and np.array_equal(A, A2)


def test_shape_checks_in_the_python_binding():
    lb = _gpu()
    nat = random_nature(6, 5, 4, 0.2, 3)
    with lb.LaboetieGPU(nat) as sim:
        sim.lb_init(1.0)
        bad = np.zeros((5, 5, 6))
        with pytest.raises(lb.LbgError):
            sim.lb_set_force_field(bad, bad, bad)
        with pytest.raises(lb.LbgError):
            sim.lb_upload(np.zeros((19,) + bad.shape), bad, bad, bad, bad)
        with pytest.raises(lb.LbgError):
            sim.info("no-such-key")
