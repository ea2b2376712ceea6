"""Multi-GPU parity (needs >= 2 GPUs on the box): z-slab handles with NCCL halo exchange against the
single-GPU result.  Populations, P and l2err must be bit-identical (no cross-node sums are involved
per node; l2err is a max); vacf differs by summation order only.

Both ranks live in this one process, one Python thread per rank (ctypes releases the GIL inside the
library), which exercises the same C-ABI calls a one-process-per-GPU launch makes.
"""
import threading

import numpy as np
import pytest

from tests.util import RTOL, random_nature

pytestmark = pytest.mark.gpu


def _ndev():
    from laboetie_b200 import api
    import ctypes
    n = ctypes.c_int()
    api.load_library().lbg_device_count(ctypes.byref(n))
    return n.value


def _run_ranks(nranks, fn):
    out, err = [None] * nranks, [None] * nranks

    def work(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa: BLE001
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("shape,nranks", [((16, 12, 21), 2), ((1, 9, 14), 2), ((33, 5, 8), 2), ((8, 8, 9), 4),
                                          ((6, 5, 10, "slit"), 2)])
@pytest.mark.parametrize("tau,nbt,in_place", [(1.0, "0", False), (0.8, "1", False), (0.8, "1", True), (1.0, "0", True)],
                         ids=["tau1-lookups", "tau0.8-table", "tau0.8-table-in-place", "tau1-lookups-in-place"])
def test_slabs_match_single_gpu(shape, nranks, tau, nbt, in_place, monkeypatch):
    import laboetie_b200 as lb
    from laboetie_b200 import api, slab
    if _ndev() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    monkeypatch.setenv("LBG_MP_NBT", nbt)   # propagate kernel: rank lookups / static neighbour table (halo fids in the table)
    lx, ly, lz = shape[:3]
    nat = random_nature(lx, ly, lz, 0.25, 77)
    if len(shape) > 3:      # solid walls at both z ends: the halo planes across the ring seam hold no fluid node
        nat[0] = 1
        nat[-1] = 1
    f = [1e-4, -2e-4, 3e-4]
    Db, ka, kd = 0.01, 0.1, 0.01
    with lb.LaboetieGPU(nat) as one:
        one.lb_init(1.0)
        d1, c1, h1 = one.lb_step(4, tau=tau, check_every=1, target_error=-1.0)
        one.lb_set_force_uniform(f)
        d2, c2, h2 = one.lb_step(11, tau=tau, check_every=1, target_error=-1.0)
        n_ref = one.lb_populations()
        mom_ref = one.lb_moments()
        prof_ref = [one.lb_profiles(a) for a in range(3)]
        one.lb_step(2, tau=tau, check_every=0)
        v0_ref = one.mp_init(Db, ka, kd, f)
        dm, cm, v_ref = one.mp_step(9)
        P_ref, A_ref = one.mp_download()
        itf_ref = None
    uid = api.comm_unique_id()

    def rank_fn(r):
        sim = slab.make_slab_sim(nat, r, nranks, device=r, unique_id=uid)
        try:
            if in_place:    # AA pattern across the slabs: forward exchange after the local step, masked return trip
                sim.lb_set_in_place(True)
            sim.lb_init(1.0)
            _, _, g1 = sim.lb_step(4, tau=tau, check_every=1, target_error=-1.0)
            sim.lb_set_force_uniform(f)
            _, _, g2 = sim.lb_step(11, tau=tau, check_every=1, target_error=-1.0)
            n = sim.lb_populations()
            mom = sim.lb_moments()
            prof_z = sim.lb_profiles(2)
            prof_xy = [sim.lb_profiles(a, raw=True) for a in (0, 1)]
            sim.lb_step(2, tau=tau, check_every=0)
            if in_place:
                assert sim.info("in_place") == 1
            v0 = sim.mp_init(Db, ka, kd, f)
            _, _, v = sim.mp_step(9)
            P, A = sim.mp_download()
            return dict(k0=sim.k0, nzl=sim.nzl, g1=g1, g2=g2, n=n, mom=mom, prof_z=prof_z, prof_xy=prof_xy, v0=v0, v=v,
                        P=P, A=A, counts=sim.counts())
        finally:
            sim.close()

    res = _run_ranks(nranks, rank_fn)
    for r in res:
        assert np.array_equal(r["g1"], h1) and np.array_equal(r["g2"], h2)      # global l2err on every rank
        sl = slice(r["k0"], r["k0"] + r["nzl"])
        assert np.array_equal(r["n"], n_ref[:, sl])
        for a, b in zip(r["mom"], mom_ref):
            assert np.array_equal(a, b[sl])
        assert np.array_equal(r["P"], P_ref[sl]) and np.array_equal(r["A"], A_ref[sl])
        assert np.allclose(r["prof_z"], prof_ref[2][sl], rtol=RTOL, atol=0)
        assert np.allclose(r["v0"], v0_ref, rtol=RTOL, atol=0)
        assert (np.abs(r["v"] - v_ref) <= RTOL * np.abs(v0_ref).max()).all()
    # x / y profiles: partial sums and counts add up over the slabs
    for ai, axis in enumerate((0, 1)):
        tot = sum(r["prof_xy"][ai] for r in res)
        got = np.concatenate([tot[:, :3], (tot[:, 3] / np.maximum(tot[:, 4], 1))[:, None]], axis=1)
        assert np.allclose(got, prof_ref[axis], rtol=RTOL, atol=1e-300)
    assert sum(r["counts"][0] for r in res) == int((nat == 0).sum())


def test_slab_equilibration_exit_step():
    """The convergence criterion uses the global max: every rank leaves the loop at the single-GPU step."""
    import laboetie_b200 as lb
    from laboetie_b200 import api, driver, slab
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import oracle as O
    nat = O.geometry(1, 3, 2, 14)
    with lb.LaboetieGPU(nat) as one:
        ref = driver.equilibration(one, [1e-5, 0, 0], chunk=113)
        jx_ref = one.lb_moments()[1]
    uid = api.comm_unique_id()

    def rank_fn(r):
        sim = slab.make_slab_sim(nat, r, 2, device=r, unique_id=uid)
        try:
            out = driver.equilibration(sim, [1e-5, 0, 0], chunk=113)
            out["jx"] = sim.lb_moments()[1]
            out["k0"], out["nzl"] = sim.k0, sim.nzl
            return out
        finally:
            sim.close()

    for r in _run_ranks(2, rank_fn):
        assert (r["t_exit"], r["t_fext"]) == (ref["t_exit"], ref["t_fext"])
        assert np.array_equal(r["l2err"], ref["l2err"])
        assert np.array_equal(r["jx"], jx_ref[r["k0"]:r["k0"] + r["nzl"]])


def test_dead_neighbour_is_an_error_not_a_hang(monkeypatch):
    """A rank whose neighbour never sends its halo planes gets LBG_ERR_NCCL after the bounded spin of the wait
    kernels (LBG_P2P_TIMEOUT_S), instead of hanging the GPU."""
    import time
    import laboetie_b200 as lb
    from laboetie_b200 import api, slab
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("LBG_P2P_TIMEOUT_S", "1")
    nat = random_nature(8, 6, 10, 0.2, 5)
    uid = api.comm_unique_id()
    gate = threading.Event()

    def rank_fn(r):
        sim = slab.make_slab_sim(nat, r, 2, device=r, unique_id=uid)
        try:
            sim.lb_init(1.0)          # collective: both ranks get here
            if r == 1:
                gate.wait(timeout=60)   # this rank "dies": it never steps
                return "idle"
            t0 = time.time()
            try:
                sim.lb_step(3, tau=1.0, check_every=1, target_error=-1.0)
            except lb.LbgError as e:
                return ("error", e.status, time.time() - t0)
            finally:
                gate.set()
            return ("no error", 0, time.time() - t0)
        finally:
            sim.close()

    res = _run_ranks(2, rank_fn)
    kind, status, dt = res[0]
    assert kind == "error" and status == 12 and dt < 30, res[0]


def test_destroy_with_an_exchange_in_flight():
    """lbg_mp_init ends with the first exchange of P; a driver that destroys its handle right away must not free the
    receive buffer under the neighbour's push (lbg_destroy drains the pending exchange first)."""
    from laboetie_b200 import api, slab
    if _ndev() < 2:
        pytest.skip("needs 2 GPUs")
    nat = random_nature(12, 7, 10, 0.25, 9)
    for _ in range(3):
        uid = api.comm_unique_id()

        def rank_fn(r):
            sim = slab.make_slab_sim(nat, r, 2, device=r, unique_id=uid)
            try:
                sim.lb_init(1.0)
                sim.lb_step(3, tau=1.0, check_every=1, target_error=-1.0)
                return sim.mp_init(0.01, 0.1, 0.01, [1e-4, 0, 0])
            finally:
                sim.close()

        a, b = _run_ranks(2, rank_fn)
        assert np.array_equal(a, b)
