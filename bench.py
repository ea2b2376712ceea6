#!/usr/bin/env python
"""bench.py -- MLUPS of the laboetie hot path (fp64 D3Q19 collide-stream + moment propagation).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

One bench "step" = one Lattice-Boltzmann step (collide + bounce-back + stream +
moments + max|dj|) plus one moment-propagation step over the whole lattice.
The timed region runs K LB steps then K MP steps (the reference's two phases are
sequential: equilibration.f90 then drop_tracers.f90) with the lattice resident in
HBM, timed with CUDA events on the library's own stream, max over ranks.
value = N_nodes * K / (t_LB + t_MP) / 1e6  [MLUPS, whole job over all GPUs].

`e2e` is the same quantity through the C ABI from host buffers: create (H2D of
the geometry), lb_init, K LB steps with the per-step l2err history read back,
the density/momentum read-back the reference does at equilibration.f90:551-554
(D2H, pinned), mp_init, K MP steps with the vacf history read back -- host
wall-clock around the calls.

Workload (config.workload): "cfg5w" = BASELINE config 5, weak scaling: synthetic
random porous medium 1024x1024x(128 per GPU), the configuration the metric
("... at 1/2/4/8 B200") is quoted on; "cfg2" (64x64x256 slit) and "cfg3" (256^3
BCC) are available with --workload.  All are far larger than L2 (126 MB), so no
L2 flush is needed between iterations (config.l2 says so).

torch is used for plumbing only (torch.distributed rendezvous/barrier, pinned
host buffers); the kernels are this repo's own (laboetie_b200/lib/liblaboetie_gpu.so).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRACER = dict(Db=0.01, ka=0.1, kd=0.01)   # README example values (README.md:129-133)
if os.environ.get("LBG_BENCH_KA"):    # tuning runs only (e.g. 0 switches adsorption off)
    TRACER["ka"] = float(os.environ["LBG_BENCH_KA"])
TAU = 1.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5w")
    ap.add_argument("--check-every", type=int, default=1, help="1 = reference semantics (l2err every step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-reps", type=int, default=3, help="complete end-to-end passes; the faster one is reported")
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="seconds of CPU stepping the sample is sized for (default: 20 for cpu_baseline, 150 for --impl reference)")
    ap.add_argument("--cpu-lattice", default="sample", choices=["sample", "full"],
                    help="--impl reference: a slab sample sized by --cpu-seconds, or the workload's whole single-GPU lattice "
                         "(needs ~60 GB of host memory and ~43 s per step for cfg5w)")
    ap.add_argument("--no-verify", action="store_true", help="skip the parity check of the launched configuration")
    ap.add_argument("--in-place", action="store_true", help="Phase A with the AA pattern (one population buffer)")
    ap.add_argument("--also", default="cfg2,cfg3", help="extra single-GPU workloads reported under 'also' (N=1 only)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        try:
            self.proc.wait(timeout=5)   # nvidia-smi holds driver-wide locks while it runs and while it goes away
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        load = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_geometry(workload, rank, nranks):
    from laboetie_b200 import synthetic as S
    from laboetie_b200 import api
    builder, lx, ly, lz_per, f_ext, desc = S.WORKLOADS[workload]
    weak = workload in S.WEAK
    lz = lz_per * nranks if weak else lz_per
    k0, nzl = api.partition(lz, nranks, rank)
    if nranks == 1:
        nat = builder(lx, ly, lz)
    else:
        nat = builder(lx, ly, lz, k0=k0 - 1, nz=nzl + 2)
    return nat, (lx, ly, lz), (k0, nzl), f_ext, desc, ("weak" if weak else "strong")


def algorithmic_bytes(nf, nif, n, check):
    """SURVEY 8d / BASELINE.md 3: LB 304 N_f + N (+48 N_f with the per-step max|dj|); MP 208 N_f + 48 N_if + N."""
    lb = (352 if check else 304) * nf + n
    mp = 208 * nf + 48 * nif + n
    return lb, mp


# --------------------------------------------------------------------------- CPU arm
def host_mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 2 ** 20
    except OSError:
        pass
    return 0.0


def workload_config(args, nranks):
    """`config` of the JSON line: what is computed.  Both arms print the same dict for the same command line (the CPU
    arm runs it on a bounded sample, `cpu_baseline.sample` says which)."""
    from laboetie_b200 import synthetic as S
    builder, lx, ly, lz_per, f_ext, desc = S.WORKLOADS[args.workload]
    lz = lz_per * nranks if args.workload in S.WEAK else lz_per
    return {"workload": args.workload, "description": desc, "lattice": [lx, ly, lz], "parallelism": f"z-slabs x{nranks}",
            "tau": TAU, **TRACER, "check_every": args.check_every,
            "phase_a_layout": "in-place (AA)" if args.in_place else "two-lattice",
            "l2": "working set >> 126 MB L2; no flush needed",
            "step": "1 LB step + 1 MP step; K LB steps then K MP steps timed"}


def cpu_run(workload, nplanes, steps, warmup=1, threads=None):
    """The reference-shaped OpenMP restatement (oracle/) on a slab sample of the workload: the lattice's whole
    x-y cross-section, `nplanes` z-planes (all of them: the whole single-GPU lattice), periodic like the lattice.
    `warmup` untimed and `steps` timed steps of each phase."""
    from oracle import oracle as O
    from laboetie_b200 import synthetic as S
    builder, lx, ly, lz_per, f_ext, desc = S.WORKLOADS[workload]
    nz = max(1, min(int(nplanes), lz_per))
    if builder in (S.bcc,):                # the BCC cell is only defined as a cube: sample the first planes of it
        nat = np.ascontiguousarray(builder(lx, ly, lz_per)[:nz])
    else:
        nat = np.ascontiguousarray(builder(lx, ly, lz_per, k0=0, nz=nz))
    if nat.all():
        nat.flat[0] = 0
    ncores = os.cpu_count() or 1
    try:      # the oracle's OpenMP regions follow omp_set_num_threads of the libgomp it is linked to
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads or ncores))
    except OSError:
        threads = None
    itf = O.detect_interfacial(nat)
    st = O.LBState(nat, 1.0, TAU)
    st.set_force_uniform(f_ext)
    n = nat.size
    for _ in range(warmup):
        st.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step()
    t_lb = (time.perf_counter() - t0) / steps
    rho, jx, jy, jz = st.rho, st.jx, st.jy, st.jz
    del st          # the reference frees its populations before Phase B too (drop_tracers.f90:85)
    mp = O.MPState(nat, itf, rho, jx, jy, jz, f_ext, TRACER["Db"], TRACER["ka"], TRACER["kd"])
    for _ in range(warmup):
        mp.propagate()
    t0 = time.perf_counter()
    for _ in range(steps):
        mp.propagate()
    t_mp = (time.perf_counter() - t0) / steps
    mlups = n / (t_lb + t_mp) / 1e6
    whole = nz == lz_per
    return dict(value=mlups, unit="MLUPS", cores=threads or ncores, kind="port",
                sample=f"{lx}x{ly}x{nz} slab of {workload} ({'the whole single-GPU lattice' if whole else 'full x-y cross-section, ' + str(nz) + ' of ' + str(lz_per) + ' z-planes'}), "
                       f"{warmup} + {steps} LB and {warmup} + {steps} MP steps, oracle/ C++/OpenMP restatement "
                       f"(no Fortran compiler: reference binary unavailable)",
                lb_mlups=n / t_lb / 1e6, mp_mlups=n / t_mp / 1e6, steps=steps, lattice=[lx, ly, nz],
                whole=bool(whole)), (t_lb + t_mp) * 1e3


def cpu_sample_planes(workload, budget_s, nsteps_total):
    """How many z-planes of the workload the CPU arm can step `nsteps_total` times (LB + MP) in about `budget_s`
    seconds on this host: a one-step probe on a thin slab gives the rate."""
    from laboetie_b200 import synthetic as S
    builder, lx, ly, lz_per, f_ext, desc = S.WORKLOADS[workload]
    plane = lx * ly
    nz0 = max(2, min(lz_per, (4 << 20) // plane))
    probe, _ = cpu_run(workload, nz0, steps=1, warmup=0)
    rate = probe["value"] * 1e6                       # nodes per second, LB + MP
    nz = int(budget_s * rate / (max(nsteps_total, 1) * plane))
    return max(min(4, lz_per), min(nz, lz_per)), probe["value"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from laboetie_b200 import synthetic as S
    builder, lx, ly, lz_per, f_ext, desc = S.WORKLOADS[args.workload]
    K, W = max(args.steps, 1), max(args.warmup, 0)
    # K timed and W warm-up steps, as asked, each on a bounded sample: the whole x-y cross-section and as many
    # z-planes as fit --cpu-seconds of stepping (the whole single-GPU lattice with --cpu-lattice full: ~43 s per step
    # for cfg5w on 16 cores and ~60 GB of host memory, same MLUPS -- profiles/bench_ref_r5l.json)
    if args.cpu_lattice == "full":
        nz, probe = lz_per, None
    else:
        nz, probe = cpu_sample_planes(args.workload, args.cpu_seconds, K + W)
    res, ms = cpu_run(args.workload, nz, steps=K, warmup=W)
    # the reference's README recommends 4 OpenMP threads (README.md:94): reported beside the all-cores figure
    res4, _ = cpu_run(args.workload, max(2, min(lz_per, (4 << 20) // (lx * ly))), steps=1, warmup=0, threads=4)
    line = {"metric": "MLUPS (fp64 D3Q19 collide-stream + moment propagation)", "value": res["value"], "unit": "MLUPS",
            "impl": "reference", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": ms,
            "cpu_4_threads": {"value": res4["value"], "unit": "MLUPS", "cores": res4["cores"], "sample": res4["sample"]},
            "higher_is_better": True, "scaling": "weak" if args.workload in S.WEAK else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, args.gpus),
            "sample_lattice": res["lattice"], "sample_is_whole_single_gpu_lattice": res["whole"],
            "ms_per_step_is_for": "one LB + one MP step of the sample lattice",
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "lb_mlups": res["lb_mlups"], "mp_mlups": res["mp_mlups"]}
    print(json.dumps(line))


# --------------------------------------------------------------------------- parity of the launched configuration
def kernel_source_hash():
    """sha1 over the kernel sources: ncu-derived traffic figures are only valid for the code they were taken on."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "laboetie_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def verify_launch(dist, rank, nranks, local, in_place=False):
    """Run a small porous lattice through exactly the path this launch uses -- one process per GPU, z-slabs, halos
    pushed through CUDA-IPC-mapped peer memory, scalar all-reduces through the peers' mailboxes -- and compare
    every per-node result with the CPU oracle on rank 0, bit for bit (the oracle is the checker here, nothing of
    it is timed).  12 LB steps with the per-step check across the force switch, 8 propagate steps with adsorption."""
    import laboetie_b200 as lb
    from laboetie_b200 import api, synthetic as S
    lx, ly, lz = 96, 40, 6 * nranks
    f_ext, tau = [1e-5, 0.0, 2e-5], 0.9
    k0, nzl = api.partition(lz, nranks, rank)
    if nranks == 1:
        sim = lb.LaboetieGPU(S.porous_spheres(lx, ly, lz, radius=5), device=local)
    else:
        sim = lb.LaboetieGPU(S.porous_spheres(lx, ly, lz, radius=5, k0=k0 - 1, nz=nzl + 2), device=local, lz_global=lz,
                             k0=k0, slab=True)
        uid = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.comm_init(nranks, rank, uid[0])
    if in_place:
        sim.lb_set_in_place(True)
    path = "single GPU" if nranks == 1 else ("ipc" if sim.info("ipc") else ("peer" if sim.info("p2p") else "nccl"))
    sim.lb_init(1.0)
    _, _, h1 = sim.lb_step(4, tau=tau, check_every=1, target_error=-1.0)
    sim.lb_set_force_uniform(f_ext)
    _, _, h2 = sim.lb_step(8, tau=tau, check_every=1, target_error=-1.0)
    mine = dict(k0=k0, nzl=nzl, hist=np.concatenate([h1, h2]), n=sim.lb_populations(), mom=sim.lb_moments())
    mine["v0"] = sim.mp_init(TRACER["Db"], TRACER["ka"], TRACER["kd"], f_ext)
    done, _, mine["v"] = sim.mp_step(8)
    mine["P"], mine["A"] = sim.mp_download()
    sim.close()
    parts = [mine]
    if dist is not None:
        parts = [None] * nranks if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    from oracle import oracle as O
    nat = S.porous_spheres(lx, ly, lz, radius=5)
    st = O.LBState(nat, 1.0, tau)
    ref_h = [st.step()[1] for _ in range(4)]
    st.set_force_uniform(f_ext)
    ref_h += [st.step()[1] for _ in range(8)]
    mp = O.MPState(nat, O.detect_interfacial(nat), st.rho, st.jx, st.jy, st.jz, f_ext, TRACER["Db"], TRACER["ka"], TRACER["kd"])
    ref_v = np.array([mp.propagate()[1] for _ in range(8)])
    worst, ok = 0.0, True
    for p in parts:
        sl = slice(p["k0"], p["k0"] + p["nzl"])
        pairs = [(p["n"], st.n[:, sl]), (p["P"], mp.P[0][sl]), (p["A"], mp.Pads[0][sl]), (p["hist"], np.array(ref_h))]
        pairs += [(a, b[sl]) for a, b in zip(p["mom"], (st.rho, st.jx, st.jy, st.jz))]
        for a, b in pairs:
            ok = ok and np.array_equal(a, b)
            worst = max(worst, float(np.abs(a - b).max()))
        scale = np.abs(mp.vacf0).max()      # cross-node sums: to summation order (north_star: 1e-12 relative)
        sum_err = max(float(np.abs(p["v"] - ref_v).max()), float(np.abs(p["v0"] - mp.vacf0).max())) / scale
        ok = ok and sum_err <= 1e-12
    return {"ok": bool(ok), "phase_a_layout": "in-place (AA)" if in_place else "two-lattice", "max_abs_diff": worst, "vacf_rel_diff": sum_err, "path": path, "lattice": [lx, ly, lz],
            "ranks": nranks, "against": "CPU oracle on rank 0 (bit for bit per node; vacf to 1e-12 relative)",
            "compared": "populations, density, momentum, l2err history (12 LB steps), P, Pads, vacf (8 MP steps)"}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nranks = args.gpus
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import laboetie_b200 as lb
    from laboetie_b200 import api

    K, W, ce = args.steps, max(args.warmup, 3), args.check_every
    nat, (lx, ly, lz), (k0, nzl), f_ext, desc, scaling = build_geometry(args.workload, rank, nranks)
    n_total = lx * ly * lz

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def make_sim():
        if nranks == 1:
            sim = lb.LaboetieGPU(nat, device=local)
            if args.in_place:
                sim.lb_set_in_place(True)
            return sim
        sim = lb.LaboetieGPU(nat, device=local, lz_global=lz, k0=k0, slab=True)
        uid = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.comm_init(nranks, rank, uid[0])
        if args.in_place:
            sim.lb_set_in_place(True)
        return sim

    # ---- parity of this very launch (process per GPU, IPC halos) against the oracle --------------
    verify = None
    if not args.no_verify:
        try:
            verify = verify_launch(dist, rank, nranks, local, in_place=args.in_place)
        except Exception as e:  # noqa: BLE001  -- a failed check must show up in the line, not kill the measurement
            verify = {"ok": False, "error": f"{type(e).__name__}: {e}"}

    def must_run(res, k, what):
        """The timed calls must execute exactly k steps (a convergence stop would inflate the rate silently)."""
        done = res[0]
        if done != k or res[1]:
            raise RuntimeError(f"{what}: {done} of {k} steps executed (converged={res[1]}); the timing would be invalid")

    # ---- device-resident measurement --------------------------------------------------------
    sim = make_sim()
    nf, nif = sim.counts()
    nf_tot, nif_tot = allsum(nf), allsum(nif)
    sim.lb_init(1.0)
    sim.lb_set_force_uniform(f_ext)
    sim.lb_step(W, tau=TAU, check_every=ce, target_error=-1.0, want_history=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sim.launches
    barrier()
    sim.sync()
    sim.timer_start()
    must_run(sim.lb_step(K, tau=TAU, check_every=ce, target_error=-1.0, want_history=False), K, "timed LB steps")
    t_lb = sim.timer_stop()
    barrier()
    l_lb = sim.launches - l0
    sim.mp_init(TRACER["Db"], TRACER["ka"], TRACER["kd"], f_ext)
    sim.mp_step(W, want_history=False)
    l0 = sim.launches
    barrier()
    sim.sync()
    sim.timer_start()
    must_run(sim.mp_step(K, want_history=False), K, "timed MP steps")
    t_mp = sim.timer_stop()
    barrier()
    l_mp = sim.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    t_lb, t_mp = allmax(t_lb), allmax(t_mp)
    sim.close()

    # ---- end to end through the C ABI from host buffers --------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            import torch
            pin = lambda: torch.empty(nat[1:-1].shape if nranks > 1 else nat.shape, dtype=torch.float64, pin_memory=True).numpy()  # noqa: E731
            bufs = [pin() for _ in range(4)]
        except Exception:
            bufs = None
        # the polling nvidia-smi of the clock sampler has just exited and 31 GB were just released: let the driver
        # settle on a throw-away handle so that neither lands inside the timed region (measured: 25-50 ms vs up to
        # 370 ms for the same lbg_create, profiles/create_timing_r5n.txt)
        with lb.LaboetieGPU(np.zeros((4, 8, 32), np.int8), device=local):
            pass
        # Three repetitions, the fastest one is reported (both times are in `seconds_all`): one CUDA call of the set-up
        # occasionally stalls for 0.3-0.4 s on some boxes of the pool (lbg_create 25-50 ms vs 370-430 ms for identical
        # work, profiles/create_timing_r5n.txt), which says nothing about the path measured.
        best = None
        seconds_all = []
        for rep in range(max(1, args.e2e_reps)):
            barrier()
            t0 = time.perf_counter()
            marks = [("start", t0)]
            mark = lambda name: marks.append((name, time.perf_counter()))   # noqa: E731
            sim = make_sim()
            mark("create")            # H2D of the geometry, device allocations, numbering
            sim.lb_init(1.0)
            sim.lb_set_force_uniform(f_ext)
            sim.sync()
            mark("lb_init")
            t_setup = time.perf_counter() - t0
            res = sim.lb_step(K, tau=TAU, check_every=1, target_error=-1.0)
            must_run(res, K, "e2e LB steps")
            mark("lb_steps")          # K steps, l2err history D2H
            if bufs is not None:
                sim.lb_moments_async(bufs)   # the reference's write-back (equilibration.f90:551-554), overlapped with Phase B
            else:
                bufs = sim.lb_moments()
            mark("moments_d2h_queued")
            v0 = sim.mp_init(TRACER["Db"], TRACER["ka"], TRACER["kd"], f_ext)
            mark("mp_init")
            res = sim.mp_step(K)
            must_run(res, K, "e2e MP steps")
            sim.sync()
            mark("mp_steps")          # K steps, vacf rows D2H
            sim.wait_transfers()
            mark("moments_d2h_tail")  # what is left of the density / momentum read-back (pinned host arrays) after Phase B
            t_e2e = allmax(time.perf_counter() - t0)
            sim.close()
            seconds_all.append(t_e2e)
            if best is None or t_e2e < best[0]:
                best = (t_e2e, t_setup, {b_[0]: b_[1] - a_[1] for a_, b_ in zip(marks, marks[1:])})
        t_e2e, t_setup, phases = best
        own = nat[1:-1].size if nranks > 1 else nat.size
        e2e = {"value": n_total * K / t_e2e / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": float(nat.nbytes * nranks) / K,
               "d2h_bytes_per_step": float((4 * 8 * own) * nranks) / K + 8 + 24,
               "seconds": t_e2e, "seconds_all": seconds_all, "repetitions": len(seconds_all),
               "setup_seconds": t_setup, "phase_seconds": phases,
               "what": "create(H2D geometry)+lb_init+K LB steps(l2err history D2H)+moments D2H(pinned, queued before and awaited after Phase B)+mp_init+K MP steps(vacf D2H); "
                       "fixed costs (setup_seconds, the moments read-back, mp_init) amortised over K; the CUDA context and, "
                       "at N>1, the NCCL bootstrap communicator exist already (one per process, reused across handles); "
                       "the faster of `repetitions` complete passes"}

    # ---- the other BASELINE configurations that fit one GPU, device-resident numbers only (N=1) ------
    also = {}
    if nranks == 1 and args.also:
        hbm0, _ = peaks()
        for wl in [w for w in args.also.split(",") if w and w != args.workload]:
            try:
                nat2, (ax, ay, az), _, f2, desc2, _ = build_geometry(wl, 0, 1)
                with lb.LaboetieGPU(nat2, device=local) as s2:
                    nf2, nif2 = s2.counts()
                    s2.lb_init(1.0)
                    s2.lb_set_force_uniform(f2)
                    s2.lb_step(W, tau=TAU, check_every=ce, target_error=-1.0, want_history=False)
                    s2.sync(); s2.timer_start()
                    s2.lb_step(4 * K, tau=TAU, check_every=ce, target_error=-1.0, want_history=False)
                    tl = s2.timer_stop() / (4 * K)
                    s2.mp_init(TRACER["Db"], TRACER["ka"], TRACER["kd"], f2)
                    s2.mp_step(W, want_history=False)
                    s2.sync(); s2.timer_start()
                    s2.mp_step(4 * K, want_history=False)
                    tm = s2.timer_stop() / (4 * K)
                bl, bm = algorithmic_bytes(nf2, nif2, nat2.size, ce == 1)
                also[wl] = {"description": desc2, "lattice": [ax, ay, az], "fluid_fraction": nf2 / nat2.size,
                            "value": nat2.size / ((tl + tm) * 1e-3) / 1e6, "lb_mlups": nat2.size / (tl * 1e-3) / 1e6,
                            "mp_mlups": nat2.size / (tm * 1e-3) / 1e6, "lb_roofline_frac": bl / (tl * 1e-3) / 1e9 / hbm0,
                            "mp_roofline_frac": bm / (tm * 1e-3) / 1e9 / hbm0, "steps": 4 * K}
            except Exception as e:  # noqa: BLE001
                also[wl] = {"error": str(e)}

    if rank != 0:
        return
    hbm, peak_src = peaks()
    check = ce == 1
    b_lb, b_mp = algorithmic_bytes(nf_tot / nranks, nif_tot / nranks, n_total / nranks, check)
    lb_gbs = b_lb / (t_lb / K * 1e-3) / 1e9
    mp_gbs = b_mp / (t_mp / K * 1e-3) / 1e9
    value = n_total * K / ((t_lb + t_mp) * 1e-3) / 1e6
    line = {
        "metric": "MLUPS (fp64 D3Q19 collide-stream + moment propagation)", "value": value, "unit": "MLUPS",
        "n_gpus": nranks, "steps": K, "warmup": W, "ms_per_step": (t_lb + t_mp) / K, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, nranks),
        "lattice_stats": {"fluid_fraction": nf_tot / n_total, "interfacial_fluid_fraction": nif_tot / n_total},
        "lb": {"mlups": n_total * K / (t_lb * 1e-3) / 1e6, "mflups": nf_tot * K / (t_lb * 1e-3) / 1e6, "ms_per_step": t_lb / K},
        "mp": {"mlups": n_total * K / (t_mp * 1e-3) / 1e6, "mflups": nf_tot * K / (t_mp * 1e-3) / 1e6, "ms_per_step": t_mp / K},
        "roofline": {"bound": "hbm", "kernel": "lb_step_kernel (pull stream + moments + collide)", "achieved": lb_gbs,
                     "peak": hbm, "unit": "GB/s", "frac": lb_gbs / hbm, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": b_lb,
                     "mp_step_kernel": {"achieved": mp_gbs, "frac": mp_gbs / hbm, "algorithmic_bytes_per_launch": b_mp},
                     # what a lattice-free kernel of the same access shape reaches (tools/microbench/streams.cu)
                     "shape_yardstick": {"lb_pull_19_to_19_gbs": 5380.4, "mp_read_25_to_3_gbs": 6716.3,
                                         "source": "profiles/streams_r3i.txt"}},
        "gpu_launches": int(l_lb + l_mp), "clocks": clocks,
    }
    if verify is not None:
        line["verify"] = verify
    if also:
        line["also"] = also
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        nzc, _ = cpu_sample_planes(args.workload, args.cpu_seconds, 4 + 1)
        res, _ = cpu_run(args.workload, nzc, steps=4, warmup=1)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    # ncu-derived DRAM traffic per launch of the two step kernels (profiles/traffic_<workload>.json, written by
    # tools/ncu_traffic.py from one `ncu --set full` capture): reported only for N=1 and only if the file was
    # taken on exactly these kernel sources -- otherwise null (a stale figure would be worse than none)
    tp = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if nranks == 1 and os.path.exists(tp):
        try:
            tr = json.load(open(tp))
            if tr.get("kernel_source_hash") == kernel_source_hash():
                line["roofline"]["traffic"] = tr.get("lb_step_kernel_bytes_per_launch")
                line["roofline"]["mp_step_kernel"]["traffic"] = tr.get("mp_step_kernel_bytes_per_launch")
                line["roofline"]["traffic_source"] = tr.get("source")
        except Exception:
            pass
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.cpu_seconds is None:
        a.cpu_seconds = 150.0 if a.impl == "reference" else 20.0
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
