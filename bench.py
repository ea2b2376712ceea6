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
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
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
def cpu_run(workload, seconds, steps=None, threads=None):
    """The reference-shaped OpenMP restatement (oracle/) on a bounded crop of the same workload."""
    from oracle import oracle as O
    from laboetie_b200 import synthetic as S
    builder, lx, ly, lz_per, f_ext, desc = S.WORKLOADS[workload]
    cx, cy, cz = min(lx, 256), min(ly, 256), min(lz_per, 32)
    full_kw = {}
    nat = builder(lx, ly, lz_per, k0=0, nz=cz)[:, :cy, :cx].copy() if workload != "cfg3" else builder(64, 64, 64)[:32]
    nat = np.ascontiguousarray(nat)
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
    st.step()  # warm-up (page faults)
    t0 = time.perf_counter()
    k = 0
    while True:
        st.step()
        k += 1
        if (steps and k >= steps) or (not steps and time.perf_counter() - t0 > seconds / 2):
            break
    t_lb = (time.perf_counter() - t0) / k
    mp = O.MPState(nat, itf, st.rho, st.jx, st.jy, st.jz, f_ext, TRACER["Db"], TRACER["ka"], TRACER["kd"])
    mp.propagate()
    t0 = time.perf_counter()
    for _ in range(k):
        mp.propagate()
    t_mp = (time.perf_counter() - t0) / k
    mlups = n / (t_lb + t_mp) / 1e6
    return dict(value=mlups, unit="MLUPS", cores=threads or ncores, kind="port",
                sample=f"{cx}x{cy}x{nat.shape[0]} crop of {workload}, {k} LB + {k} MP steps, "
                       f"oracle/ C++/OpenMP restatement (no Fortran compiler: reference binary unavailable)",
                lb_mlups=n / t_lb / 1e6, mp_mlups=n / t_mp / 1e6, steps=k), (t_lb + t_mp) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = max(args.warmup, 0)
    res, ms = cpu_run(args.workload, args.cpu_seconds, steps=max(1, min(args.steps, 4)))
    # the reference's README recommends 4 OpenMP threads (README.md:94): reported beside the all-cores figure
    res4, _ = cpu_run(args.workload, args.cpu_seconds, steps=max(1, min(args.steps, 2)), threads=4)
    line = {"metric": "MLUPS (fp64 D3Q19 collide-stream + moment propagation)", "value": res["value"], "unit": "MLUPS",
            "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "timed_steps": res["steps"], "warmup": warm,
            "ms_per_step": ms, "cpu_4_threads": {"value": res4["value"], "unit": "MLUPS", "cores": res4["cores"]},
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "tau": TAU, **TRACER},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "lb_mlups": res["lb_mlups"], "mp_mlups": res["mp_mlups"]}
    print(json.dumps(line))


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
        return sim

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
    sim.lb_step(K, tau=TAU, check_every=ce, target_error=-1.0, want_history=False)
    t_lb = sim.timer_stop()
    barrier()
    l_lb = sim.launches - l0
    sim.mp_init(TRACER["Db"], TRACER["ka"], TRACER["kd"], f_ext)
    sim.mp_step(W, want_history=False)
    l0 = sim.launches
    barrier()
    sim.sync()
    sim.timer_start()
    sim.mp_step(K, want_history=False)
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
        done, conv, hist = sim.lb_step(K, tau=TAU, check_every=1, target_error=-1.0)
        mark("lb_steps")          # K steps, l2err history D2H
        if bufs is not None:
            sim._ck(sim._L.lbg_lb_download_moments(sim._h, *bufs))
        else:
            bufs = sim.lb_moments()
        mark("moments_d2h")       # density and momentum density into pinned host arrays
        v0 = sim.mp_init(TRACER["Db"], TRACER["ka"], TRACER["kd"], f_ext)
        mark("mp_init")
        done, conv, vac = sim.mp_step(K)
        sim.sync()
        mark("mp_steps")          # K steps, vacf rows D2H
        t_e2e = allmax(time.perf_counter() - t0)
        sim.close()
        phases = {b[0]: b[1] - a[1] for a, b in zip(marks, marks[1:])}
        own = nat[1:-1].size if nranks > 1 else nat.size
        e2e = {"value": n_total * K / t_e2e / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": float(nat.nbytes * nranks) / K,
               "d2h_bytes_per_step": float((4 * 8 * own) * nranks) / K + 8 + 24,
               "seconds": t_e2e, "setup_seconds": t_setup, "phase_seconds": phases,
               "what": "create(H2D geometry)+lb_init+K LB steps(l2err history D2H)+moments D2H(pinned)+mp_init+K MP steps(vacf D2H); fixed costs (setup_seconds, the moments read-back, mp_init) amortised over K"}

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
        "config": {"workload": args.workload, "description": desc, "lattice": [lx, ly, lz], "parallelism": f"z-slabs x{nranks}",
                   "fluid_fraction": nf_tot / n_total, "interfacial_fluid_fraction": nif_tot / n_total, "tau": TAU, **TRACER,
                   "check_every": ce, "phase_a_layout": "in-place (AA)" if args.in_place else "two-lattice",
                   "l2": "working set >> 126 MB L2; no flush needed",
                   "step": "1 LB step + 1 MP step; K LB steps then K MP steps timed"},
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
    if also:
        line["also"] = also
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline:
        res, _ = cpu_run(args.workload, args.cpu_seconds)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    # optional ncu-derived DRAM traffic for the dominant kernel (profiles/traffic_<workload>.json)
    tp = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(tp):
        try:
            tr = json.load(open(tp))
            line["roofline"]["traffic"] = tr.get("lb_step_kernel_bytes_per_launch")
            line["roofline"]["mp_step_kernel"]["traffic"] = tr.get("mp_step_kernel_bytes_per_launch")
        except Exception:
            pass
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
