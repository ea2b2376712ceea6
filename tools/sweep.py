"""A/B runs of environment-selected kernel variants in ONE process (one box, one geometry build).

  python tools/sweep.py cfg5w 20 "LBG_LB_TPC=0 LBG_MP_TPC=0" "LBG_LB_TPC=1 LBG_MP_TPC=4" ...

Each quoted argument is a set of NAME=VALUE pairs put into the environment before the handle is created (the
library reads its tuning variables in lbg_create*).  Prints ms per LB / MP step (CUDA events) and the fraction of
the HBM roofline by the SURVEY 8d byte formulas.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laboetie_b200 as lb  # noqa: E402
from laboetie_b200 import synthetic as S  # noqa: E402

wl = sys.argv[1]
K = int(sys.argv[2])
configs = sys.argv[3:] or [""]
builder, lx, ly, lz, f_ext, desc = S.WORKLOADS[wl]
nat = builder(lx, ly, lz)
try:
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    hbm = 6650.0
ce = int(os.environ.get("SWEEP_CHECK_EVERY", "1"))
for cfg in configs:
    saved = dict(os.environ)
    for kv in cfg.split():
        k, v = kv.split("=")
        os.environ[k] = v
    with lb.LaboetieGPU(nat) as sim:
        nf, nif = sim.counts()
        sim.lb_init(1.0)
        sim.lb_set_force_uniform(f_ext)
        sim.lb_step(3, tau=1.0, check_every=ce, target_error=-1.0, want_history=False)
        sim.sync(); sim.timer_start()
        sim.lb_step(K, tau=1.0, check_every=ce, target_error=-1.0, want_history=False)
        tl = sim.timer_stop() / K
        sim.mp_init(0.01, 0.1, 0.01, f_ext)
        sim.mp_step(3, want_history=False)
        sim.sync(); sim.timer_start()
        d, c, v = sim.mp_step(K)
        tm = sim.timer_stop() / K
    bl = (352 if ce == 1 else 304) * nf + nat.size
    bm = 208 * nf + 48 * nif + nat.size
    print(f"{wl} [{cfg}] lb {tl:.3f} ms frac {bl / tl * 1e-6 / hbm:.3f} | mp {tm:.3f} ms frac {bm / tm * 1e-6 / hbm:.3f} "
          f"| value {nat.size / (tl + tm) * 1e-3:.0f} MLUPS | vacf[-1] {v[-1].tolist()}", flush=True)
    os.environ.clear()
    os.environ.update(saved)
