"""Short run for ncu: a few LB steps and MP steps on one workload (no timing, no CPU work).

  ncu --set full -k regex:lb_step_kernel -s 3 -c 2 -o gpurun_out/prof_lb python tools/profile_run.py cfg5w
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import laboetie_b200 as lb  # noqa: E402
from laboetie_b200 import synthetic as S  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5w"
nlb = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nmp = int(sys.argv[3]) if len(sys.argv) > 3 else 6
ce = int(sys.argv[4]) if len(sys.argv) > 4 else 1
builder, lx, ly, lz, f_ext, desc = S.WORKLOADS[wl]
nat = builder(lx, ly, lz)
with lb.LaboetieGPU(nat) as sim:
    print(wl, desc, "fluid, interfacial fluid:", sim.counts(), "of", nat.size, flush=True)
    sim.lb_init(1.0)
    sim.lb_set_force_uniform(f_ext)
    sim.lb_step(nlb, tau=1.0, check_every=ce, target_error=-1.0, want_history=False)
    sim.mp_init(0.01, 0.1, 0.01, f_ext)
    sim.mp_step(nmp, want_history=False)
    print("launches", sim.launches)
