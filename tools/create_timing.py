"""Repeated create / step / destroy of the benchmark lattice with LBG_TIMING=1: where does set-up time go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["LBG_TIMING"] = "1"
import numpy as np
import laboetie_b200 as lb
from laboetie_b200 import synthetic as S
nat = S.porous_spheres(1024, 1024, 128)
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
for it in range(4):
    if it == 2 and mode == "torch":
        import torch
        bufs = [torch.empty(nat.shape, dtype=torch.float64, pin_memory=True).numpy() for _ in range(4)]
        print("pinned 4 GB", flush=True)
    t0 = time.perf_counter()
    sim = lb.LaboetieGPU(nat)
    t1 = time.perf_counter()
    sim.lb_init(1.0); sim.lb_step(3, want_history=False); sim.sync()
    t2 = time.perf_counter()
    sim.close()
    t3 = time.perf_counter()
    print(f"iter {it}: create {1e3*(t1-t0):.1f} ms  init+3 steps {1e3*(t2-t1):.1f} ms  close {1e3*(t3-t2):.1f} ms", flush=True)
