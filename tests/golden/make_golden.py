"""Writes the oracle-output fixtures under tests/golden/ (run from the repo root).

These are outputs of the CPU oracle, not of the reference binary (which cannot be
built here: no Fortran compiler) -- see tests/golden/README.md.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tuto_cfg1(mp_steps=400):
    """BASELINE config 1: tuto one-disk 1x50x50, f_ext=(0,1e-5,0), Db=0.01, ka=0.1, kd=0.01 (README example values)."""
    nat = O.read_geom_in(os.path.join(HERE, "geom.in_chromat_1disks-dia10-1x50x50_v1"), 1, 50, 50)
    itf = O.detect_interfacial(nat)
    f = [0.0, 1e-5, 0.0]
    r = O.equilibration(nat, f, tau=1.0, target_error=1e-10)
    assert r["rc"] == 0
    mp = O.MPState(nat, itf, r["rho"], r["jx"], r["jy"], r["jz"], f, 0.01, 0.1, 0.01)
    vacf = [mp.vacf0.copy()]
    for _ in range(mp_steps):
        rc, v, conv = mp.propagate()
        assert rc == 0 and not conv
        vacf.append(v)
    np.savez_compressed(os.path.join(HERE, "tuto_cfg1_oracle.npz"), t_exit=r["t_exit"], t_fext=r["t_fext"],
                        l2err=r["l2err"], rho=r["rho"], jx=r["jx"], jy=r["jy"], jz=r["jz"],
                        prof_z=O.profiles(r["rho"], r["jx"], r["jy"], r["jz"], 2), mp_steps=mp_steps,
                        P=mp.P[0], Pads=mp.Pads[0], vacf=np.array(vacf))
    print("tuto_cfg1: t_exit", r["t_exit"], "t_fext", r["t_fext"], "vacf[-1]", vacf[-1])


if __name__ == "__main__":
    tuto_cfg1()
