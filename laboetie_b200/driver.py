"""Python mirror of the reference's two phase drivers, on top of the C ABI.

The reference keeps lb.in parsing, geometry set-up, the equilibration state
machine and the output files in Fortran (equilibration.f90, drop_tracers.f90);
only the loop bodies move to the GPU.  This module restates that control flow so
that the parity tests and the benchmark read like the reference's own drivers.
The compiled mirror (same control flow, plus the lb.in / geom.in readers and the
output files) is the C++ program in laboetie_b200/driver/.
"""
import numpy as np

from .api import LaboetieGPU


def equilibration(sim: LaboetieGPU, f_ext, tau=1.0, target_error=1e-10, rho0=1.0, max_steps=10**9, chunk=4096,
                  check_every=1):
    """Phase A, equilibration.f90:143-491 (uniform-force branch :381-386).

    Steps until l2err <= target_error with t > 2 (:346), switches the external
    force on at the first convergence (:377-386) and leaves the loop at the
    second (:373-374).  Returns dict(t_exit, t_fext, l2err, rc).
    """
    sim.lb_init(rho0)                       # init_simu.f90:24-39
    without_fext = False
    t_fext = 0
    hist = []
    t = 0
    while t < max_steps:
        done, conv, h = sim.lb_step(min(chunk, max_steps - t), tau=tau, check_every=check_every,
                                    target_error=target_error)
        hist.append(h)
        t += done
        if not conv:
            continue
        if not without_fext:                # first convergence: enable the force (:377-386)
            without_fext = True
            t_fext = t + 1
            sim.lb_set_force_uniform(f_ext)
        else:                               # second convergence: stationary state found (:373-374)
            return dict(rc=0, t_exit=t, t_fext=t_fext, l2err=np.concatenate(hist))
    return dict(rc=2, t_exit=t, t_fext=t_fext, l2err=np.concatenate(hist) if hist else np.zeros(0))


def compensating_force_field(nature, f_ext, particle_diameter=1, particle_coordinates=None, geometry_label=0):
    """The `compensate_f_ext` force field of equilibration.f90:388-487 (host side, as in the reference's
    driver): f_ext spread over the lattice points of a spherical particle of odd diameter, compensated by a
    uniform background in the bulk cell (geometryLabel = -1).  nature: int8 (lz, ly, lx).
    Returns (fx, fy, fz, nodes_in_particle)."""
    nature = np.asarray(nature)
    lz, ly, lx = nature.shape
    pd = int(particle_diameter)
    if pd % 2 == 0:
        raise ValueError("particle diameter must be odd")                                   # :391-395
    if lx % 2 == 0 or ly % 2 == 0 or lz % 2 == 0:
        raise ValueError("when compensate_f_ext, there should be odd number of nodes in all directions")   # :397-402
    px, py, pz = particle_coordinates or (lx // 2 + 1, ly // 2 + 1, lz // 2 + 1)            # :414 (1-based)
    pdr = pd // 2
    fluid = nature == 0
    inside = np.zeros(nature.shape, bool)
    for i in range(px - pdr, px + pdr + 1):                                                   # :419-432
        for j in range(py - pdr, py + pdr + 1):
            for k in range(pz - pdr, pz + pdr + 1):
                if (i - px) ** 2 + (j - py) ** 2 + (k - pz) ** 2 > (pd / 2.0) ** 2:
                    continue
                if not fluid[k - 1, j - 1, i - 1]:
                    raise ValueError("Dominika's particle at a solid node")                  # :435-438
                inside[k - 1, j - 1, i - 1] = True
    l = int(inside.sum())
    nfl = int(fluid.sum())
    out = []
    for d in range(3):
        f = np.where(inside, float(f_ext[d]), 0.0)
        # the reference recognises particle nodes by comparing all three components with f_ext (:450,468)
        out.append(f)
    match = (out[0] == f_ext[0]) & (out[1] == f_ext[1]) & (out[2] == f_ext[2])
    res = []
    for d in range(3):
        if geometry_label == -1:                                                              # :449-458
            f = np.where(match, -float(f_ext[d]) / nfl + out[d] / l, -float(f_ext[d]) / nfl)
        else:                                                                                 # :467-477
            f = np.where(match, out[d] / l, 0.0)
        res.append(np.where(fluid, f, 0.0))                                                   # :480-484
    return res[0], res[1], res[2], l


def equilibration_compensated(sim: LaboetieGPU, nature, f_ext, tau=1.0, target_error=1e-10, rho0=1.0,
                              particle_diameter=1, particle_coordinates=None, geometry_label=0, max_steps=10**9):
    """Phase A with compensate_f_ext = T (equilibration.f90:185-188,388-487): as `equilibration`, but the
    force switched on at the first convergence is the particle + background field, and the momentum at the
    particle centre is recorded at the start of every later step (output/v_centralnode.dat, :187)."""
    lz, ly, lx = np.asarray(nature).shape
    px, py, pz = particle_coordinates or (lx // 2 + 1, ly // 2 + 1, lz // 2 + 1)
    sim.lb_init(rho0)
    hist, probe = [], []
    without_fext, t, t_fext = False, 0, 0
    while t < max_steps:
        if without_fext:
            probe.append((t + 1 - t_fext,) + tuple(sim.lb_probe(px - 1, py - 1, pz - 1)[:3]))
            n = 1                                # v_centralnode.dat wants the probe before every step
        else:
            n = min(4096, max_steps - t)
        done, conv, h = sim.lb_step(n, tau=tau, check_every=1, target_error=target_error)
        hist.append(h)
        t += done
        if not conv:
            continue
        if not without_fext:
            without_fext = True
            t_fext = t + 1
            fx, fy, fz, _ = compensating_force_field(nature, f_ext, particle_diameter, particle_coordinates,
                                                     geometry_label)
            sim.lb_set_force_field(fx, fy, fz)
        else:
            return dict(rc=0, t_exit=t, t_fext=t_fext, l2err=np.concatenate(hist), v_centralnode=np.array(probe))
    return dict(rc=2, t_exit=t, t_fext=t_fext, l2err=np.concatenate(hist), v_centralnode=np.array(probe))


def drop_tracers(sim: LaboetieGPU, f_ext, Db, ka, kd, max_steps, chunk=4096):
    """Phase B, drop_tracers.f90:20-55.  max_steps < 0 means run until converged (:40).

    Returns dict(steps, converged, vacf) with vacf[0] the init value (vacf.dat row 0).
    """
    if max_steps == 0:                      # :20-21
        return dict(steps=0, converged=False, vacf=np.zeros((0, 3)))
    v0 = sim.mp_init(Db, ka, kd, f_ext)     # update_tracer_population + init (:29-35)
    rows = [v0[None, :]]
    if max_steps < 0:
        max_steps = np.iinfo(np.int32).max
    it, conv = 0, False
    while it < max_steps and not conv:
        done, conv, v = sim.mp_step(min(chunk, max_steps - it))
        rows.append(v)
        it += done
    return dict(steps=it, converged=conv, vacf=np.concatenate(rows))
