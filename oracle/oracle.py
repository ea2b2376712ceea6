"""ctypes front-end to the CPU oracle (oracle/laboetie_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- see the header of laboetie_oracle.cpp.  Imported by
tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl
reference legs; never by the product package laboetie_b200.

PARITY UNPINNED: the reference has no golden vectors and cannot be built here.

Array conventions (numpy, C-contiguous):
  nature, interfacial  int8  (lz, ly, lx)            == Fortran (i,j,k), i fastest
  n (Phase A)          f64   (19, lz, ly, lx)        == Fortran n(i,j,k,l)
  ntr (Phase B)        f64   (lz, ly, lx, 19)        == Fortran n(l,i,j,k)
  density, jx, jy, jz  f64   (lz, ly, lx)
  P, Pads              f64   (2, lz, ly, lx, 3)      == Propagated_Quantity(x:z,i,j,k,now:next)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liblaboetie_oracle.so")
    src = os.path.join(_HERE, "laboetie_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liblaboetie_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        I, D = C.c_int, C.c_double
        dims = [I, I, I]
        L.orc_lbm_table.argtypes = [i32p, f64p, f64p, f64p, i32p]
        L.orc_geometry.argtypes = [I] + dims + [i8p]
        L.orc_read_geom_in.argtypes = [C.c_char_p] + dims + [i8p]
        L.orc_read_pbm.argtypes = [C.c_char_p] + dims + [i8p]
        L.orc_detect_interfacial.argtypes = dims + [i8p, i8p]
        L.orc_detect_interfacial.restype = None
        L.orc_init_populations.argtypes = dims + [i8p, D, f64p, f64p]
        L.orc_init_populations.restype = None
        L.orc_collide.argtypes = dims + [i8p, D] + [f64p] * 8
        L.orc_bounce_back.argtypes = dims + [i8p, f64p]
        L.orc_bounce_back.restype = None
        L.orc_stream.argtypes = dims + [f64p]
        L.orc_stream.restype = None
        L.orc_moments.argtypes = dims + [f64p] * 11 + [C.POINTER(D)]
        L.orc_lb_step.argtypes = dims + [i8p, D] + [f64p] * 11 + [C.POINTER(D)]
        L.orc_equilibration.argtypes = dims + [i8p, D, D, f64p, I] + [f64p] * 5 + [f64p, I, C.POINTER(I), C.POINTER(I)]
        L.orc_compensate_force.argtypes = dims + [i8p, f64p, I, I, I, I, I, f64p, f64p, f64p, C.POINTER(I)]
        L.orc_profiles.argtypes = dims + [f64p] * 4 + [I, f64p]
        L.orc_profiles.restype = None
        L.orc_total_flux.argtypes = dims + [f64p] * 3 + [f64p]
        L.orc_total_flux.restype = None
        L.orc_update_tracer_population.argtypes = dims + [i8p] + [f64p] * 4 + [f64p, D, f64p]
        L.orc_mp_init.argtypes = dims + [i8p, i8p, f64p, f64p, D, D, D, f64p, f64p, f64p, C.POINTER(I)]
        L.orc_mp_propagate.argtypes = dims + [i8p, i8p, f64p, f64p, D, D, D, I, I, f64p, f64p, f64p, C.POINTER(I)]
        L.orc_drop_tracers.argtypes = dims + [i8p, i8p] + [f64p] * 4 + [f64p, D, D, D, I, f64p, f64p, f64p]
        _LIB = L
    return _LIB


def _dims(a):
    lz, ly, lx = a.shape[-3:]
    return lx, ly, lz


def lbm_table():
    c = np.zeros((19, 3), np.int32)
    a0, a1, a2 = (np.zeros(19) for _ in range(3))
    inv = np.zeros(19, np.int32)
    lib().orc_lbm_table(c, a0, a1, a2, inv)
    return c, a0, a1, a2, inv


def geometry(label, lx, ly, lz):
    nat = np.zeros((lz, ly, lx), np.int8)
    rc = lib().orc_geometry(label, lx, ly, lz, nat)
    if rc:
        raise ValueError(f"orc_geometry(label={label}) -> {rc}")
    return nat


def read_geom_in(path, lx, ly, lz):
    nat = np.zeros((lz, ly, lx), np.int8)
    rc = lib().orc_read_geom_in(os.fsencode(path), lx, ly, lz, nat)
    if rc:
        raise ValueError(f"orc_read_geom_in({path}) -> {rc}")
    return nat


def read_pbm(path, lx, ly, lz):
    nat = np.zeros((lz, ly, lx), np.int8)
    rc = lib().orc_read_pbm(os.fsencode(path), lx, ly, lz, nat)
    if rc:
        raise ValueError(f"orc_read_pbm({path}) -> {rc}")
    return nat


def detect_interfacial(nature):
    out = np.zeros_like(nature)
    lib().orc_detect_interfacial(*_dims(nature), nature, out)
    return out


def init_populations(nature, rho0=1.0):
    n = np.zeros((19,) + nature.shape)
    rho = np.zeros(nature.shape)
    lib().orc_init_populations(*_dims(nature), nature, rho0, n, rho)
    return n, rho


class LBState:
    """The arrays equilibration.f90 keeps across steps (:59-96)."""

    def __init__(self, nature, rho0=1.0, tau=1.0):
        self.nature = np.ascontiguousarray(nature, np.int8)
        self.tau = tau
        self.n, self.rho = init_populations(self.nature, rho0)
        z = lambda: np.zeros(self.nature.shape)
        self.jx, self.jy, self.jz = z(), z(), z()
        self.jxo, self.jyo, self.jzo = z(), z(), z()
        self.fx, self.fy, self.fz = z(), z(), z()
        self.t = 0

    def set_force_uniform(self, f):
        """equilibration.f90:381-386"""
        fl = self.nature == 0
        self.fx[fl], self.fy[fl], self.fz[fl] = f[0], f[1], f[2]

    def step(self):
        """One body of the time loop; returns (rc, l2err); rc=1 means ANY(n<0)."""
        err = C.c_double()
        rc = lib().orc_lb_step(*_dims(self.nature), self.nature, self.tau, self.n, self.rho, self.jx, self.jy, self.jz,
                               self.jxo, self.jyo, self.jzo, self.fx, self.fy, self.fz, C.byref(err))
        self.t += 1
        return rc, err.value


def equilibration(nature, f_ext, tau=1.0, target_error=1e-10, rho0=1.0, max_steps=10**9, hist_cap=1 << 20):
    """Full Phase A (equilibration.f90:143-491). Returns dict."""
    nature = np.ascontiguousarray(nature, np.int8)
    n, rho = init_populations(nature, rho0)
    jx, jy, jz = (np.zeros(nature.shape) for _ in range(3))
    hist = np.zeros(hist_cap)
    te, tf = C.c_int(), C.c_int()
    rc = lib().orc_equilibration(*_dims(nature), nature, tau, target_error, np.asarray(f_ext, np.float64), max_steps,
                                 n, rho, jx, jy, jz, hist, hist_cap, C.byref(te), C.byref(tf))
    return dict(rc=rc, n=n, rho=rho, jx=jx, jy=jy, jz=jz, t_exit=te.value, t_fext=tf.value,
                l2err=hist[: min(te.value, hist_cap)].copy())


def compensate_force(nature, f_ext, pd=1, centre=None, geometry_label=0):
    """equilibration.f90:388-487; centre is 1-based (px,py,pz), default (n/2+1) as the reference."""
    lx, ly, lz = _dims(nature)
    if centre is None:
        centre = (lx // 2 + 1, ly // 2 + 1, lz // 2 + 1)
    fx, fy, fz = (np.zeros(nature.shape) for _ in range(3))
    l = C.c_int()
    rc = lib().orc_compensate_force(lx, ly, lz, nature, np.asarray(f_ext, np.float64), pd, *centre, geometry_label,
                                    fx, fy, fz, C.byref(l))
    if rc:
        raise ValueError(f"orc_compensate_force -> {rc}")
    return fx, fy, fz, l.value


def profiles(rho, jx, jy, jz, axis):
    lx, ly, lz = _dims(rho)
    out = np.zeros(((lx, ly, lz)[axis], 4))
    lib().orc_profiles(lx, ly, lz, rho, jx, jy, jz, axis, out)
    return out


def total_flux(jx, jy, jz):
    out = np.zeros(3)
    lib().orc_total_flux(*_dims(jx), jx, jy, jz, out)
    return out


def update_tracer_population(nature, rho, jx, jy, jz, f_ext, Db):
    ntr = np.zeros(nature.shape + (19,))
    rc = lib().orc_update_tracer_population(*_dims(nature), nature, rho, jx, jy, jz, np.asarray(f_ext, np.float64), Db, ntr)
    if rc:
        raise ValueError("The diffusion coefficient (tracer_Db in input file) is invalid")
    return ntr


class MPState:
    """moment_propagation module state: init (:30-160) then propagate (:164-289)."""

    def __init__(self, nature, interfacial, rho, jx, jy, jz, f_ext, Db, ka, kd):
        self.nature = np.ascontiguousarray(nature, np.int8)
        self.interfacial = np.ascontiguousarray(interfacial, np.int8)
        self.rho = np.ascontiguousarray(rho)
        self.Db, self.ka, self.kd = Db, ka, kd
        self.ntr = update_tracer_population(self.nature, self.rho, jx, jy, jz, f_ext, Db)
        self.P = np.zeros((2,) + self.nature.shape + (3,))
        self.Pads = np.zeros((2,) + self.nature.shape + (3,))
        self.vacf0 = np.zeros(3)
        ads = C.c_int()
        rc = lib().orc_mp_init(*_dims(self.nature), self.nature, self.interfacial, self.ntr, self.rho, Db, ka, kd,
                               self.P, self.Pads, self.vacf0, C.byref(ads))
        if rc:
            raise ValueError(f"orc_mp_init -> {rc}")
        self.ads = ads.value
        self.it = 0

    def propagate(self):
        """Returns (rc, vacf[3], is_converged); rc=1 means 'restpart is negative'."""
        self.it += 1
        v = np.zeros(3)
        conv = C.c_int()
        rc = lib().orc_mp_propagate(*_dims(self.nature), self.nature, self.interfacial, self.ntr, self.rho, self.Db,
                                    self.ka, self.kd, self.ads, self.it, self.P, self.Pads, v, C.byref(conv))
        return rc, v, bool(conv.value)
