"""ctypes binding of include/laboetie_gpu.h.

Array conventions (numpy, C-contiguous), identical to the reference's memory order:
  nature, interfacial   int8  (lz, ly, lx)        Fortran (i,j,k), i fastest
  populations n         f64   (19, lz, ly, lx)    Fortran n(i,j,k,l)
  rho, jx, jy, jz       f64   (lz, ly, lx)
  P, Pads               f64   (lz, ly, lx, 3)     Propagated_Quantity(x:z,i,j,k,now)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

STATUS = {
    0: "LBG_OK", 1: "LBG_ERR_NEGATIVE_POPULATION", 2: "LBG_ERR_RESTPART_NEGATIVE", 3: "LBG_ERR_RELAXATION_TIME",
    4: "LBG_ERR_TRACER_DB", 5: "LBG_ERR_TRACER_KA_KD", 6: "LBG_ERR_ALL_SOLID", 7: "LBG_ERR_INVALID_ARG",
    8: "LBG_ERR_STATE", 9: "LBG_ERR_UNSUPPORTED", 10: "LBG_ERR_NO_DEVICE", 11: "LBG_ERR_CUDA", 12: "LBG_ERR_NCCL",
    13: "LBG_ERR_NOMEM",
}

# every symbol include/laboetie_gpu.h declares
SYMBOLS = [
    "lbg_abi_version", "lbg_status_string", "lbg_last_error", "lbg_device_count", "lbg_partition", "lbg_halo_plan",
    "lbg_create", "lbg_create_slab", "lbg_create_geometry", "lbg_get_nature", "lbg_destroy", "lbg_comm_unique_id", "lbg_comm_init", "lbg_get_interfacial",
    "lbg_get_counts", "lbg_lb_set_in_place", "lbg_lb_init", "lbg_lb_upload", "lbg_lb_set_force_uniform", "lbg_lb_set_force_field",
    "lbg_lb_step", "lbg_lb_time", "lbg_lb_download_moments", "lbg_lb_download_moments_async", "lbg_wait_transfers", "lbg_lb_download_populations", "lbg_lb_profiles",
    "lbg_lb_total_flux", "lbg_lb_slice", "lbg_lb_probe", "lbg_mp_init", "lbg_mp_init_from_moments", "lbg_mp_step", "lbg_mp_download", "lbg_timer_start",
    "lbg_timer_stop", "lbg_launch_count", "lbg_get_info", "lbg_sync",
]


class LbgError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


def lib_path():
    # LBG_LIB lets a tuning run point at an alternative build of the same library
    return os.environ.get("LBG_LIB") or os.path.join(_HERE, "lib", "liblaboetie_gpu.so")


def load_library():
    """Load the CUDA library.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise LbgError(10, f"{path} is missing: build it with `python -m laboetie_b200.build` "
                           "(the product has no CPU path)")
    L = C.CDLL(path)
    I, D, P = C.c_int, C.c_double, C.c_void_p
    f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    i8 = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
    L.lbg_status_string.restype = C.c_char_p
    L.lbg_status_string.argtypes = [I]
    L.lbg_last_error.restype = C.c_char_p
    L.lbg_last_error.argtypes = [P]
    L.lbg_device_count.argtypes = [C.POINTER(I)]
    L.lbg_partition.argtypes = [I, I, I, C.POINTER(I), C.POINTER(I)]
    L.lbg_halo_plan.argtypes = [C.POINTER(I * 5), C.POINTER(I * 5)]
    L.lbg_create.argtypes = [C.POINTER(P), I, I, I, i8, I]
    L.lbg_create_slab.argtypes = [C.POINTER(P), I, I, I, I, I, i8, I]
    L.lbg_create_geometry.argtypes = [C.POINTER(P), I, I, I, I, I, I, I]
    L.lbg_get_nature.argtypes = [P, i8]
    L.lbg_destroy.argtypes = [P]
    L.lbg_comm_unique_id.argtypes = [C.c_char_p]
    L.lbg_comm_init.argtypes = [P, I, I, C.c_char_p]
    L.lbg_get_interfacial.argtypes = [P, i8]
    L.lbg_get_counts.argtypes = [P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.lbg_lb_set_in_place.argtypes = [P, I]
    L.lbg_lb_init.argtypes = [P, D]
    L.lbg_lb_upload.argtypes = [P, f64, f64, f64, f64, f64]
    L.lbg_lb_set_force_uniform.argtypes = [P, C.POINTER(D * 3)]
    L.lbg_lb_set_force_field.argtypes = [P, f64, f64, f64]
    L.lbg_lb_step.argtypes = [P, D, I, I, D, P, C.POINTER(I), C.POINTER(I)]
    L.lbg_lb_time.argtypes = [P, C.POINTER(C.c_int64)]
    L.lbg_lb_download_moments.argtypes = [P, f64, f64, f64, f64]
    L.lbg_lb_download_moments_async.argtypes = [P, f64, f64, f64, f64]
    L.lbg_wait_transfers.argtypes = [P]
    L.lbg_lb_download_populations.argtypes = [P, f64]
    L.lbg_lb_profiles.argtypes = [P, I, I, f64]
    L.lbg_lb_total_flux.argtypes = [P, f64]
    L.lbg_lb_slice.argtypes = [P, I, I, P, P, P, P]
    L.lbg_lb_probe.argtypes = [P, I, I, I, f64]
    L.lbg_mp_init.argtypes = [P, D, D, D, C.POINTER(D * 3), f64]
    L.lbg_mp_init_from_moments.argtypes = [P, f64, f64, f64, f64, D, D, D, C.POINTER(D * 3), f64]
    L.lbg_mp_step.argtypes = [P, I, P, C.POINTER(I), C.POINTER(I)]
    L.lbg_mp_download.argtypes = [P, P, P]
    L.lbg_timer_start.argtypes = [P]
    L.lbg_timer_stop.argtypes = [P, C.POINTER(C.c_float)]
    L.lbg_launch_count.argtypes = [P, C.POINTER(C.c_int64)]
    L.lbg_get_info.argtypes = [P, C.c_char_p, C.POINTER(C.c_int64)]
    L.lbg_sync.argtypes = [P]
    if L.lbg_abi_version() != 1:
        raise LbgError(7, "ABI version mismatch between api.py and liblaboetie_gpu.so")
    _LIB = L
    return L


def partition(lz, nranks, rank):
    k0, nzl = C.c_int(), C.c_int()
    rc = load_library().lbg_partition(lz, nranks, rank, C.byref(k0), C.byref(nzl))
    if rc:
        raise LbgError(rc, "lbg_partition")
    return k0.value, nzl.value


def halo_plan():
    up, down = (C.c_int * 5)(), (C.c_int * 5)()
    load_library().lbg_halo_plan(C.byref(up), C.byref(down))
    return list(up), list(down)


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = load_library().lbg_comm_unique_id(buf)
    if rc:
        raise LbgError(rc, load_library().lbg_last_error(None).decode())
    return buf.raw


class LaboetieGPU:
    """One GPU, one z-slab [k0, k0+nzl) of an (lx, ly, lz) lattice (the whole lattice by default)."""

    def __init__(self, nature=None, device=0, lz_global=None, k0=0, slab=False, label=None, shape=None, nzl=None):
        """nature: int8 (lz, ly, lx) -- or label + shape=(lx, ly, lz) to build geometryLabel -1/1/2/3 on the
        device (lbg_create_geometry), optionally only planes [k0, k0+nzl)."""
        self._L = load_library()
        self._h = C.c_void_p()
        if nature is None:
            lx, ly, lz = shape
            nzl = lz if nzl is None else nzl
            rc = self._L.lbg_create_geometry(C.byref(self._h), int(label), lx, ly, lz, k0, nzl, device)
            self.nzl, self.lz = nzl, lz
            if rc:
                self._h = C.c_void_p()
                raise LbgError(rc, self._L.lbg_last_error(None).decode())
            self.lx, self.ly, self.k0 = lx, ly, k0
            self.shape = (self.nzl, ly, lx)
            return
        nature = np.ascontiguousarray(nature, np.int8)
        if not slab:
            lz, ly, lx = nature.shape
            rc = self._L.lbg_create(C.byref(self._h), lx, ly, lz, nature, device)
            self.nzl, self.lz = lz, lz
        else:
            nzl2, ly, lx = nature.shape  # nzl + 2 planes: k0-1 .. k0+nzl
            rc = self._L.lbg_create_slab(C.byref(self._h), lx, ly, lz_global, k0, nzl2 - 2, nature, device)
            self.nzl, self.lz = nzl2 - 2, lz_global
        if rc:
            self._h = C.c_void_p()
            raise LbgError(rc, self._L.lbg_last_error(None).decode())
        self.lx, self.ly, self.k0 = lx, ly, k0
        self.shape = (self.nzl, ly, lx)

    # -- plumbing ---------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise LbgError(rc, self._L.lbg_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._L.lbg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def comm_init(self, nranks, rank, unique_id):
        self._ck(self._L.lbg_comm_init(self._h, nranks, rank, unique_id))

    def sync(self):
        self._ck(self._L.lbg_sync(self._h))

    def timer_start(self):
        self._ck(self._L.lbg_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self._L.lbg_timer_stop(self._h, C.byref(ms)))
        return ms.value

    @property
    def launches(self):
        n = C.c_int64()
        self._ck(self._L.lbg_launch_count(self._h, C.byref(n)))
        return n.value

    def info(self, key):
        """Which code path the handle runs (lbg_get_info): 'mp_neighbour_table', 'lb_variant', 'p2p', 'ipc', ..."""
        v = C.c_int64()
        self._ck(self._L.lbg_get_info(self._h, key.encode(), C.byref(v)))
        return v.value

    def _field(self, x, what, lead=()):
        """C-contiguous fp64 array of exactly the slab's shape: the library reads nown values per array."""
        a = np.ascontiguousarray(x, np.float64)
        if a.shape != tuple(lead) + self.shape:
            raise LbgError(7, f"{what}: shape {a.shape}, expected {tuple(lead) + self.shape}")
        return a

    # -- geometry ---------------------------------------------------------
    def interfacial(self):
        out = np.zeros(self.shape, np.int8)
        self._ck(self._L.lbg_get_interfacial(self._h, out))
        return out

    def nature(self):
        out = np.zeros(self.shape, np.int8)
        self._ck(self._L.lbg_get_nature(self._h, out))
        return out

    def counts(self):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self._L.lbg_get_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- Phase A ----------------------------------------------------------
    def lb_set_in_place(self, on=True):
        """AA pattern: one population buffer; call before lb_init / lb_upload."""
        self._ck(self._L.lbg_lb_set_in_place(self._h, int(bool(on))))

    def lb_init(self, rho0=1.0):
        self._ck(self._L.lbg_lb_init(self._h, rho0))

    def lb_upload(self, n, rho, jx, jy, jz):
        a = [self._field(n, "lb_upload n", (19,))] + [self._field(x, "lb_upload moments") for x in (rho, jx, jy, jz)]
        self._ck(self._L.lbg_lb_upload(self._h, *a))

    def lb_set_force_uniform(self, f):
        self._ck(self._L.lbg_lb_set_force_uniform(self._h, C.byref((C.c_double * 3)(*[float(v) for v in f]))))

    def lb_set_force_field(self, fx, fy, fz):
        a = [self._field(x, "lb_set_force_field") for x in (fx, fy, fz)]
        self._ck(self._L.lbg_lb_set_force_field(self._h, *a))

    def lb_step(self, nsteps, tau=1.0, check_every=1, target_error=1e-10, want_history=True):
        """Returns (steps_done, converged, l2err history)."""
        hist = np.full(max(nsteps, 1), np.nan) if want_history else None
        done, conv = C.c_int(), C.c_int()
        rc = self._L.lbg_lb_step(self._h, tau, nsteps, check_every, target_error,
                                 hist.ctypes.data_as(C.c_void_p) if want_history else None, C.byref(done), C.byref(conv))
        self.last_steps_done = done.value
        self._ck(rc)
        return done.value, bool(conv.value), (hist[: done.value] if want_history else None)

    @property
    def t(self):
        t = C.c_int64()
        self._ck(self._L.lbg_lb_time(self._h, C.byref(t)))
        return t.value

    def lb_moments(self):
        out = [np.zeros(self.shape) for _ in range(4)]
        self._ck(self._L.lbg_lb_download_moments(self._h, *out))
        return out

    def lb_moments_async(self, out):
        """Queue the read-back of density and momentum into the four (preferably pinned) arrays `out`; they are
        valid after wait_transfers().  Phase B may start meanwhile."""
        for a in out:
            assert a.shape == self.shape and a.dtype == np.float64 and a.flags.c_contiguous
        self._ck(self._L.lbg_lb_download_moments_async(self._h, *out))

    def wait_transfers(self):
        self._ck(self._L.lbg_wait_transfers(self._h))

    def lb_populations(self):
        n = np.zeros((19,) + self.shape)
        self._ck(self._L.lbg_lb_download_populations(self._h, n))
        return n

    def lb_profiles(self, axis, raw=False):
        rows = (self.lx, self.ly, self.nzl)[axis]
        out = np.zeros((rows, 5 if raw else 4))
        self._ck(self._L.lbg_lb_profiles(self._h, axis, int(raw), out))
        return out

    def lb_total_flux(self):
        out = np.zeros(3)
        self._ck(self._L.lbg_lb_total_flux(self._h, out))
        return out

    def lb_slice(self, axis, index):
        """rho, jx, jy, jz on one plane (axis 0: x = index -> (nzl, ly); 1: y = index -> (nzl, lx); 2: z = index -> (ly, lx))."""
        shape = [(self.nzl, self.ly), (self.nzl, self.lx), (self.ly, self.lx)][axis]
        out = [np.zeros(shape) for _ in range(4)]
        self._ck(self._L.lbg_lb_slice(self._h, axis, index, *[a.ctypes.data_as(C.c_void_p) for a in out]))
        return out

    def lb_probe(self, i, j, k):
        out = np.zeros(4)
        self._ck(self._L.lbg_lb_probe(self._h, i, j, k, out))
        return out

    # -- Phase B ----------------------------------------------------------
    def mp_init(self, Db, ka, kd, f_ext):
        v0 = np.zeros(3)
        self._ck(self._L.lbg_mp_init(self._h, Db, ka, kd, C.byref((C.c_double * 3)(*[float(v) for v in f_ext])), v0))
        return v0

    def mp_init_from_moments(self, rho, jx, jy, jz, Db, ka, kd, f_ext):
        """Phase B from the driver's own density / momentum arrays (shape (nzl, ly, lx)); no LB state needed."""
        v0 = np.zeros(3)
        arrs = [self._field(a, "mp_init_from_moments") for a in (rho, jx, jy, jz)]
        self._ck(self._L.lbg_mp_init_from_moments(self._h, *arrs, Db, ka, kd,
                                                  C.byref((C.c_double * 3)(*[float(v) for v in f_ext])), v0))
        return v0

    def mp_step(self, nsteps, want_history=True):
        """Returns (steps_done, converged, vacf history (steps_done, 3))."""
        v = np.zeros((max(nsteps, 1), 3)) if want_history else None
        done, conv = C.c_int(), C.c_int()
        rc = self._L.lbg_mp_step(self._h, nsteps, v.ctypes.data_as(C.c_void_p) if want_history else None,
                                 C.byref(done), C.byref(conv))
        self._ck(rc)
        return done.value, bool(conv.value), (v[: done.value] if want_history else None)

    def mp_download(self, want_ads=True):
        P = np.zeros(self.shape + (3,))
        A = np.zeros(self.shape + (3,)) if want_ads else None
        self._ck(self._L.lbg_mp_download(self._h, P.ctypes.data_as(C.c_void_p),
                                         A.ctypes.data_as(C.c_void_p) if want_ads else None))
        return P, A
