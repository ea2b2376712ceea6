/* laboetie_gpu.h -- C ABI of the B200 (sm_100a) implementation of laboetie's
 * time-stepping hot path: D3Q19 BGK collision with Guo forcing, streaming with
 * halfway bounce-back, the density/momentum moments with their convergence
 * scalar, and tracer moment propagation with adsorption/desorption.
 *
 * The reference (maxlevesque/laboetie, Fortran 2008) has no FFI of its own; the
 * boundary is introduced at the three seams its drivers already have.  Each
 * entry point below names the reference code it replaces (file:line relative to
 * the reference tree).  The Fortran driver binds these through ISO_C_BINDING
 * (fortran/laboetie_gpu_iface.f90); a C++ mirror of the driver
 * (laboetie_b200/driver) and a ctypes mirror (laboetie_b200/api.py) bind the
 * same symbols.
 *
 * Conventions
 *   - plain C types only; every function returns an lbg_status (0 == ok);
 *   - host arrays are owned by the caller, never retained after return;
 *   - lattice arrays are in the reference's memory order: (i,j,k) with i (x)
 *     fastest, 0-based linear index i + lx*(j + ly*k); populations are
 *     n(i,j,k,l) with l slowest, l = 0..18 standing for the reference's 1..19
 *     (module_lbmodel.f90:66-86);
 *   - nature: int8, 0 = fluid, 1 = solid (module_system.f90:32);
 *   - one handle = one GPU = one z-slab [k0, k0+nzl) of the lattice; calls on a
 *     handle come from one host thread (the reference's drivers are serial);
 *   - there is no CPU fallback: without a CUDA device lbg_create* fails with
 *     LBG_ERR_NO_DEVICE.
 */
#ifndef LABOETIE_GPU_H
#define LABOETIE_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBG_ABI_VERSION 1
#define LBG_NVEL 19
#define LBG_UNIQUE_ID_BYTES 128

typedef struct lbg_handle_s* lbg_handle;

typedef enum {
  LBG_OK = 0,
  LBG_ERR_NEGATIVE_POPULATION = 1, /* equilibration.f90:248  "the population n(x,y,z,vel) < 0" */
  LBG_ERR_RESTPART_NEGATIVE = 2,   /* module_moment_propagation.f90:257 "somewhere restpart is negative" */
  LBG_ERR_RELAXATION_TIME = 3,     /* module_collision.f90:38-39  relaxation_time must be >0 and >=0.5 */
  LBG_ERR_TRACER_DB = 4,           /* drop_tracers.f90:89   tracer_Db <= epsilon */
  LBG_ERR_TRACER_KA_KD = 5,        /* module_moment_propagation.f90:43-44  ka or kd < 0 */
  LBG_ERR_ALL_SOLID = 6,           /* supercell_definition.f90:84-86 */
  LBG_ERR_INVALID_ARG = 7,
  LBG_ERR_STATE = 8,               /* call order violated (e.g. lb_step after mp_init) */
  LBG_ERR_UNSUPPORTED = 9,         /* first_order_only=T, tracer_Ds/=0, tracer_z/=0 (reference stops or is UB) */
  LBG_ERR_NO_DEVICE = 10,
  LBG_ERR_CUDA = 11,
  LBG_ERR_NCCL = 12,
  LBG_ERR_NOMEM = 13
} lbg_status;

/* ---- library ----------------------------------------------------------- */
int lbg_abi_version(void);
const char* lbg_status_string(int status);
/* last CUDA/NCCL error text recorded on this handle (or on the library if h==NULL) */
const char* lbg_last_error(lbg_handle h);
int lbg_device_count(int* count);

/* ---- decomposition (host logic, no GPU needed) -------------------------- */
/* Contiguous z-slabs, the reference's own OpenMP decomposition axis
 * (module_moment_propagation.f90:207).  rank r of nranks owns planes
 * [k0, k0+nzl); the first lz % nranks ranks get one extra plane. */
int lbg_partition(int lz, int nranks, int rank, int* k0, int* nzl);
/* Velocity indices (0-based) that cross a z-face: up[5] have cz=+1, down[5] have cz=-1. */
int lbg_halo_plan(int up[5], int down[5]);

/* ---- lifetime ----------------------------------------------------------- */
/* Whole lattice on one GPU.  Builds on the device what detectInterfacialNodes
 * (supercell_definition.f90:115-147) and the il/jl/kl neighbour tables
 * (equilibration.f90:109-119) provide: per-node link masks and the interfacial flag. */
int lbg_create(lbg_handle* h, int lx, int ly, int lz, const int8_t* nature, int device);
/* One z-slab of a decomposed lattice.  nature_halo has nzl+2 planes: global
 * planes k0-1 .. k0+nzl (periodic, module_system.f90:99-112). */
int lbg_create_slab(lbg_handle* h, int lx, int ly, int lz_global, int k0, int nzl,
                    const int8_t* nature_halo, int device);
/* the reference's exactly-representable geometry builders on the device
 * (supercell_definition.f90:50-59): geometryLabel -1 (bulk), 1 (slit, module_geometry.f90:158-166),
 * 2 (cylinder along z, :253-277), 3 (BCC spheres, :206-245).  Builds planes [k0, k0+nzl) of the
 * (lx, ly, lz_global) lattice plus their periodic halo planes without any host array; k0 = 0 and
 * nzl = lz_global gives the whole lattice on one GPU.  lbg_get_nature reads node%nature back. */
int lbg_create_geometry(lbg_handle* h, int label, int lx, int ly, int lz_global, int k0, int nzl, int device);
int lbg_get_nature(lbg_handle h, int8_t* nature);
int lbg_destroy(lbg_handle h);
/* Joins the slabs of one node into a ring.  Rank 0 obtains the 128-byte id (an NCCL unique id), the host
 * distributes it (torch.distributed / MPI / a file), every rank calls comm_init (collective).  The NCCL
 * communicator only bootstraps the exchange of CUDA IPC handles (a process keeps it and reuses it for later
 * handles of the same rank): per step the halo planes are pushed by the copy engines over NVLink into small
 * receive buffers the ring neighbours map, and the scalars (l2err, negative flag, vacf, counts) are all-reduced
 * through the peers' mailboxes; LBG_HALO=nccl selects ncclSend/ncclRecv + ncclAllReduce instead.
 * All stepping / set-up calls on slab handles are collective; the read-back calls are local. */
int lbg_comm_unique_id(void* id_out /* LBG_UNIQUE_ID_BYTES */);
int lbg_comm_init(lbg_handle h, int nranks, int rank, const void* id);

/* geometry read-back (own planes): interfacial flag as the reference defines it */
int lbg_get_interfacial(lbg_handle h, int8_t* interfacial);
/* counts over own planes: fluid nodes, interfacial fluid nodes */
int lbg_get_counts(lbg_handle h, int64_t* n_fluid, int64_t* n_interfacial_fluid);

/* ---- Phase A: Lattice-Boltzmann flow (equilibration.f90) ---------------- */
/* init_simu.f90:24-39: n_l = rho0*w_l on fluid, 0 on solid; density = rho0 on
 * fluid; equilibration.f90:75-80: j = 0; f_ext = 0 (equilibration.f90:94-98). */
int lbg_lb_init(lbg_handle h, double rho0);
/* Restart from a host state (own planes): populations n(i,j,k,l), and the
 * density / momentum density the next collide will consume. */
int lbg_lb_upload(lbg_handle h, const double* n, const double* rho, const double* jx, const double* jy,
                  const double* jz);
/* Memory-lean option (north_star "AA-pattern in place"): with on != 0 Phase A keeps ONE population
 * buffer (152 instead of 304 bytes per fluid node) and alternates a local and a pull/push kernel.
 * Results, l2err history and exit step are identical to the default two-lattice mode; the per-step
 * convergence scalar then costs a separate moments pass on checked steps (504 vs 352 bytes per node).
 * Call before lbg_lb_init / lbg_lb_upload.  Single-slab handles only. */
int lbg_lb_set_in_place(lbg_handle h, int on);
/* equilibration.f90:381-386: the same force on every fluid node, 0 on solid. */
int lbg_lb_set_force_uniform(lbg_handle h, const double f[3]);
/* equilibration.f90:388-487 (compensate_f_ext): arbitrary per-node force (own planes). */
int lbg_lb_set_force_field(lbg_handle h, const double* fx, const double* fy, const double* fz);
/* Up to nsteps bodies of the time loop equilibration.f90:143-350:
 * collide (module_collision.f90:15-130, second-order branch), bounce-back
 * (:204-222), streaming (:227-243), ANY(n<0) guard (:248), density (:254),
 * momentum density (:266-300) and l2err = max|j - j_old| (:339-343).
 * check_every = 1 reproduces the reference: l2err is evaluated every step and
 * the call returns right after the first step t (counted from lbg_lb_init, as
 * the reference's t) with l2err <= target_error and t > 2 (:346); the state is
 * then exactly the state after that step.  check_every = k > 1 evaluates it on
 * steps with t % k == 0 only; check_every = 0 never does.
 * l2err_hist (may be NULL) receives one value per executed step (NaN on
 * unchecked steps).  *converged is set when the criterion stopped the call. */
int lbg_lb_step(lbg_handle h, double tau, int nsteps, int check_every, double target_error,
                double* l2err_hist, int* steps_done, int* converged);
/* global step counter t (number of completed LB steps since lbg_lb_init/upload) */
int lbg_lb_time(lbg_handle h, int64_t* t);
/* equilibration.f90:551-554: density and momentum density after the last step (own planes). */
int lbg_lb_download_moments(lbg_handle h, double* rho, double* jx, double* jy, double* jz);
/* The same read-back without waiting for it: the call returns once the work is queued (the four host arrays
 * should be page-locked, otherwise the copies are synchronous anyway) and the arrays are valid after
 * lbg_wait_transfers.  Density and momentum stay resident, so the driver can go on with lbg_mp_init /
 * lbg_mp_step -- the reference's drop_tracers phase -- while 32 bytes per node cross PCIe. */
int lbg_lb_download_moments_async(lbg_handle h, double* rho, double* jx, double* jy, double* jz);
int lbg_wait_transfers(lbg_handle h);
/* system::n after the last completed step (parity / debugging; own planes). */
int lbg_lb_download_populations(lbg_handle h, double* n);
/* equilibration.f90:161-172,505-516: for each index p along axis (0=x,1=y,2=z)
 * out[4p..4p+3] = SUM(jx), SUM(jy), SUM(jz), SUM(density)/MAX(COUNT(density>eps),1).
 * Slab handles return partial sums/counts for x and y in out (5 values per row:
 * the three sums, the density sum and the count) when raw != 0. */
int lbg_lb_profiles(lbg_handle h, int axis, int raw, double* out);
/* equilibration.f90:260: SUM(jx), SUM(jy), SUM(jz) (own planes) */
int lbg_lb_total_flux(lbg_handle h, double out[3]);
/* equilibration.f90:526-548: one plane of density / momentum density for the 2-D field files
 * (mass-flux_field_2d_at_x.eq.1.dat, f_ext-field.dat, vel-field_central.dat) without reading the whole lattice back.
 * axis 0: x = index, arrays (ly, nzl) with j fastest; axis 1: y = index, (lx, nzl), i fastest; axis 2: own plane
 * k = index, (lx, ly), i fastest.  0-based index; any of the four outputs may be NULL; 0 on solid nodes. */
int lbg_lb_slice(lbg_handle h, int axis, int index, double* rho, double* jx, double* jy, double* jz);
/* equilibration.f90:187: jx,jy,jz,density at one node (0-based, own-plane k) */
int lbg_lb_probe(lbg_handle h, int i, int j, int k, double out[4]);

/* ---- Phase B: tracer moment propagation (drop_tracers.f90) -------------- */
/* update_tracer_population (drop_tracers.f90:63-105) + moment_propagation::init
 * (module_moment_propagation.f90:30-160) from the resident density / momentum
 * density.  f_ext is the lb.in force re-read at drop_tracers.f90:92.
 * tracer_Ds and tracer_z must be 0 (the reference stops otherwise).
 * vacf0 receives vacf(:,t=0).  The LB populations are released (the reference
 * deallocates n at drop_tracers.f90:85). */
int lbg_mp_init(lbg_handle h, double tracer_Db, double tracer_ka, double tracer_kd, const double f_ext[3],
                double vacf0[3]);
/* The same, from the driver's own arrays: drop_tracers.f90:63-105 reads density and momentum density
 * from node%solventdensity / node%solventflux (written back at equilibration.f90:551-554), so a driver
 * that restarts from saved fields, or that ran Phase A elsewhere, starts Phase B here without any
 * Lattice-Boltzmann state on the device.  rho, jx, jy, jz: (lx,ly,lz) arrays, i fastest (own planes). */
int lbg_mp_init_from_moments(lbg_handle h, const double* rho, const double* jx, const double* jy, const double* jz,
                             double tracer_Db, double tracer_ka, double tracer_kd, const double f_ext[3],
                             double vacf0[3]);
/* Up to nsteps calls of propagate (module_moment_propagation.f90:164-289).
 * vacf (may be NULL) receives vacf(:,now) of each executed step, 3 per step.
 * Returns after the first step `it` (counted from lbg_mp_init) with
 * it>2, all|vacf| < 1/(2 lx ly lz/Db) and all|vacf| < 1e-12 (:284). */
int lbg_mp_step(lbg_handle h, int nsteps, double* vacf, int* steps_done, int* converged);
/* Propagated_Quantity(x:z,i,j,k,now) and ..._Adsorbed(...,now), reference AoS order (own planes). */
int lbg_mp_download(lbg_handle h, double* P, double* Pads);

/* ---- measurement -------------------------------------------------------- */
/* CUDA events on the stream the kernels run on. */
int lbg_timer_start(lbg_handle h);
int lbg_timer_stop(lbg_handle h, float* milliseconds);
/* number of kernels this library has launched on this handle */
int lbg_launch_count(lbg_handle h, int64_t* launches);
/* Which code path a handle runs (diagnostics for tests and bench lines; no reference counterpart).  Keys:
 * "mp_neighbour_table" (Phase B resolves neighbours from its static table: 1, through rank lookups: 0),
 * "lb_variant" (100*pipelined + 10*tiles-per-chunk-is-dynamic + blocks per SM of the Phase-A kernel),
 * "in_place", "p2p" (peer-to-peer halos: 1, NCCL send/recv: 0), "ipc" (a neighbour's memory is mapped
 * through CUDA IPC, i.e. one process per GPU), "nranks", "rank", "fluid_nodes_with_halo".
 * Unknown key: LBG_ERR_INVALID_ARG. */
int lbg_get_info(lbg_handle h, const char* key, int64_t* value);
int lbg_sync(lbg_handle h);

#ifdef __cplusplus
}
#endif
#endif /* LABOETIE_GPU_H */
