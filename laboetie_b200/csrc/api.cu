// api.cu -- host side of the C ABI declared in include/laboetie_gpu.h.
//
// One handle owns one GPU and one z-slab.  Storage is fluid-compacted (lbg_internal.h): per fluid
// node of the slab (own planes + one halo plane each side), fp64 SoA with stride nfa:
//   f[2]   2 x 19   populations, two-lattice (source / destination of a step)
//   mom    4        density, jx, jy, jz as the driver sees them
//   jpp[2] 2 x 3    momentum density of the last two steps (for max|j - j_old|)
//   gidx   1 u32    dense node index, bit 31 = interfacial
// plus the rank structure words[] (8 bytes per 32 lattice nodes).  Phase B reuses f[0] for the 18 link
// probabilities and f[1] for the remaining fraction / u*, both time levels of Propagated_Quantity and
// of the adsorbed quantity (the reference drops its populations there too, drop_tracers.f90:85).
// The driver's dense (i,j,k) arrays are scattered / gathered at the phase boundaries.
//
// Multi-GPU: the slab ring exchanges, per step, the 5 populations leaving each z-face.  A plane's fluid nodes are one
// contiguous fid range, so a face is 5 contiguous runs that the copy engines push over NVLink into the
// neighbour's small IPC-mapped receive buffer (unpacked into the halo ranges by halo_unpack_kernel), overlapped
// with the interior planes' kernel; scalars are all-reduced by a one-warp kernel over the peers' mailboxes.
// NCCL (loaded with dlopen only when lbg_comm_init is called) bootstraps the handle exchange and is the fallback
// transport (LBG_HALO=nccl).
#include <dlfcn.h>
#include <nccl.h>
#include <time.h>
#include <unistd.h>

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/laboetie_gpu.h"
#include "lattice.cuh"

using namespace lbg;

namespace {

constexpr int SLOT_CAP = 4096;  // steps per batch (one host sync per batch)
constexpr int STAGE_SLOTS = 4;     // dense staging arrays for host transfers: the four moment arrays can be in flight at once
constexpr int F0_ARRAYS = 19 + 4;  // arrays in the f[0] allocation: 19 populations + density, jx, jy, jz

thread_local std::string g_last_error;  // per thread: several slabs may be driven from threads of one process

// LBG_TIMING=1: host-side phase timings of the set-up entry points on stderr (e2e tuning)
struct PhaseTimer {
  bool on;
  const char* what;
  double t0, tl;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  }
  explicit PhaseTimer(const char* w) : on(std::getenv("LBG_TIMING") != nullptr), what(w), t0(0), tl(0) {
    if (on) t0 = tl = now();
  }
  void lap(const char* name) {
    if (!on) return;
    const double t = now();
    std::fprintf(stderr, "[lbg timing] %s: %-28s %8.2f ms\n", what, name, (t - tl) * 1e3);
    tl = t;
  }
  ~PhaseTimer() {
    if (on) std::fprintf(stderr, "[lbg timing] %s: total %8.2f ms\n", what, (now() - t0) * 1e3);
  }
};

struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
#define LBG_SYM(f) *(void**)(&f) = dlsym(lib, "nccl" #f)
    LBG_SYM(GetUniqueId);
    LBG_SYM(CommInitRank);
    LBG_SYM(CommDestroy);
    LBG_SYM(GroupStart);
    LBG_SYM(GroupEnd);
    LBG_SYM(Send);
    LBG_SYM(Recv);
    LBG_SYM(AllReduce);
    LBG_SYM(GetErrorString);
#undef LBG_SYM
    return GetUniqueId && CommInitRank && CommDestroy && GroupStart && GroupEnd && Send && Recv && AllReduce;
  }
} g_nccl;

// communicators this process has created (see lbg_comm_init)
struct CommEntry {
  int nranks, rank, device;
  ncclComm_t comm;
  bool in_use;
};
std::mutex g_comm_mu;
std::vector<CommEntry> g_comms;

struct Force {
  int mode = FORCE_NONE;  // FORCE_NONE / FORCE_UNIFORM / FORCE_FIELD
  double u[3] = {0, 0, 0};
  double* field = nullptr;  // 3 arrays, stride nfa (owned)
};

enum Phase { PH_CREATED = 0, PH_LB = 1, PH_MP = 2 };

// Strip order for the propagate kernel: segments (strip s, plane p) of consecutive fids, strip-major.
struct SegTable {
  int nseg = 0, ntiles = 0;
  int* tile_cum = nullptr;
  long long* seg_begin = nullptr;
  long long* seg_end = nullptr;
  void release() {
    cudaFree(tile_cum);
    cudaFree(seg_begin);
    cudaFree(seg_end);
    *this = SegTable();
  }
};

}  // namespace

struct lbg_handle_s {
  int device = 0;
  int sm_count = 148;
  Geo geo{};
  int lz_global = 0, k0 = 0;
  long long nown = 0;    // lattice nodes of the own planes
  long long nf = 0;      // fluid nodes of the slab incl. halo planes
  long long nwords = 0;
  std::vector<long long> pstart;  // first fid of plane p, p = 0..nzl+2
  std::string err;
  long long launches = 0;

  int nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t st = nullptr, st_comm = nullptr, st_ar = nullptr;  // compute, halo traffic, scalar all-reduces
  cudaEvent_t ev_ready = nullptr, ev_halo = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  cudaEvent_t ev_ar[2] = {nullptr, nullptr};  // lagged vacf all-reduces of Phase B
  cudaEvent_t ev_arb = nullptr;               // blocking all-reduce on the all-reduce stream
  bool transfers_pending = false;  // an asynchronous read-back is using the staging buffer (lbg_wait_transfers)
  cudaEvent_t ev_stage[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // read-back pipeline (copy_own_to_host_many)
  bool halo_pending = false;
  // peer-to-peer halos (NVLink, copy engines): neighbours' population buffers and arrival flags
  bool p2p = false;
  // Halo planes arrive in a small receive buffer the two ring neighbours map (CUDA IPC) -- mapping the lattices
  // themselves costs 0.3-0.9 s per 12 GB buffer and handle -- and are copied from there into the halo ranges of
  // the arrays by a small kernel on the compute stream (halo_unpack_kernel).  Layout, in doubles:
  // [parity 2][side 2: 0 = from the lower neighbour, 1 = from the upper][array HALO_ARRAYS][halo_cap].
  double* halo_stage = nullptr;
  long long halo_cap = 0;
  double* peer_stage[2] = {nullptr, nullptr};                       // [down/up] neighbour's receive buffer
  long long peer_cap[2] = {0, 0};
  unsigned int* peer_flags[2] = {nullptr, nullptr};                 // [down/up]
  bool peer_ipc[2] = {false, false};                                // opened with cudaIpcOpenMemHandle
  struct PendingUnpack {
    bool on = false;
    double* base = nullptr;
    int par = 0, nlo = 0, nhi = 0;
    int reverse = 0;   // 1: the in-place scheme's return trip (masked merge into the own boundary planes)
    int lo_list[5] = {0, 0, 0, 0, 0}, hi_list[5] = {0, 0, 0, 0, 0};
  } unpack;
  // `mail` is one small allocation every peer of the job maps: [0..15] 32-bit halo arrival flags (flags[0] = halo data
  // from the lower neighbour has arrived up to this sequence number, [1] = upper), then from MAIL_OFF (in 64-bit
  // words) the scalar all-reduce mailbox: 2 parities x nranks senders x MAIL_WORDS words {sequence, payload...}.
  unsigned long long* mail = nullptr;
  std::vector<unsigned long long*> peer_mail;   // every rank's mailbox base (own included), host copy
  std::vector<bool> peer_mail_ipc;
  unsigned long long** d_peer_mail = nullptr;   // the same on the device
  unsigned long long ar_seq = 0;                // all-reduces issued so far
  long long spin_clocks = 40000000000LL;        // bounded spins of the peer-to-peer waits (~20 s; LBG_P2P_TIMEOUT_S)
  int comm_slot = -1;                           // entry of the process-wide communicator cache in use (-1: own comm)
  unsigned int* flags = nullptr;   // = (unsigned int*)mail
  unsigned int xseq = 0;           // exchanges issued so far
  unsigned int xwait = 0;          // sequence number the next kernels must see in flags[]
  int* p2p_err = nullptr;          // set by a wait kernel that timed out

  uint2* words = nullptr;
  uint32_t* gidx = nullptr;
  double* f[2] = {nullptr, nullptr};
  double* mom = nullptr;
  double* jpp[2] = {nullptr, nullptr};
  double* stage = nullptr;  // dense staging buffer for host transfers (3 * nown doubles, lazily)
  unsigned long long* l2_slots = nullptr;
  double* vacf_slots = nullptr;
  double* partial = nullptr;
  Ctrl* ctrl = nullptr;
  int* mp_err = nullptr;
  unsigned long long* counts = nullptr;
  // pinned host staging
  unsigned long long* h_l2 = nullptr;
  double* h_vacf = nullptr;
  Ctrl* h_ctrl = nullptr;
  unsigned char* h_small = nullptr;  // scratch for small reads (read_small_async)

  d3q19::Consts k{};
  int grid_lb = 148, grid_mp = 148;
  int lb_minb = 2;  // register-allocation variant of the LB step kernel (see lb_kernels.cu)
  int lb_pipe = 1;             // software-pipelined step kernel (lb_kernels.cu lb_step_pipe_kernel)
  int mp_arith = 1;            // Phase-B rank-lookup path: arithmetic neighbour ids on regular nodes (LBG_MP_ARITH=0 disables)
  int mp_tpc = 0;              // Phase-B kernel: consecutive tiles per CTA of a grid that covers the tiles (0: persistent grid)
  int lb_tpc = 0;              // tiles per chunk of the dynamic tile schedule of the Phase-A kernels (Geo::tpc; 0 = static)
  int64_t n_fluid = 0, n_if_fluid = 0;  // own planes

  Phase phase = PH_CREATED;
  // Phase A
  bool in_place = false;     // AA pattern: populations live in f[0] only (lb_aa_kernels.cu)
  bool aa_swapped = false;   // layout of f[0] in in-place mode: false = N(t), true = S(t)
  int grid_aa = 148;
  long long t = 0;
  bool precollision = true;  // f[src] holds n(t) (not yet collided) since init/upload
  int src = 0;               // f[src] = n*(t) (or n(t) if precollision); f[1-src] = n*(t+1) when collided_ok
  int jc = 0;                // jpp[jc] = j(t) when j_valid_step == t
  bool collided_ok = false;
  double collided_tau = 0;
  unsigned long long force_version = 0, collided_force_version = 0;
  long long j_valid_step = -1, mom_valid_step = 0;
  Force fcur, fprev;
  bool prev_equals_cur = true;
  // Phase B
  long long it = 0;
  int pc = 0;  // P[pc] = now
  double Db = 0, ka = 0, kd = 0;
  int ads = 0;
  int mp_bad = 0;
  double* q = nullptr;
  uint32_t* nbt01 = nullptr;  // Phase-B neighbour table (lbg_internal.h NBT_*): words 0,1 in the spare array of f[0],
  uint32_t* nbt27 = nullptr;  // words 2..7 in the three spare arrays of f[1]
  int mp_use_nbt = 1;         // off on narrow lattices, where every warp holds nodes of the periodic x seam
  double* s = nullptr;
  double* P[2] = {nullptr, nullptr};
  double* A[2] = {nullptr, nullptr};
  uint32_t* rwords = nullptr; // Phase B, rank-lookup path: "regular" bits per group of 32 fids (mp_kernels.cu)
  uint2* awords = nullptr;    // compact adsorbed storage: per group of 32 fids {interfacial bits, first slot} (lbg_internal.h)
  long long a_stride = 0;     // slots per component of A[.]
  SegTable strips;  // over the planes one propagate launch covers (all own planes, or the interior ones)
  SegTable lb_strips;  // the same for the Phase-A step kernel (tiles aligned to 32 fids)
  bool lb_strips_built = false;
};

namespace {

int fail(lbg_handle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_last_error = msg;
  return code;
}

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(h, e_ == cudaErrorMemoryAllocation ? LBG_ERR_NOMEM : LBG_ERR_CUDA,                \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                               \
  } while (0)

#define NK(call)                                                                                     \
  do {                                                                                               \
    ncclResult_t r_ = (call);                                                                        \
    if (r_ != ncclSuccess)                                                                           \
      return fail(h, LBG_ERR_NCCL,                                                                   \
                  std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
  } while (0)

#define RET(call)               \
  do {                          \
    int rc_ = (call);           \
    if (rc_ != LBG_OK) return rc_; \
  } while (0)

// module_lbmodel.f90:122-136, every operation rounded to fp64 like the Fortran PARAMETERs
d3q19::Consts make_consts() {
  d3q19::Consts k;
  volatile double one = 1.0, three = 3.0, e18 = 18.0, e36 = 36.0;
  const double csq = one / three;
  const double w[3] = {one / three, one / e18, one / e36};
  volatile double csq2 = csq * csq;
  for (int i = 0; i < 3; ++i) {
    k.a0[i] = w[i];
    k.a1[i] = w[i] / csq;
    k.a2[i] = w[i] / (2 * csq2);
    k.two_a2[i] = 2.0 * k.a2[i];
  }
  k.csq = csq;
  volatile double c1 = one - csq;
  k.c1 = c1;
  k.mcsq = 0.0 - csq;
  return k;
}

// fid ranges
long long own_begin(const lbg_handle h) { return h->pstart[1]; }
long long own_end(const lbg_handle h) { return h->pstart[h->geo.nzl + 1]; }

int up_rank(const lbg_handle h) { return (h->rank + 1) % h->nranks; }
int down_rank(const lbg_handle h) { return (h->rank + h->nranks - 1) % h->nranks; }

// peer flag store after the pushes of one exchange (copy engine traffic precedes it in stream order)
__global__ void p2p_signal_kernel(unsigned int* flag_a, unsigned int* flag_b, unsigned int seq) {
  __threadfence_system();
  if (flag_a) *(volatile unsigned int*)flag_a = seq;
  if (flag_b) *(volatile unsigned int*)flag_b = seq;
  __threadfence_system();
}

// wait until both neighbours' halo data of exchange `seq` have landed; bounded spin (no GPU hang if a
// neighbour died): on time-out the control block's stop flag is raised and *err set
__global__ void p2p_wait_kernel(const unsigned int* flags, unsigned int seq, Ctrl* ctrl, int* err, long long max_clocks) {
  const long long t0 = clock64();
  for (;;) {
    const unsigned int a = *(volatile const unsigned int*)&flags[0];
    const unsigned int b = *(volatile const unsigned int*)&flags[1];
    if ((int)(a - seq) >= 0 && (int)(b - seq) >= 0) break;
    if (clock64() - t0 > max_clocks) {  // LBG_P2P_TIMEOUT_S, default ~20 s
      ctrl->stop = 1;
      *err = 2;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

constexpr int HALO_ARRAYS = 5;  // arrays per face and exchange: 5 populations (Phase A), 3 (P), 4 (density, momentum)

struct UnpackArgs {
  const double* stage_lo;  // receive slots filled by the lower neighbour (array i at i * cap)
  const double* stage_hi;  // ... by the upper neighbour
  double* base;            // arrays of stride nfa
  long long nfa, cap;
  long long lo_begin, lo_cnt, hi_begin, hi_cnt;  // destination fid ranges: my lower / upper halo plane (forward),
                                                 // my bottom / top own plane (reverse)
  int nlo, nhi;
  int lo_list[5], hi_list[5];
  // reverse trip of the in-place (AA) scheme: array m of a boundary-plane node r is taken only if its owner
  // node r - c_m (in the halo plane, i.e. in the sender's own boundary plane) is fluid -- a slot whose owner is
  // solid was never written by the sender and holds my own bounce-back value (tests/test_aa_slab_protocol_model.py)
  int masked;
  Geo geo;
};

// receive buffer -> halo ranges of the arrays.  It sits between two steps on the compute stream, so it is
// spread over the whole GPU: chunks of 4 * BLOCK elements, four independent loads per thread in flight.
__global__ void __launch_bounds__(BLOCK) halo_unpack_kernel(const __grid_constant__ UnpackArgs a) {
  constexpr int U = 4;
  const long long per_lo = (a.lo_cnt + U * BLOCK - 1) / (U * BLOCK), per_hi = (a.hi_cnt + U * BLOCK - 1) / (U * BLOCK);
  const long long nchunks = a.nlo * per_lo + a.nhi * per_hi;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const bool lo = c < a.nlo * per_lo;
    const long long cc = lo ? c : c - a.nlo * per_lo;
    const long long per = lo ? per_lo : per_hi;
    const int i = (int)(cc / per);
    const long long q0 = (cc - (long long)i * per) * (U * BLOCK) + threadIdx.x;
    const double* src = (lo ? a.stage_lo : a.stage_hi) + (long long)i * a.cap;
    double* dst = a.base + (long long)(lo ? a.lo_list[i] : a.hi_list[i]) * a.nfa + (lo ? a.lo_begin : a.hi_begin);
    const long long n = lo ? a.lo_cnt : a.hi_cnt;
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = (q0 + u * BLOCK < n) ? __ldcs(src + q0 + u * BLOCK) : 0.0;
    if (!a.masked) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (q0 + u * BLOCK < n) dst[q0 + u * BLOCK] = v[u];
    } else {
      const int m = lo ? a.lo_list[i] : a.hi_list[i];
      const int X = d3q19::cx(m), Y = d3q19::cy(m), Z = d3q19::cz(m);
      const long long fid0 = lo ? a.lo_begin : a.hi_begin;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = q0 + u * BLOCK;
        if (q >= n) continue;
        const int g = (int)(a.geo.gidx[fid0 + q] & GIDX_MASK);
        const Nb nb = neighbours(a.geo, g);
        const int o = (X > 0 ? nb.oxm : (X < 0 ? nb.oxp : 0)) + (Y > 0 ? nb.oym : (Y < 0 ? nb.oyp : 0)) +
                      (Z > 0 ? nb.ozm : (Z < 0 ? nb.ozp : 0));
        int fo;
        if (lookup(a.geo, g + o, fo)) dst[q] = v[u];
      }
    }
  }
}

int wait_halo(lbg_handle h);

// reverse = false: my top own plane of the up_list arrays -> the upper neighbour's lower halo, my bottom own plane of
// the down_list arrays -> the lower neighbour's upper halo.  reverse = true (in-place scheme after a pull/push
// step): my upper HALO plane of the up_list arrays -> the upper neighbour's bottom own plane, my lower halo plane of
// the down_list arrays -> the lower neighbour's top own plane, merged there under the owner mask.
int halo_exchange_p2p(lbg_handle h, double* base, const int* up_list, int nup, const int* down_list, int ndown,
                      bool reverse = false) {
  const Geo& g = h->geo;
  const std::vector<long long>& ps = h->pstart;
  const int nz = g.nzl;
  if (nup > HALO_ARRAYS || ndown > HALO_ARRAYS) return fail(h, LBG_ERR_INVALID_ARG, "halo exchange: too many arrays");
  // the previous exchange must have been consumed (its receive slots of the same parity come up again two
  // exchanges later, and the neighbours may only overwrite them after my unpack: see the header comment)
  if (h->unpack.on || h->xwait) RET(wait_halo(h));
  const long long up_src = reverse ? ps[nz + 1] : ps[nz], down_src = reverse ? ps[0] : ps[1];
  const size_t top_cnt = (size_t)(reverse ? ps[nz + 2] - ps[nz + 1] : ps[nz + 1] - ps[nz]);
  const size_t bot_cnt = (size_t)(reverse ? ps[1] - ps[0] : ps[2] - ps[1]);
  CK(cudaEventRecord(h->ev_ready, h->st));
  CK(cudaStreamWaitEvent(h->st_comm, h->ev_ready, 0));
  ++h->xseq;
  const int par = (int)(h->xseq & 1u);
  for (int i = 0; i < nup && top_cnt; ++i) {   // -> the upper neighbour: its side 0 (data from its lower neighbour)
    const double* src = base + (long long)up_list[i] * g.nfa + up_src;
    double* dst = h->peer_stage[1] + ((long long)(par * 2 + 0) * HALO_ARRAYS + i) * h->peer_cap[1];
    CK(cudaMemcpyAsync(dst, src, top_cnt * sizeof(double), cudaMemcpyDeviceToDevice, h->st_comm));
  }
  for (int i = 0; i < ndown && bot_cnt; ++i) {  // -> the lower neighbour: its side 1 (data from its upper neighbour)
    const double* src = base + (long long)down_list[i] * g.nfa + down_src;
    double* dst = h->peer_stage[0] + ((long long)(par * 2 + 1) * HALO_ARRAYS + i) * h->peer_cap[0];
    CK(cudaMemcpyAsync(dst, src, bot_cnt * sizeof(double), cudaMemcpyDeviceToDevice, h->st_comm));
  }
  // upper neighbour: I am its lower side -> flags[0]; lower neighbour: I am its upper side -> flags[1]
  p2p_signal_kernel<<<1, 1, 0, h->st_comm>>>(h->peer_flags[1] + 0, h->peer_flags[0] + 1, h->xseq);
  h->launches += 1;
  CK(cudaEventRecord(h->ev_halo, h->st_comm));
  h->halo_pending = true;
  h->xwait = h->xseq;
  // what arrives for me: the lower neighbour's top plane of the up_list arrays, the upper neighbour's bottom
  // plane of the down_list arrays
  h->unpack.on = true;
  h->unpack.base = base;
  h->unpack.par = par;
  h->unpack.reverse = reverse ? 1 : 0;
  h->unpack.nlo = nup;
  h->unpack.nhi = ndown;
  for (int i = 0; i < nup; ++i) h->unpack.lo_list[i] = up_list[i];
  for (int i = 0; i < ndown; ++i) h->unpack.hi_list[i] = down_list[i];
  return LBG_OK;
}

// Exchange boundary planes with the ring neighbours.  up_list / down_list name the arrays (index into
// base, stride nfa) whose top own plane goes to the upper neighbour's lower halo / whose bottom own
// plane goes to the lower neighbour's upper halo.  Runs on st_comm after everything enqueued on st.
int halo_exchange(lbg_handle h, double* base, const int* up_list, int nup, const int* down_list, int ndown,
                  bool reverse = false) {
  if (h->nranks == 1) return LBG_OK;
  if (h->p2p) return halo_exchange_p2p(h, base, up_list, nup, down_list, ndown, reverse);
  if (reverse) return fail(h, LBG_ERR_UNSUPPORTED, "the in-place (AA) scheme across slabs needs the peer-to-peer halo transport");
  const Geo& g = h->geo;
  const std::vector<long long>& ps = h->pstart;
  const int nz = g.nzl;
  const size_t top_cnt = (size_t)(ps[nz + 1] - ps[nz]), bot_cnt = (size_t)(ps[2] - ps[1]);
  const size_t lo_halo_cnt = (size_t)(ps[1] - ps[0]), hi_halo_cnt = (size_t)(ps[nz + 2] - ps[nz + 1]);
  CK(cudaEventRecord(h->ev_ready, h->st));
  CK(cudaStreamWaitEvent(h->st_comm, h->ev_ready, 0));
  NK(g_nccl.GroupStart());
  for (int i = 0; i < nup; ++i) {
    double* a = base + (long long)up_list[i] * g.nfa;
    if (top_cnt) NK(g_nccl.Send(a + ps[nz], top_cnt, ncclDouble, up_rank(h), h->comm, h->st_comm));
    if (lo_halo_cnt) NK(g_nccl.Recv(a + ps[0], lo_halo_cnt, ncclDouble, down_rank(h), h->comm, h->st_comm));
  }
  for (int i = 0; i < ndown; ++i) {
    double* a = base + (long long)down_list[i] * g.nfa;
    if (bot_cnt) NK(g_nccl.Send(a + ps[1], bot_cnt, ncclDouble, down_rank(h), h->comm, h->st_comm));
    if (hi_halo_cnt) NK(g_nccl.Recv(a + ps[nz + 1], hi_halo_cnt, ncclDouble, up_rank(h), h->comm, h->st_comm));
  }
  NK(g_nccl.GroupEnd());
  CK(cudaEventRecord(h->ev_halo, h->st_comm));
  h->halo_pending = true;
  return LBG_OK;
}

int wait_halo(lbg_handle h) {
  if (h->halo_pending) {
    CK(cudaStreamWaitEvent(h->st, h->ev_halo, 0));
    h->halo_pending = false;
  }
  if (h->p2p && h->xwait) {  // the neighbours' pushes of the latest exchange must have landed
    p2p_wait_kernel<<<1, 1, 0, h->st>>>(h->flags, h->xwait, h->ctrl, h->p2p_err, h->spin_clocks);
    h->launches += 1;
    h->xwait = 0;
  }
  if (h->p2p && h->unpack.on) {  // receive buffer -> halo ranges
    const std::vector<long long>& ps = h->pstart;
    const int nz = h->geo.nzl;
    UnpackArgs a{};
    a.stage_lo = h->halo_stage + (long long)(h->unpack.par * 2 + 0) * HALO_ARRAYS * h->halo_cap;
    a.stage_hi = h->halo_stage + (long long)(h->unpack.par * 2 + 1) * HALO_ARRAYS * h->halo_cap;
    a.base = h->unpack.base;
    a.nfa = h->geo.nfa;
    a.cap = h->halo_cap;
    if (!h->unpack.reverse) {
      a.lo_begin = ps[0];
      a.lo_cnt = ps[1] - ps[0];
      a.hi_begin = ps[nz + 1];
      a.hi_cnt = ps[nz + 2] - ps[nz + 1];
    } else {
      a.lo_begin = ps[1];
      a.lo_cnt = ps[2] - ps[1];
      a.hi_begin = ps[nz];
      a.hi_cnt = ps[nz + 1] - ps[nz];
    }
    a.masked = h->unpack.reverse;
    a.geo = h->geo;
    a.nlo = h->unpack.nlo;
    a.nhi = h->unpack.nhi;
    for (int i = 0; i < 5; ++i) {
      a.lo_list[i] = h->unpack.lo_list[i];
      a.hi_list[i] = h->unpack.hi_list[i];
    }
    const long long nmax = a.lo_cnt > a.hi_cnt ? a.lo_cnt : a.hi_cnt;
    if (nmax > 0 && a.nlo + a.nhi > 0) {
      const long long per_lo = (a.lo_cnt + 4 * BLOCK - 1) / (4 * BLOCK), per_hi = (a.hi_cnt + 4 * BLOCK - 1) / (4 * BLOCK);
      long long gx = a.nlo * per_lo + a.nhi * per_hi;
      if (gx > 8LL * h->sm_count) gx = 8LL * h->sm_count;
      if (gx > 0) {
        halo_unpack_kernel<<<(unsigned)gx, BLOCK, 0, h->st>>>(a);
        h->launches += 1;
      }
    }
    h->unpack.on = false;
  }
  return LBG_OK;
}

// after a batch has been synchronised: did a halo wait give up?
int check_p2p(lbg_handle h) {
  if (!h->p2p) return LBG_OK;
  int e = 0;
  CK(cudaMemcpy(&e, h->p2p_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) return fail(h, LBG_ERR_NCCL, "timed out waiting for a neighbour's halo planes (peer-to-peer exchange)");
  return LBG_OK;
}

const int UP_L[5] = {5, 11, 12, 15, 16};     // cz = +1  (reference l = 6,12,13,16,17)
const int DOWN_L[5] = {6, 13, 14, 17, 18};   // cz = -1  (reference l = 7,14,15,18,19)

// ---- scalar all-reduce over peer memory ------------------------------------------------------------------
// The per-step scalars (l2err + negative flag as a 2-word MAX, vacf[3] as a SUM, counts) are a few words; what
// an NCCL all-reduce costs there is latency: a stream hand-off, NCCL's kernel, the hand-off back, with the GPU
// idle in between.  With the peers' mailboxes mapped, ONE small kernel on the compute stream does it: lane r
// stores this rank's words and then the sequence number into rank r's mailbox (NVLink stores), waits until
// rank r's words of the same sequence number have landed in its own mailbox, and lane 0 combines the nranks
// contributions in rank order -- the same order on every rank, so every rank gets the same bits (max is exact
// anyway; the vacf sum is deterministic and identical everywhere, which the common stop decision needs).
// Two parities: a rank can be at most one all-reduce ahead of a peer (it needs that peer's contribution to
// finish the current one), so slot (seq & 1) of all-reduce seq-2 has been consumed before seq overwrites it.
constexpr int MAIL_OFF = 8;     // 64-bit words reserved in front of the mailbox (halo arrival flags)
constexpr int MAIL_WORDS = 8;   // {sequence number, up to 7 payload words}
enum ArOp { AR_U64_MAX = 0, AR_U64_SUM = 1, AR_F64_SUM = 2 };

__global__ void p2p_allreduce_kernel(unsigned long long* const* peer_mail, int nranks, int rank,
                                     unsigned long long seq, unsigned long long* buf, int n, int op, Ctrl* ctrl,
                                     int* err, long long max_clocks) {
  const int par = (int)(seq & 1ull);
  for (int r = threadIdx.x; r < nranks; r += blockDim.x) {
    volatile unsigned long long* slot = peer_mail[r] + MAIL_OFF + ((size_t)par * nranks + rank) * MAIL_WORDS;
    for (int i = 0; i < n; ++i) slot[1 + i] = buf[i];
    __threadfence_system();
    slot[0] = seq;
  }
  __threadfence_system();
  const long long t0 = clock64();
  for (int r = threadIdx.x; r < nranks; r += blockDim.x) {
    const volatile unsigned long long* slot = peer_mail[rank] + MAIL_OFF + ((size_t)par * nranks + r) * MAIL_WORDS;
    while (slot[0] < seq) {
      if (clock64() - t0 > max_clocks) {  // a peer died; give up instead of hanging the GPU
        ctrl->stop = 1;
        *err = 3;
        break;
      }
      __nanosleep(100);
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const volatile unsigned long long* base = peer_mail[rank] + MAIL_OFF + (size_t)par * nranks * MAIL_WORDS;
    for (int i = 0; i < n; ++i) {
      unsigned long long acc = base[1 + i];
      for (int r = 1; r < nranks; ++r) {
        const unsigned long long v = base[(size_t)r * MAIL_WORDS + 1 + i];
        if (op == AR_U64_MAX) acc = v > acc ? v : acc;
        else if (op == AR_U64_SUM) acc += v;
        else acc = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)acc) + __longlong_as_double((long long)v));
      }
      buf[i] = acc;
    }
  }
}

// all-reduce in place across the slabs, ordered after st; the result is visible to st after wait_halo.
// Peer-to-peer mode: one kernel on the compute stream (above).  NCCL mode (LBG_HALO=nccl): ncclAllReduce on the
// communication stream.
int allreduce(lbg_handle h, void* buf, size_t n, ArOp op) {
  if (h->nranks == 1) return LBG_OK;
  if (h->p2p) {
    if (n > MAIL_WORDS - 1) return fail(h, LBG_ERR_INVALID_ARG, "allreduce: too many words");
    ++h->ar_seq;
    p2p_allreduce_kernel<<<1, 32, 0, h->st>>>(h->d_peer_mail, h->nranks, h->rank, h->ar_seq, (unsigned long long*)buf,
                                              (int)n, (int)op, h->ctrl, h->p2p_err, h->spin_clocks);
    h->launches += 1;
    return LBG_OK;
  }
  cudaStream_t sa = h->st_comm;
  CK(cudaEventRecord(h->ev_ready, h->st));
  CK(cudaStreamWaitEvent(sa, h->ev_ready, 0));
  NK(g_nccl.AllReduce(buf, buf, n, op == AR_F64_SUM ? ncclDouble : ncclUint64, op == AR_U64_MAX ? ncclMax : ncclSum, h->comm, sa));
  CK(cudaEventRecord(h->ev_halo, sa));
  h->halo_pending = true;
  return LBG_OK;
}

// With one-sided pushes a neighbour may write into my halo ranges as soon as IT is ready.  After a
// (re)initialisation that clears the buffers, nobody may start pushing before every rank is done
// clearing: a barrier across the ring (an all-reduce that every rank reaches after its own memsets).
int ring_barrier(lbg_handle h) {
  if (h->nranks == 1) return LBG_OK;
  CK(cudaMemsetAsync(h->counts, 0, sizeof(unsigned long long), h->st));
  RET(allreduce(h, h->counts, 1, AR_U64_SUM));
  RET(wait_halo(h));
  CK(cudaStreamSynchronize(h->st));
  RET(check_p2p(h));
  return LBG_OK;
}

__global__ void plane_starts_kernel(Geo geo, long long ndense, long long total, long long* out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > geo.nzl + 2) return;
  const long long g = (long long)p * geo.plane;
  if (g >= ndense) {
    out[p] = total;
    return;
  }
  const uint2 w = geo.words[g >> 5];
  const unsigned bit = (unsigned)(g & 31);
  out[p] = (long long)w.y + __popc(w.x & ((1u << bit) - 1u));
}

// nature_halo == NULL: the geometry is built on the device from `label` (lbg_create_geometry)
int create_common(lbg_handle* out, int lx, int ly, int lz_global, int k0, int nzl, const int8_t* nature_halo,
                  int device, bool zwrap, int label = 0) {
  lbg_handle h = nullptr;
  if (!out || lx < 1 || ly < 1 || lz_global < 1 || nzl < 1 || k0 < 0 || k0 + nzl > lz_global)
    return fail(nullptr, LBG_ERR_INVALID_ARG, "lbg_create: invalid argument");
  const long long plane = (long long)lx * ly;
  const long long ndense = plane * (nzl + 2);
  if (ndense > (long long)INT_MAX - 64) return fail(nullptr, LBG_ERR_INVALID_ARG, "lbg_create: slab too large for 32-bit node index");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(nullptr, LBG_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(nullptr, LBG_ERR_INVALID_ARG, "lbg_create: bad device index");
  PhaseTimer tm("lbg_create");
  h = new lbg_handle_s();
  h->device = device;
  h->geo.lx = lx;
  h->geo.ly = ly;
  h->geo.plane = (int)plane;
  set_div_magic(h->geo);
  h->geo.nzl = nzl;
  h->geo.zwrap = zwrap ? 1 : 0;
  h->lz_global = lz_global;
  h->k0 = k0;
  h->nown = plane * nzl;
  h->k = make_consts();
  auto bail = [&](int rc) {
    std::string e = h->err;
    lbg_destroy(h);
    g_last_error = e;
    return rc;
  };
#define CKB(call)                                                                            \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
      return bail(e_ == cudaErrorMemoryAllocation ? LBG_ERR_NOMEM : LBG_ERR_CUDA);           \
    }                                                                                        \
  } while (0)
  CKB(cudaSetDevice(device));
  cudaDeviceProp prop;
  CKB(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  tm.lap("device properties");
  CKB(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  CKB(cudaStreamCreateWithFlags(&h->st_comm, cudaStreamNonBlocking));
  CKB(cudaStreamCreateWithFlags(&h->st_ar, cudaStreamNonBlocking));
  CKB(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  CKB(cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
  CKB(cudaEventCreateWithFlags(&h->ev_arb, cudaEventDisableTiming));
  CKB(cudaEventCreateWithFlags(&h->ev_ar[0], cudaEventDisableTiming));
  CKB(cudaEventCreateWithFlags(&h->ev_ar[1], cudaEventDisableTiming));
  CKB(cudaEventCreate(&h->ev_t0));
  CKB(cudaEventCreate(&h->ev_t1));
  h->grid_mp = occupancy_grid_mp(h->sm_count);
  CKB(cudaMalloc(&h->l2_slots, 2 * SLOT_CAP * sizeof(unsigned long long)));
  CKB(cudaMalloc(&h->vacf_slots, 3 * SLOT_CAP * sizeof(double)));
  CKB(cudaMalloc(&h->ctrl, sizeof(Ctrl)));
  CKB(cudaMalloc(&h->mp_err, sizeof(int)));
  CKB(cudaMalloc(&h->counts, 2 * sizeof(unsigned long long)));
  tm.lap("streams, events, small device buffers");
  CKB(cudaMallocHost(&h->h_l2, 2 * SLOT_CAP * sizeof(unsigned long long)));
  CKB(cudaMallocHost(&h->h_vacf, 3 * SLOT_CAP * sizeof(double)));
  CKB(cudaMallocHost(&h->h_ctrl, sizeof(Ctrl)));
  CKB(cudaMallocHost(&h->h_small, 64 * 1024));
  CKB(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->st));
  CKB(cudaMemsetAsync(h->counts, 0, 2 * sizeof(unsigned long long), h->st));

  // ---- numbering: nature -> fluid bits + ranks -> plane starts -> gidx
  h->nwords = (ndense + 31) / 32;
  CKB(cudaMalloc(&h->words, (size_t)h->nwords * sizeof(uint2)));
  h->geo.words = h->words;
  int8_t* nat_d = nullptr;
  CKB(cudaMalloc(&nat_d, (size_t)ndense));
  tm.lap("pinned host buffers");
  if (nature_halo && zwrap) {
    // whole lattice from lbg_create: `nature_halo` holds the lz own planes only
    CKB(cudaMemcpyAsync(nat_d + plane, nature_halo, (size_t)plane * nzl, cudaMemcpyHostToDevice, h->st));
    CKB(cudaMemcpyAsync(nat_d, nat_d + plane * nzl, (size_t)plane, cudaMemcpyDeviceToDevice, h->st));
    CKB(cudaMemcpyAsync(nat_d + plane * (nzl + 1), nat_d + plane, (size_t)plane, cudaMemcpyDeviceToDevice, h->st));
  } else if (nature_halo) {
    CKB(cudaMemcpyAsync(nat_d, nature_halo, (size_t)ndense, cudaMemcpyHostToDevice, h->st));
  } else {
    h->launches += launch_build_nature(label, lx, ly, lz_global, k0, nzl, nat_d, h->st);
  }
  h->launches += launch_build_bits((int)plane, nzl, nat_d, h->words, h->nwords, h->st);
  h->launches += launch_scan_ranks(h->words, h->nwords, h->counts, h->st);
  unsigned long long total = 0;
  CKB(cudaMemcpyAsync(&total, h->counts, sizeof(total), cudaMemcpyDeviceToHost, h->st));
  CKB(cudaStreamSynchronize(h->st));
  CKB(cudaGetLastError());
  cudaFree(nat_d);
  tm.lap("nature H2D + bits + ranks");
  h->nf = (long long)total;
  h->geo.nfa = (h->nf + 31) / 32 * 32;
  if (h->geo.nfa == 0) h->geo.nfa = 32;
  {
    long long* d_ps = nullptr;
    CKB(cudaMalloc(&d_ps, (size_t)(nzl + 3) * sizeof(long long)));
    plane_starts_kernel<<<(nzl + 3 + 127) / 128, 128, 0, h->st>>>(h->geo, ndense, h->nf, d_ps);
    h->launches += 1;
    h->pstart.resize(nzl + 3);
    CKB(cudaMemcpyAsync(h->pstart.data(), d_ps, (size_t)(nzl + 3) * sizeof(long long), cudaMemcpyDeviceToHost, h->st));
    CKB(cudaStreamSynchronize(h->st));
    cudaFree(d_ps);
  }
  CKB(cudaMalloc(&h->gidx, (size_t)h->geo.nfa * sizeof(uint32_t)));
  CKB(cudaMemsetAsync(h->gidx, 0, (size_t)h->geo.nfa * sizeof(uint32_t), h->st));
  h->geo.gidx = h->gidx;
  h->launches += launch_build_gidx(h->geo, h->nwords, h->gidx, h->st);
  CKB(cudaMemsetAsync(h->counts, 0, 2 * sizeof(unsigned long long), h->st));
  h->launches += launch_count_interfacial(h->geo, own_begin(h), own_end(h), h->counts, h->st);
  unsigned long long nif = 0;
  CKB(cudaMemcpyAsync(&nif, h->counts, sizeof(nif), cudaMemcpyDeviceToHost, h->st));
  CKB(cudaStreamSynchronize(h->st));
  CKB(cudaGetLastError());
  h->n_fluid = (int64_t)(own_end(h) - own_begin(h));
  h->n_if_fluid = (int64_t)nif;
  tm.lap("plane starts + gidx + counts");

  // ---- fields
  const size_t nb = (size_t)h->geo.nfa * sizeof(double);
  // f[0] and the moments share one allocation: the ring neighbours map it once (CUDA IPC) and can then push
  // halo planes of the populations AND of density / momentum (mp_init) straight into it.
  // f[1] is allocated when a second lattice is first needed.
  CKB(cudaMalloc(&h->f[0], (size_t)F0_ARRAYS * nb));
  h->mom = h->f[0] + 19 * h->geo.nfa;
  CKB(cudaMalloc(&h->jpp[0], 3 * nb));
  CKB(cudaMalloc(&h->jpp[1], 3 * nb));
  // one vacf partial per CTA of a propagate launch: at most one per tile of the slab
  CKB(cudaMalloc(&h->partial, 3 * ((size_t)(h->geo.nfa / BLOCK) + 2 + (size_t)h->grid_mp + 8) * sizeof(double)));
  if (const char* e = std::getenv("LBG_MP_TPC")) h->mp_tpc = std::atoi(e) > 0 ? std::atoi(e) : 0;
  if (const char* e = std::getenv("LBG_MP_ARITH")) h->mp_arith = std::atoi(e) ? 1 : 0;
  // Variant of the Phase-A step kernel (lb_kernels.cu), measured per workload in profiles/ab_r5a.txt:
  //   large slabs: plain kernel, 3 CTAs per SM (80 registers), tiles handed out two at a time by an atomic
  //                counter (cfg5w: 5.44 ms vs 5.70 ms for the static 2-CTA variant);
  //   small slabs (launches of 70-350 us, tail dominated): two-stage software pipeline, 2 CTAs per SM, static
  //                tile-stride schedule (cfg3: 0.330 vs 0.344 ms; cfg2: 0.068 vs 0.071 ms).
  // LBG_LB_PIPE / LBG_LB_TPC / LBG_LB_MINB override for tuning runs.
  const bool small = h->n_fluid < (8LL << 20);
  h->lb_pipe = small ? 1 : 0;
  h->lb_tpc = small ? 0 : 2;
  h->lb_minb = small ? 2 : 3;
  if (const char* e = std::getenv("LBG_LB_PIPE")) h->lb_pipe = std::atoi(e) ? 1 : 0;
  if (const char* e = std::getenv("LBG_LB_TPC")) h->lb_tpc = std::atoi(e) > 0 ? std::atoi(e) : 0;
  if (const char* e = std::getenv("LBG_LB_MINB")) h->lb_minb = std::atoi(e) >= 3 ? 3 : 2;
  h->geo.tpc = h->lb_tpc;
  h->grid_lb = occupancy_grid_lb(h->sm_count, h->lb_minb);
  h->grid_aa = occupancy_grid_aa(h->sm_count);
  // the small per-32-fid index arrays of Phase B (12 bytes per 32 nodes): allocated here so that lbg_mp_init, which
  // may run beside a queued multi-GB read-back, allocates nothing
  CKB(cudaMalloc(&h->awords, (size_t)(h->geo.nfa >> 5) * sizeof(uint2)));
  CKB(cudaMalloc(&h->rwords, (size_t)(h->geo.nfa >> 5) * sizeof(uint32_t)));
  tm.lap("field allocations");
#undef CKB
  *out = h;
  return LBG_OK;
}

// Cut planes [p_lo, p_hi] into strips of rows such that one (strip, plane) segment streams about 16 MB;
// returns nseg == 0 when a plane is small enough for plain order to keep its neighbours in L2 anyway.
// aligned: a segment's tiles start on the 32-fid boundary below its first fid (Phase A kernels)
int build_strips(lbg_handle h, int p_lo, int p_hi, SegTable* t, const char* env = "LBG_MP_STRIP_ROWS", bool aligned = false) {
  t->release();
  const Geo& g = h->geo;
  const int np = p_hi - p_lo + 1;
  if (np < 3) return LBG_OK;
  // Measured (profiles/variants_r1q.txt): plain fid order is faster for the propagate kernel on every workload
  // so far -- the three planes already survive in the 126 MB L2 -- so strips are opt-in there
  // (LBG_MP_STRIP_ROWS=<rows>).
  double rows_d = 0;
  if (aligned && h->n_fluid > 0 && (double)h->n_if_fluid >= 0.2 * (double)h->n_fluid) {
    // Phase A on a lattice with many walls: a (strip, plane) segment of about 48 MB of traffic keeps the bounce-back
    // sectors in L2 until the neighbouring plane's segment reads them (profiles/ab_r5i.txt: -1 % on the porous
    // benchmark lattice, -5.6 % on Bernoulli noise; +2 % on an all-fluid slit, where it stays off)
    const double row_fluid = (double)h->n_fluid / ((double)g.nzl * g.ly);
    rows_d = std::floor(48e6 / (352.0 * (row_fluid > 1 ? row_fluid : 1)) / 8.0) * 8.0;
    if (rows_d < 16) rows_d = 16;
  }
  if (const char* e = std::getenv(env)) rows_d = std::atof(e);
  if (rows_d <= 0 || rows_d >= g.ly) return LBG_OK;
  if (rows_d < 2) rows_d = 2;
  const int rows = (int)rows_d;
  const int nstrips = (g.ly + rows - 1) / rows;
  if (nstrips < 2) return LBG_OK;
  std::vector<long long> idx((size_t)(nstrips + 1) * np), fids(idx.size());
  for (int s = 0; s <= nstrips; ++s)
    for (int p = 0; p < np; ++p) {
      const long long y = (long long)s * rows < g.ly ? (long long)s * rows : g.ly;
      idx[(size_t)s * np + p] = (long long)(p_lo + p) * g.plane + y * g.lx;
    }
  long long *d_idx = nullptr, *d_out = nullptr;
  CK(cudaMalloc(&d_idx, idx.size() * sizeof(long long)));
  CK(cudaMalloc(&d_out, idx.size() * sizeof(long long)));
  CK(cudaMemcpyAsync(d_idx, idx.data(), idx.size() * sizeof(long long), cudaMemcpyHostToDevice, h->st));
  h->launches += launch_rank_at(g, d_idx, (int)idx.size(), (long long)g.plane * (g.nzl + 2), h->nf, d_out, h->st);
  CK(cudaMemcpyAsync(fids.data(), d_out, idx.size() * sizeof(long long), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  cudaFree(d_idx);
  cudaFree(d_out);
  std::vector<int> cum(1, 0);
  std::vector<long long> sb, se;
  for (int s = 0; s < nstrips; ++s)
    for (int p = 0; p < np; ++p) {
      const long long b = fids[(size_t)s * np + p], e = fids[(size_t)(s + 1) * np + p];
      if (e <= b) continue;
      sb.push_back(b);
      se.push_back(e);
      cum.push_back(cum.back() + (int)((e - (aligned ? tile_base(b) : b) + BLOCK - 1) / BLOCK));
    }
  if (sb.empty()) return LBG_OK;
  CK(cudaMalloc(&t->tile_cum, cum.size() * sizeof(int)));
  CK(cudaMalloc(&t->seg_begin, sb.size() * sizeof(long long)));
  CK(cudaMalloc(&t->seg_end, se.size() * sizeof(long long)));
  CK(cudaMemcpyAsync(t->tile_cum, cum.data(), cum.size() * sizeof(int), cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(t->seg_begin, sb.data(), sb.size() * sizeof(long long), cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(t->seg_end, se.data(), se.size() * sizeof(long long), cudaMemcpyHostToDevice, h->st));
  CK(cudaStreamSynchronize(h->st));
  t->nseg = (int)sb.size();
  t->ntiles = cum.back();
  return LBG_OK;
}

int ensure_second_lattice(lbg_handle h, bool zero = true) {
  if (!h->f[1]) {
    CK(cudaMalloc(&h->f[1], 19 * (size_t)h->geo.nfa * sizeof(double)));
    if (zero) CK(cudaMemsetAsync(h->f[1], 0, 19 * (size_t)h->geo.nfa * sizeof(double), h->st));
  }
  return LBG_OK;
}

int ensure_stage(lbg_handle h) {
  if (!h->stage) CK(cudaMalloc(&h->stage, (size_t)STAGE_SLOTS * (size_t)h->nown * sizeof(double)));
  return LBG_OK;
}

// which force was in effect for the last completed step (enters j(t) as f/2)
const Force& force_of_last_step(const lbg_handle h) { return h->prev_equals_cur ? h->fcur : h->fprev; }

int ensure_field(lbg_handle h, Force& f) {
  if (!f.field) CK(cudaMalloc(&f.field, 3 * (size_t)h->geo.nfa * sizeof(double)));
  return LBG_OK;
}

// make `f` usable as a field (materialise a uniform force on the fluid nodes)
int as_field(lbg_handle h, Force& f, double** scratch, const double** out) {
  if (f.mode == FORCE_FIELD) {
    *out = f.field;
    return LBG_OK;
  }
  if (!*scratch) CK(cudaMalloc(scratch, 3 * (size_t)h->geo.nfa * sizeof(double)));
  const double zero[3] = {0, 0, 0};
  h->launches += launch_fill_force(h->geo, h->nf, f.mode == FORCE_NONE ? zero : f.u, *scratch, h->st);
  *out = *scratch;
  return LBG_OK;
}

struct ForceSel {
  int mode;
  double fj[3], fc[3];
  const double *fj_field, *fc_field;
};

int select_force(lbg_handle h, Force& fj, Force& fc, ForceSel* s, double** scratch_j, double** scratch_c) {
  std::memset(s, 0, sizeof(*s));
  if (fj.mode != FORCE_FIELD && fc.mode != FORCE_FIELD) {
    s->mode = (fj.mode == FORCE_NONE && fc.mode == FORCE_NONE) ? FORCE_NONE : FORCE_UNIFORM;
    for (int d = 0; d < 3; ++d) {
      s->fj[d] = fj.mode == FORCE_NONE ? 0.0 : fj.u[d];
      s->fc[d] = fc.mode == FORCE_NONE ? 0.0 : fc.u[d];
    }
    return LBG_OK;
  }
  s->mode = FORCE_FIELD;
  RET(as_field(h, fj, scratch_j, &s->fj_field));
  if (&fj == &fc) s->fc_field = s->fj_field;
  else RET(as_field(h, fc, scratch_c, &s->fc_field));
  return LBG_OK;
}

struct StepFlags {
  bool check, writej;
  int batch_idx, prev_checked, prev_may_stop;
  double target;
};

// Enqueue kernel K: f[fin] -> f[1-fin] over own planes (+ halo exchange of the result).
int enqueue_lb_kernel(lbg_handle h, int fin, double tau, const ForceSel& fs, int jold, const StepFlags& fl) {
  LBArgs a{};
  a.geo = h->geo;
  a.k = h->k;
  a.fin = h->f[fin];
  a.fout = h->f[1 - fin];
  a.w1 = 1.0 - 1.0 / tau;
  a.w2 = 1.0 / tau;
  a.w3 = 1.0 - 1.0 / (2.0 * tau);
  for (int d = 0; d < 3; ++d) {
    a.fj[d] = fs.fj[d];
    a.fc[d] = fs.fc[d];
  }
  a.fj_field = fs.fj_field;
  a.fc_field = fs.fc_field;
  a.jold = h->jpp[jold];
  a.jnew = h->jpp[1 - jold];
  a.l2_slots = h->l2_slots;
  a.batch_idx = fl.batch_idx;
  a.prev_checked = fl.prev_checked;
  a.prev_may_stop = fl.prev_may_stop;
  a.target = fl.target;
  a.ctrl = h->ctrl;
  a.pipe = h->lb_pipe;
  a.neg_flag_local = (h->nranks > 1 && !fl.prev_checked) ? 1 : 0;
  const bool tau1 = (tau == 1.0);
  const std::vector<long long>& ps = h->pstart;
  const int nz = h->geo.nzl;
  RET(wait_halo(h));
  auto launch = [&](long long b, long long e, bool use_strips = false) {
    a.fid_begin = b;
    a.fid_end = e;
    a.nseg = 0;
    if (use_strips && h->lb_strips.nseg > 0 && !h->lb_pipe) {
      a.nseg = h->lb_strips.nseg;
      a.ntiles = h->lb_strips.ntiles;
      a.tile_cum = h->lb_strips.tile_cum;
      a.seg_begin = h->lb_strips.seg_begin;
      a.seg_end = h->lb_strips.seg_end;
    }
    h->launches += launch_lb_step(a, tau1, fs.mode, fl.check, fl.writej, h->lb_minb, h->grid_lb, h->st);
  };
  if (!h->lb_strips_built) {  // strip order over the planes the big launch covers (opt-in, LBG_LB_STRIP_ROWS)
    h->lb_strips_built = true;
    if (h->nranks == 1) RET(build_strips(h, 1, nz, &h->lb_strips, "LBG_LB_STRIP_ROWS", true));
    else RET(build_strips(h, 2, nz - 1, &h->lb_strips, "LBG_LB_STRIP_ROWS", true));
  }
  if (h->nranks == 1) {
    launch(own_begin(h), own_end(h), true);
  } else {
    // boundary planes first, so their populations can travel while the interior runs
    launch(ps[1], ps[2]);
    if (nz > 1) launch(ps[nz], ps[nz + 1]);
    RET(halo_exchange(h, h->f[1 - fin], UP_L, 5, DOWN_L, 5));
    if (nz > 2) launch(ps[2], ps[nz], true);
    if (fl.check) {
      // global max of l2err and of the negative-population flag (one all-reduce of the slot pair),
      // before the next step looks at them
      RET(wait_halo(h));
      RET(allreduce(h, h->l2_slots + 2 * fl.batch_idx, 2, AR_U64_MAX));
    }
  }
  return LBG_OK;
}

// Make f[1-src] = n*(t+1) valid for (tau, current force); optionally make j(t) available in jpp[jc].
int ensure_collided(lbg_handle h, double tau, bool need_j) {
  const bool stale = !h->collided_ok || tau != h->collided_tau || h->collided_force_version != h->force_version;
  const bool j_missing = need_j && h->j_valid_step != h->t;
  if (!stale && !j_missing) return LBG_OK;
  double *scr_j = nullptr, *scr_c = nullptr;
  int rc = LBG_OK;
  if (h->precollision) {
    ForceSel fs;
    rc = select_force(h, h->fcur, h->fcur, &fs, &scr_c, &scr_c);
    if (rc == LBG_OK) {
      CollideArgs a{};
      a.geo = h->geo;
      a.k = h->k;
      a.fin = h->f[h->src];
      a.fout = h->f[1 - h->src];
      a.mom = h->mom;
      a.fid_begin = own_begin(h);
      a.fid_end = own_end(h);
      a.w1 = 1.0 - 1.0 / tau;
      a.w2 = 1.0 / tau;
      a.w3 = 1.0 - 1.0 / (2.0 * tau);
      for (int d = 0; d < 3; ++d) a.fc[d] = fs.fc[d];
      a.fc_field = fs.fc_field;
      rc = wait_halo(h);
      if (rc == LBG_OK) {
        h->launches += launch_collide(a, tau == 1.0, fs.mode, h->grid_lb, h->st);
        rc = halo_exchange(h, h->f[1 - h->src], UP_L, 5, DOWN_L, 5);
      }
      if (rc == LBG_OK && j_missing) {
        const size_t nb = (size_t)h->geo.nfa * sizeof(double);
        cudaError_t e = cudaMemcpyAsync(h->jpp[h->jc], h->mom + h->geo.nfa, 3 * nb, cudaMemcpyDeviceToDevice, h->st);
        if (e != cudaSuccess) rc = fail(h, LBG_ERR_CUDA, cudaGetErrorString(e));
        h->j_valid_step = h->t;
      }
    }
  } else {
    // redo K(t): pull n*(t) from f[src], force of step t for j(t), current force for the collision
    Force& fj = h->prev_equals_cur ? h->fcur : h->fprev;
    ForceSel fs;
    rc = select_force(h, fj, h->fcur, &fs, &scr_j, &scr_c);
    if (rc == LBG_OK) {
      StepFlags fl{};
      fl.check = false;
      fl.writej = j_missing;
      // K(t) writes j(t) into jpp[1-jold]; keep jc pointing at j(t)
      rc = enqueue_lb_kernel(h, h->src, tau, fs, 1 - h->jc, fl);
      if (rc == LBG_OK && j_missing) h->j_valid_step = h->t;
    }
  }
  if (scr_j || scr_c) {
    cudaStreamSynchronize(h->st);
    cudaFree(scr_j);
    cudaFree(scr_c);
  }
  if (rc != LBG_OK) return rc;
  h->collided_ok = true;
  h->collided_tau = tau;
  h->collided_force_version = h->force_version;
  return LBG_OK;
}

int refresh_moments(lbg_handle h, double* pops_out) {
  if (h->phase != PH_LB) return fail(h, LBG_ERR_STATE, "no Lattice-Boltzmann state (call lbg_lb_init first)");
  if (h->precollision) return LBG_OK;  // mom / f[src] hold the state the driver set
  if (h->mom_valid_step == h->t && !pops_out) return LBG_OK;
  const Force& fj = force_of_last_step(h);
  if (h->in_place) {
    LBArgs a{};
    a.geo = h->geo;
    a.k = h->k;
    a.fin = h->f[0];
    a.fout = h->f[0];
    a.fid_begin = own_begin(h);
    a.fid_end = own_end(h);
    for (int d = 0; d < 3; ++d) a.fj[d] = fj.mode == FORCE_NONE ? 0.0 : fj.u[d];
    a.fj_field = fj.field;
    a.ctrl = h->ctrl;
    RET(wait_halo(h));
    CK(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->st));
    h->launches += launch_aa_moments(a, fj.mode, h->aa_swapped, false, false, h->mom, pops_out, h->grid_aa, h->st);
    h->mom_valid_step = h->t;
    return LBG_OK;
  }
  MomArgs a{};
  a.geo = h->geo;
  a.fin = h->f[h->src];
  a.mom = h->mom;
  a.pops = pops_out;
  a.fid_begin = own_begin(h);
  a.fid_end = own_end(h);
  for (int d = 0; d < 3; ++d) a.fj[d] = fj.mode == FORCE_NONE ? 0.0 : fj.u[d];
  a.fj_field = fj.field;
  RET(wait_halo(h));
  h->launches += launch_moments(a, fj.mode, h->grid_lb, h->st);
  h->mom_valid_step = h->t;
  return LBG_OK;
}

int snapshot_prev_force(lbg_handle h) {
  if (!h->prev_equals_cur) return LBG_OK;
  h->fprev.mode = h->fcur.mode;
  for (int d = 0; d < 3; ++d) h->fprev.u[d] = h->fcur.u[d];
  if (h->fcur.mode == FORCE_FIELD) {
    RET(ensure_field(h, h->fprev));
    CK(cudaMemcpyAsync(h->fprev.field, h->fcur.field, 3 * (size_t)h->geo.nfa * sizeof(double),
                       cudaMemcpyDeviceToDevice, h->st));
  }
  h->prev_equals_cur = false;
  return LBG_OK;
}

// compact arrays -> the driver's dense (i,j,k) arrays over the own planes (0 on solid nodes).
// The staging buffer holds three dense arrays: the scatter kernel of array i+1 (compute stream) runs while the
// copy engine moves array i to the host (communication stream).  pull_l >= 0: src[i] is ignored and array i is
// n(t)(., pull_l + i) rebuilt from the post-collision populations `pull_from` (launch_pull_to_dense).
// Small device -> host reads that may run while a multi-GB read-back occupies the copy engine (a cudaMemcpy of a
// few bytes would queue behind it for tens of ms): a one-block kernel stores the bytes into page-locked host memory.
__global__ void to_host_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int nbytes) {
  for (int i = threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}

constexpr size_t SMALL_CAP = 64 * 1024;

// queue a small read into the pinned scratch at byte offset `off`; valid after the next synchronisation of h->st
int read_small_async(lbg_handle h, size_t off, const void* dev_src, size_t bytes) {
  if (off + bytes > SMALL_CAP) return fail(h, LBG_ERR_INVALID_ARG, "read_small: too large");
  to_host_kernel<<<1, 256, 0, h->st>>>((const unsigned char*)dev_src, h->h_small + off, (int)bytes);
  h->launches += 1;
  return LBG_OK;
}

int wait_transfers(lbg_handle h) {
  if (h->transfers_pending) {
    CK(cudaStreamSynchronize(h->st_ar));
    h->transfers_pending = false;
  }
  return LBG_OK;
}

// async: return once everything is enqueued; the host arrays are valid after lbg_wait_transfers.  The copies run on
// their own stream (not the halo stream: a multi-GB read-back must not sit in front of the next halo planes).
int copy_own_to_host_many(lbg_handle h, double* const* dst, const double* const* src, int count,
                          const double* pull_from = nullptr, bool async = false) {
  RET(wait_transfers(h));
  RET(ensure_stage(h));
  for (int e = 0; e < 2 * STAGE_SLOTS; ++e)
    if (!h->ev_stage[e]) CK(cudaEventCreateWithFlags(&h->ev_stage[e], cudaEventDisableTiming));
  int issued = 0;
  for (int i = 0; i < count; ++i) {
    if (!dst[i]) continue;
    const int slot = issued % STAGE_SLOTS;
    double* stg = h->stage + (size_t)slot * h->nown;
    if (issued >= STAGE_SLOTS) CK(cudaStreamWaitEvent(h->st, h->ev_stage[STAGE_SLOTS + slot], 0));  // the slot's previous copy is done
    if (pull_from) h->launches += launch_pull_to_dense(h->geo, pull_from, i, stg, h->st);
    else h->launches += launch_scatter_to_dense(h->geo, src[i], stg, h->st);
    CK(cudaEventRecord(h->ev_stage[slot], h->st));
    CK(cudaStreamWaitEvent(h->st_ar, h->ev_stage[slot], 0));
    CK(cudaMemcpyAsync(dst[i], stg, (size_t)h->nown * sizeof(double), cudaMemcpyDeviceToHost, h->st_ar));
    CK(cudaEventRecord(h->ev_stage[STAGE_SLOTS + slot], h->st_ar));
    ++issued;
  }
  if (async) {
    h->transfers_pending = issued > 0;
    return LBG_OK;
  }
  CK(cudaStreamSynchronize(h->st_ar));
  CK(cudaStreamSynchronize(h->st));
  return LBG_OK;
}

int copy_own_to_device(lbg_handle h, double* arr, const double* src) {
  RET(wait_transfers(h));
  RET(ensure_stage(h));
  CK(cudaMemcpyAsync(h->stage, src, (size_t)h->nown * sizeof(double), cudaMemcpyHostToDevice, h->st));
  h->launches += launch_gather_from_dense(h->geo, h->stage, arr, h->st);
  CK(cudaStreamSynchronize(h->st));
  return LBG_OK;
}

}  // namespace

// ===========================================================================
extern "C" {

int lbg_abi_version(void) { return LBG_ABI_VERSION; }

const char* lbg_status_string(int s) {
  switch (s) {
    case LBG_OK: return "ok";
    case LBG_ERR_NEGATIVE_POPULATION: return "In equilibration, the population n(x,y,z,vel) < 0";
    case LBG_ERR_RESTPART_NEGATIVE: return "somewhere restpart is negative";
    case LBG_ERR_RELAXATION_TIME: return "relaxation_time must be > 0.5";
    case LBG_ERR_TRACER_DB: return "The diffusion coefficient (tracer_Db in input file) is invalid";
    case LBG_ERR_TRACER_KA_KD: return "I detected tracer%ka or tracer%kd to be <0 in module moment_propagation";
    case LBG_ERR_ALL_SOLID: return "All nodes are solid: no fluid, no fluid dynamics!";
    case LBG_ERR_INVALID_ARG: return "invalid argument";
    case LBG_ERR_STATE: return "call order violated";
    case LBG_ERR_UNSUPPORTED: return "unsupported option";
    case LBG_ERR_NO_DEVICE: return "no CUDA device (there is no CPU path)";
    case LBG_ERR_CUDA: return "CUDA error";
    case LBG_ERR_NCCL: return "NCCL error";
    case LBG_ERR_NOMEM: return "out of device memory";
    default: return "unknown status";
  }
}

const char* lbg_last_error(lbg_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int lbg_device_count(int* count) {
  if (!count) return LBG_ERR_INVALID_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
  *count = n;
  return LBG_OK;
}

int lbg_partition(int lz, int nranks, int rank, int* k0, int* nzl) {
  if (lz < 1 || nranks < 1 || rank < 0 || rank >= nranks || nranks > lz || !k0 || !nzl) return LBG_ERR_INVALID_ARG;
  const int base = lz / nranks, extra = lz % nranks;
  *nzl = base + (rank < extra ? 1 : 0);
  *k0 = rank * base + (rank < extra ? rank : extra);
  return LBG_OK;
}

int lbg_halo_plan(int up[5], int down[5]) {
  if (!up || !down) return LBG_ERR_INVALID_ARG;
  for (int i = 0; i < 5; ++i) {
    up[i] = UP_L[i];
    down[i] = DOWN_L[i];
  }
  return LBG_OK;
}

int lbg_create(lbg_handle* out, int lx, int ly, int lz, const int8_t* nature, int device) {
  if (!nature || lx < 1 || ly < 1 || lz < 1) return fail(nullptr, LBG_ERR_INVALID_ARG, "lbg_create: invalid argument");
  const size_t plane = (size_t)lx * ly;
  if (!std::memchr(nature, 0, plane * lz)) return fail(nullptr, LBG_ERR_ALL_SOLID, lbg_status_string(LBG_ERR_ALL_SOLID));
  // the periodic halo planes (k = -1 == lz-1, k = lz == 0) are copied on the device side of create_common
  return create_common(out, lx, ly, lz, 0, lz, nature, device, true);
}

int lbg_create_slab(lbg_handle* out, int lx, int ly, int lz_global, int k0, int nzl, const int8_t* nature_halo,
                    int device) {
  if (!nature_halo) return fail(nullptr, LBG_ERR_INVALID_ARG, "lbg_create_slab: nature is NULL");
  return create_common(out, lx, ly, lz_global, k0, nzl, nature_halo, device, false);
}

int lbg_create_geometry(lbg_handle* out, int label, int lx, int ly, int lz_global, int k0, int nzl, int device) {
  if (label != -1 && label != 1 && label != 2 && label != 3)
    return fail(nullptr, LBG_ERR_UNSUPPORTED, "lbg_create_geometry builds geometryLabel -1, 1, 2 and 3 only");
  if (label == 2 && (lx != ly || lx < 3)) return fail(nullptr, LBG_ERR_INVALID_ARG, "wall=2 is for cylinders, which should have same lx and ly (>= 3)");
  if (label == 3 && (lx != ly || lx != lz_global)) return fail(nullptr, LBG_ERR_INVALID_ARG, "with wall = 3, i.e. cfc cell, the supercell should be cubic with lx=ly=lz");
  if (label == 1 && lz_global < 3 && lx * ly > 0) {
    // two walls and nothing between them: supercell_definition.f90:84-86
    return fail(nullptr, LBG_ERR_ALL_SOLID, lbg_status_string(LBG_ERR_ALL_SOLID));
  }
  const bool whole = (k0 == 0 && nzl == lz_global);
  return create_common(out, lx, ly, lz_global, k0, nzl, nullptr, device, whole, label);
}

int lbg_get_nature(lbg_handle h, int8_t* out) {
  if (!h || !out) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(wait_transfers(h));
  RET(ensure_stage(h));   // scratch: no allocation per call
  int8_t* d = reinterpret_cast<int8_t*>(h->stage);
  h->launches += launch_dense_nature(h->geo, d, h->st);
  CK(cudaMemcpyAsync(out, d, (size_t)h->nown, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return LBG_OK;
}

int lbg_destroy(lbg_handle h) {
  if (!h) return LBG_OK;
  cudaSetDevice(h->device);
  h->transfers_pending = false;
  // an exchange whose planes have not been consumed yet (e.g. lbg_mp_init directly followed by lbg_destroy): the
  // neighbours' pushes into my receive buffer must have landed before it is freed (bounded wait, errors ignored)
  if (h->p2p && (h->unpack.on || h->xwait)) wait_halo(h);
  if (h->st) cudaStreamSynchronize(h->st);
  if (h->st_comm) cudaStreamSynchronize(h->st_comm);
  if (h->st_ar) cudaStreamSynchronize(h->st_ar);
  for (int s = 0; s < 2; ++s)
    if (h->peer_ipc[s]) cudaIpcCloseMemHandle(h->peer_stage[s]);
  cudaFree(h->halo_stage);
  for (size_t r = 0; r < h->peer_mail.size(); ++r)
    if (h->peer_mail_ipc[r]) cudaIpcCloseMemHandle(h->peer_mail[r]);
  cudaFree(h->mail);
  cudaFree(h->d_peer_mail);
  cudaFree(h->p2p_err);
  if (h->comm) {  // the communicator stays in the process-wide cache for the next handle
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (h->comm_slot >= 0 && (size_t)h->comm_slot < g_comms.size()) g_comms[(size_t)h->comm_slot].in_use = false;
  }
  cudaFree(h->words);
  cudaFree(h->gidx);
  cudaFree(h->awords);
  cudaFree(h->rwords);
  cudaFree(h->f[0]);
  cudaFree(h->f[1]);
  cudaFree(h->jpp[0]);
  cudaFree(h->jpp[1]);
  cudaFree(h->stage);
  cudaFree(h->l2_slots);
  cudaFree(h->vacf_slots);
  cudaFree(h->partial);
  cudaFree(h->ctrl);
  cudaFree(h->mp_err);
  cudaFree(h->counts);
  cudaFree(h->fcur.field);
  cudaFree(h->fprev.field);
  h->strips.release();
  h->lb_strips.release();
  cudaFreeHost(h->h_l2);
  cudaFreeHost(h->h_vacf);
  cudaFreeHost(h->h_ctrl);
  cudaFreeHost(h->h_small);
  if (h->ev_ready) cudaEventDestroy(h->ev_ready);
  if (h->ev_halo) cudaEventDestroy(h->ev_halo);
  if (h->ev_arb) cudaEventDestroy(h->ev_arb);
  if (h->ev_ar[0]) cudaEventDestroy(h->ev_ar[0]);
  if (h->ev_ar[1]) cudaEventDestroy(h->ev_ar[1]);
  for (int e = 0; e < 8; ++e)
    if (h->ev_stage[e]) cudaEventDestroy(h->ev_stage[e]);
  if (h->ev_t0) cudaEventDestroy(h->ev_t0);
  if (h->ev_t1) cudaEventDestroy(h->ev_t1);
  if (h->st) cudaStreamDestroy(h->st);
  if (h->st_comm) cudaStreamDestroy(h->st_comm);
  if (h->st_ar) cudaStreamDestroy(h->st_ar);
  // releasing tens of GB is partly deferred by the driver: pay for it here, not in whatever CUDA call comes next
  cudaDeviceSynchronize();
  delete h;
  return LBG_OK;
}

int lbg_comm_unique_id(void* id_out) {
  lbg_handle h = nullptr;
  if (!id_out) return LBG_ERR_INVALID_ARG;
  if (!g_nccl.load()) return fail(nullptr, LBG_ERR_NCCL, "cannot load libnccl.so.2");
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == LBG_UNIQUE_ID_BYTES, "ncclUniqueId size");
  NK(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return LBG_OK;
}

int lbg_comm_init(lbg_handle h, int nranks, int rank, const void* id) {
  if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return LBG_ERR_INVALID_ARG;
  if (nranks == 1) return LBG_OK;
  if (h->comm) return fail(h, LBG_ERR_STATE, "lbg_comm_init: the handle already belongs to a communicator");
  if (h->geo.zwrap) return fail(h, LBG_ERR_STATE, "lbg_comm_init needs a handle made by lbg_create_slab");
  if (!g_nccl.load()) return fail(h, LBG_ERR_NCCL, "cannot load libnccl.so.2");
  CK(cudaSetDevice(h->device));
  PhaseTimer tm("lbg_comm_init");
  // The NCCL communicator is only the bootstrap channel (handle exchange, agreement) and the fallback transport.
  // Creating one costs seconds on 8 GPUs, so a process keeps the communicators it made and a later handle with
  // the same (nranks, rank, device) reuses a free one instead of the fresh id (every rank of the job takes the same
  // decision as long as the ranks create and destroy their handles in the same order; LBG_COMM_CACHE=0 disables).
  {
    bool use_cache = true;
    if (const char* e = std::getenv("LBG_COMM_CACHE")) use_cache = std::atoi(e) != 0;
    std::lock_guard<std::mutex> lk(g_comm_mu);
    if (use_cache)
      for (size_t i = 0; i < g_comms.size(); ++i) {
        CommEntry& c = g_comms[i];
        if (!c.in_use && c.nranks == nranks && c.rank == rank && c.device == h->device) {
          c.in_use = true;
          h->comm = c.comm;
          h->comm_slot = (int)i;
          break;
        }
      }
  }
  if (!h->comm) {
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    NK(g_nccl.CommInitRank(&h->comm, nranks, uid, rank));
    std::lock_guard<std::mutex> lk(g_comm_mu);
    g_comms.push_back(CommEntry{nranks, rank, h->device, h->comm, true});
    h->comm_slot = (int)g_comms.size() - 1;
  }
  tm.lap("communicator");
  h->nranks = nranks;
  h->rank = rank;
  if (const char* e = std::getenv("LBG_P2P_TIMEOUT_S")) {
    const double sec = std::atof(e);
    if (sec > 0) h->spin_clocks = (long long)(sec * 2.0e9);
  }
  // ---- peer-to-peer: map the ring neighbours' population buffers and every rank's mailbox (same node, NVLink)
  {
    struct PeerInfo {
      int pid, dev;
      cudaIpcMemHandle_t hs, ml;
      unsigned long long hs_ptr, ml_ptr;
      long long cap;
      int ok;
    };
    int want = 1;
    if (const char* e = std::getenv("LBG_HALO")) want = std::strcmp(e, "nccl") != 0;
    {
      const std::vector<long long>& ps = h->pstart;
      const int nzq = h->geo.nzl;
      long long m = 0;  // halo planes (forward trips) and own boundary planes (return trip of the in-place scheme)
      for (long long c : {ps[1] - ps[0], ps[nzq + 2] - ps[nzq + 1], ps[2] - ps[1], ps[nzq + 1] - ps[nzq]}) m = c > m ? c : m;
      h->halo_cap = (m + 31) / 32 * 32 + 32;
      CK(cudaMalloc(&h->halo_stage, (size_t)(2 * 2 * HALO_ARRAYS) * (size_t)h->halo_cap * sizeof(double)));
    }
    const size_t mail_words = MAIL_OFF + 2 * (size_t)nranks * MAIL_WORDS;
    CK(cudaMalloc(&h->mail, mail_words * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(h->mail, 0, mail_words * sizeof(unsigned long long), h->st));
    h->flags = reinterpret_cast<unsigned int*>(h->mail);
    CK(cudaMalloc(&h->p2p_err, sizeof(int)));
    CK(cudaMemsetAsync(h->p2p_err, 0, sizeof(int), h->st));
    tm.lap("receive buffer + mailbox");
    PeerInfo me{};
    me.pid = (int)getpid();
    me.dev = h->device;
    me.ok = want;
    if (cudaIpcGetMemHandle(&me.hs, h->halo_stage) != cudaSuccess) me.ok = 0;
    if (cudaIpcGetMemHandle(&me.ml, h->mail) != cudaSuccess) me.ok = 0;
    cudaGetLastError();
    me.hs_ptr = (unsigned long long)h->halo_stage;
    me.ml_ptr = (unsigned long long)h->mail;
    me.cap = h->halo_cap;
    std::vector<PeerInfo> all((size_t)nranks);
    PeerInfo *d_me = nullptr, *d_all = nullptr;
    CK(cudaMalloc(&d_me, sizeof(PeerInfo)));
    CK(cudaMalloc(&d_all, sizeof(PeerInfo) * (size_t)nranks));
    CK(cudaMemcpyAsync(d_me, &me, sizeof(PeerInfo), cudaMemcpyHostToDevice, h->st));
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    *(void**)(&AllGather) = dlsym(g_nccl.lib, "ncclAllGather");
    if (!AllGather) return fail(h, LBG_ERR_NCCL, "ncclAllGather not found");
    NK(AllGather(d_me, d_all, sizeof(PeerInfo), ncclChar, h->comm, h->st));
    CK(cudaMemcpyAsync(all.data(), d_all, sizeof(PeerInfo) * (size_t)nranks, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    cudaFree(d_me);
    cudaFree(d_all);
    tm.lap("handle all-gather");
    bool ok = true;
    for (const PeerInfo& p : all) ok = ok && p.ok;
    auto enable_peer = [&](int dev) {
      if (dev == h->device) return true;
      cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
      cudaGetLastError();
      return e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
    };
    // every rank's mailbox
    h->peer_mail.assign((size_t)nranks, nullptr);
    h->peer_mail_ipc.assign((size_t)nranks, false);
    for (int r = 0; r < nranks && ok; ++r) {
      const PeerInfo& p = all[(size_t)r];
      if (r == rank) {
        h->peer_mail[(size_t)r] = h->mail;
      } else if (p.pid == me.pid) {  // several slabs driven from one process: plain peer access
        ok = enable_peer(p.dev);
        h->peer_mail[(size_t)r] = (unsigned long long*)p.ml_ptr;
      } else {
        void* m = nullptr;
        if (cudaIpcOpenMemHandle(&m, p.ml, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          ok = false;
          cudaGetLastError();
        }
        h->peer_mail[(size_t)r] = (unsigned long long*)m;
        h->peer_mail_ipc[(size_t)r] = m != nullptr;
      }
    }
    tm.lap("mailboxes mapped");
    const int nbr[2] = {down_rank(h), up_rank(h)};
    for (int s = 0; s < 2 && ok; ++s) {
      const PeerInfo& p = all[(size_t)nbr[s]];
      h->peer_cap[s] = p.cap;
      h->peer_flags[s] = reinterpret_cast<unsigned int*>(h->peer_mail[(size_t)nbr[s]]);
      if (p.pid == me.pid) {
        h->peer_stage[s] = (double*)p.hs_ptr;
        h->peer_ipc[s] = false;
      } else if (s == 1 && nbr[1] == nbr[0]) {  // two slabs: the same neighbour on both sides, map once
        h->peer_stage[1] = h->peer_stage[0];
        h->peer_ipc[1] = false;
      } else {
        void* a = nullptr;
        if (cudaIpcOpenMemHandle(&a, p.hs, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          ok = false;
          cudaGetLastError();
        }
        h->peer_stage[s] = (double*)a;
        h->peer_ipc[s] = a != nullptr;
      }
    }
    tm.lap("neighbour receive buffers mapped");
    if (ok) {
      CK(cudaMalloc(&h->d_peer_mail, sizeof(unsigned long long*) * (size_t)nranks));
      CK(cudaMemcpyAsync(h->d_peer_mail, h->peer_mail.data(), sizeof(unsigned long long*) * (size_t)nranks,
                         cudaMemcpyHostToDevice, h->st));
    }
    // every rank must take the same path: agree on the outcome (this all-reduce is also the barrier that
    // orders every rank's mailbox memset before anybody's first push)
    int* d_ok = h->mp_err;
    int okv = ok ? 1 : 0;
    CK(cudaMemcpyAsync(d_ok, &okv, sizeof(int), cudaMemcpyHostToDevice, h->st));
    NK(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, h->comm, h->st));
    CK(cudaMemcpyAsync(&okv, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->p2p = okv != 0;
    tm.lap("agreement");
  }
  // The LB step kernel is persistent (one wave of resident blocks).  NCCL's send/recv kernel needs SM
  // room to run beside the interior planes' kernel, otherwise the exchange waits for that kernel to end:
  // leave 32 block slots free (measured at N=2: 7.7 -> 6.2 ms per step; profiles/multigpu_r1.txt).
  // The propagate kernel gets the same treatment: its halo exchange and the lagged vacf all-reduce run
  // beside the interior kernel.
  // With peer-to-peer halos the copy engines move the planes and the scalar all-reduces are one-warp
  // kernels between the step kernels: nothing is reserved.
  int reserve = h->p2p ? 0 : 32;
  if (const char* e = std::getenv("LBG_GRID_RESERVE")) reserve = std::atoi(e);
  if (reserve > 0 && reserve < h->grid_lb) h->grid_lb -= reserve;
  int reserve_mp = h->p2p ? 0 : 32;
  if (const char* e = std::getenv("LBG_GRID_RESERVE_MP")) reserve_mp = std::atoi(e);
  if (reserve_mp > 0 && reserve_mp < h->grid_mp) h->grid_mp -= reserve_mp;
  return LBG_OK;
}

int lbg_get_interfacial(lbg_handle h, int8_t* out) {
  if (!h || !out) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(wait_transfers(h));
  RET(ensure_stage(h));
  int8_t* d = reinterpret_cast<int8_t*>(h->stage);
  h->launches += launch_dense_interfacial(h->geo, d, h->st);
  CK(cudaMemcpyAsync(out, d, (size_t)h->nown, cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return LBG_OK;
}

int lbg_get_counts(lbg_handle h, int64_t* nf, int64_t* nif) {
  if (!h) return LBG_ERR_INVALID_ARG;
  if (nf) *nf = h->n_fluid;
  if (nif) *nif = h->n_if_fluid;
  return LBG_OK;
}

// --------------------------------------------------------------------------- Phase A
static void reset_lb_state(lbg_handle h) {
  h->phase = PH_LB;
  h->t = 0;
  h->precollision = true;
  h->src = 0;
  h->jc = 0;
  h->collided_ok = false;
  h->j_valid_step = -1;
  h->mom_valid_step = 0;
  h->fcur.mode = FORCE_NONE;
  h->fprev.mode = FORCE_NONE;
  for (int d = 0; d < 3; ++d) h->fcur.u[d] = h->fprev.u[d] = 0.0;
  h->prev_equals_cur = true;
  h->force_version++;
}

static int zero_lb_fields(lbg_handle h) {
  const size_t nb = (size_t)h->geo.nfa * sizeof(double);
  if (h->in_place) {  // memory-lean mode: one lattice
    if (h->f[1]) {
      CK(cudaStreamSynchronize(h->st));
      cudaFree(h->f[1]);
      h->f[1] = nullptr;
    }
  } else {
    RET(ensure_second_lattice(h));
    CK(cudaMemsetAsync(h->f[1], 0, 19 * nb, h->st));
  }
  h->aa_swapped = false;
  CK(cudaMemsetAsync(h->f[0], 0, 19 * nb, h->st));
  CK(cudaMemsetAsync(h->mom, 0, 4 * nb, h->st));
  CK(cudaMemsetAsync(h->jpp[0], 0, 3 * nb, h->st));
  CK(cudaMemsetAsync(h->jpp[1], 0, 3 * nb, h->st));
  return LBG_OK;
}

int lbg_lb_init(lbg_handle h, double rho0) {
  if (!h) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(zero_lb_fields(h));
  h->launches += launch_lb_init(h->geo, own_begin(h), own_end(h), rho0, h->k.a0, h->f[0], h->mom, h->st);
  CK(cudaGetLastError());
  reset_lb_state(h);
  RET(ring_barrier(h));
  return LBG_OK;
}

int lbg_lb_upload(lbg_handle h, const double* n, const double* rho, const double* jx, const double* jy,
                  const double* jz) {
  if (!h || !n || !rho || !jx || !jy || !jz) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(zero_lb_fields(h));
  for (int l = 0; l < 19; ++l) RET(copy_own_to_device(h, h->f[0] + (long long)l * h->geo.nfa, n + (size_t)l * h->nown));
  const double* m[4] = {rho, jx, jy, jz};
  for (int c = 0; c < 4; ++c) RET(copy_own_to_device(h, h->mom + (long long)c * h->geo.nfa, m[c]));
  reset_lb_state(h);
  RET(ring_barrier(h));
  return LBG_OK;
}

int lbg_lb_set_force_uniform(lbg_handle h, const double f[3]) {
  if (!h || !f) return LBG_ERR_INVALID_ARG;
  if (h->phase != PH_LB) return fail(h, LBG_ERR_STATE, "lbg_lb_set_force_uniform before lbg_lb_init");
  CK(cudaSetDevice(h->device));
  RET(snapshot_prev_force(h));
  const bool zero = (f[0] == 0.0 && f[1] == 0.0 && f[2] == 0.0);
  h->fcur.mode = zero ? FORCE_NONE : FORCE_UNIFORM;
  for (int d = 0; d < 3; ++d) h->fcur.u[d] = f[d];
  h->force_version++;
  return LBG_OK;
}

int lbg_lb_set_force_field(lbg_handle h, const double* fx, const double* fy, const double* fz) {
  if (!h || !fx || !fy || !fz) return LBG_ERR_INVALID_ARG;
  if (h->phase != PH_LB) return fail(h, LBG_ERR_STATE, "lbg_lb_set_force_field before lbg_lb_init");
  CK(cudaSetDevice(h->device));
  RET(snapshot_prev_force(h));
  RET(ensure_field(h, h->fcur));
  CK(cudaMemsetAsync(h->fcur.field, 0, 3 * (size_t)h->geo.nfa * sizeof(double), h->st));
  const double* src[3] = {fx, fy, fz};
  for (int d = 0; d < 3; ++d) RET(copy_own_to_device(h, h->fcur.field + (long long)d * h->geo.nfa, src[d]));
  h->fcur.mode = FORCE_FIELD;
  h->force_version++;
  return LBG_OK;
}

int lbg_lb_time(lbg_handle h, int64_t* t) {
  if (!h || !t) return LBG_ERR_INVALID_ARG;
  *t = h->t;
  return LBG_OK;
}

// ANY(n<0) (equilibration.f90:248) is a global test: the slabs agree on the first step of the batch that saw a
// negative population anywhere (an unchecked step's flag never stopped a kernel, so every rank ran the same steps)
static int agree_on_negative_step(lbg_handle h, int chunk, int* executed, int* neg) {
  unsigned long long w = *neg ? (unsigned long long)(chunk - *executed + 1) : 0ull;   // larger = earlier step
  CK(cudaMemcpyAsync(h->counts, &w, sizeof(w), cudaMemcpyHostToDevice, h->st));
  RET(allreduce(h, h->counts, 1, AR_U64_MAX));
  RET(wait_halo(h));
  CK(cudaMemcpyAsync(&w, h->counts, sizeof(w), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  RET(check_p2p(h));
  if (w) {
    const int first = chunk - (int)w + 1;
    if (first < *executed || !*neg) *executed = first;
    *neg = 1;
  }
  return LBG_OK;
}

// In-place (AA) stepping, see lb_aa_kernels.cu.  One kernel per step, plus a moments pass on the steps
// whose l2err is asked for (and after the last step of a call, for the negative-population guard).
static int lb_step_in_place(lbg_handle h, double tau, int nsteps, int check_every, double target_error,
                            double* l2err_hist, int* steps_done, int* converged) {
  auto checked = [&](long long s) { return check_every > 0 && (s % check_every) == 0; };
  int total = 0;
  while (total < nsteps) {
    const int chunk = (nsteps - total) < SLOT_CAP ? (nsteps - total) : SLOT_CAP;
    CK(cudaMemsetAsync(h->l2_slots, 0, 2 * (size_t)chunk * sizeof(unsigned long long), h->st));
    CK(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->st));
    double *scr_j = nullptr, *scr_c = nullptr;
    Force& f_last = h->prev_equals_cur ? h->fcur : h->fprev;
    ForceSel fs_first, fs_rest;
    RET(select_force(h, f_last, h->fcur, &fs_first, &scr_j, &scr_c));   // first kernel: j(t) with the force of step t
    RET(select_force(h, h->fcur, h->fcur, &fs_rest, &scr_c, &scr_c));
    auto base_args = [&](const ForceSel& fs) {
      LBArgs a{};
      a.geo = h->geo;
      a.k = h->k;
      a.fin = h->f[0];
      a.fout = h->f[0];
      a.fid_begin = own_begin(h);
      a.fid_end = own_end(h);
      a.w1 = 1.0 - 1.0 / tau;
      a.w2 = 1.0 / tau;
      a.w3 = 1.0 - 1.0 / (2.0 * tau);
      for (int d = 0; d < 3; ++d) {
        a.fj[d] = fs.fj[d];
        a.fc[d] = fs.fc[d];
      }
      a.fj_field = fs.fj_field;
      a.fc_field = fs.fc_field;
      a.l2_slots = h->l2_slots;
      a.target = target_error;
      a.ctrl = h->ctrl;
      return a;
    };
    // j(t) for the first check of this call
    if (checked(h->t + 1) && h->j_valid_step != h->t) {
      if (h->precollision) {
        CK(cudaMemcpyAsync(h->jpp[h->jc], h->mom + h->geo.nfa, 3 * (size_t)h->geo.nfa * sizeof(double),
                           cudaMemcpyDeviceToDevice, h->st));
      } else {
        ForceSel fl;
        RET(select_force(h, f_last, f_last, &fl, &scr_j, &scr_j));
        LBArgs a = base_args(fl);
        a.l2_slots = nullptr;
        a.jnew = h->jpp[h->jc];
        RET(wait_halo(h));
        h->launches += launch_aa_moments(a, fl.mode, h->aa_swapped, false, true, nullptr, nullptr, h->grid_aa, h->st);
      }
      h->j_valid_step = h->t;
    }
    bool swapped = h->aa_swapped;
    int jold = h->jc;
    for (int i = 0; i < chunk; ++i) {
      const long long s = h->t + 1 + i;
      const ForceSel& fs = (i == 0) ? fs_first : fs_rest;
      LBArgs a = base_args(fs);
      a.batch_idx = i;
      a.prev_checked = (i > 0 && checked(s - 1)) ? 1 : 0;
      a.prev_may_stop = (s - 1 > 2) ? 1 : 0;
      const bool first = h->precollision && i == 0;
      a.neg_flag_local = h->nranks > 1 ? (a.prev_checked ? 2 : 1) : 0;
      RET(wait_halo(h));
      if (h->nranks == 1) {
        h->launches += launch_aa_step(a, tau == 1.0, fs.mode, swapped, first, h->mom, h->grid_aa, h->st);
      } else {
        // Across slabs (tests/test_aa_slab_protocol_model.py): boundary planes first, then their exchange runs
        // beside the interior planes' kernel.  After a local step (N -> S) slot inv(l) holds direction l, so the top
        // plane's cz = -1 slots feed the upper neighbour's pull and the bottom plane's cz = +1 slots the lower
        // one's (the two-lattice lists, swapped).  After a pull/push step (S -> N) the pushes that landed in my
        // halo planes travel back into the neighbours' boundary planes, merged there under the owner mask.
        const std::vector<long long>& ps = h->pstart;
        const int nz = h->geo.nzl;
        auto launch = [&](long long b, long long e) {
          a.fid_begin = b;
          a.fid_end = e;
          h->launches += launch_aa_step(a, tau == 1.0, fs.mode, swapped, first, h->mom, h->grid_aa, h->st);
        };
        launch(ps[1], ps[2]);
        if (nz > 1) launch(ps[nz], ps[nz + 1]);
        if (!swapped) RET(halo_exchange(h, h->f[0], DOWN_L, 5, UP_L, 5));
        else RET(halo_exchange(h, h->f[0], UP_L, 5, DOWN_L, 5, true));
        if (nz > 2) launch(ps[2], ps[nz]);
      }
      swapped = !swapped;
      const bool chk = checked(s), wj = chk || checked(s + 1), last = (i == chunk - 1);
      if (chk || wj || last) {
        LBArgs m = base_args(fs_rest);   // the state of step s carries the force of step s in j
        m.batch_idx = i;
        m.jold = h->jpp[jold];
        m.jnew = h->jpp[1 - jold];
        RET(wait_halo(h));   // the S layout is read through the halo planes; the N layout needs the merged return trip
        h->launches += launch_aa_moments(m, fs_rest.mode, swapped, chk, wj, nullptr, nullptr, h->grid_aa, h->st);
        if (chk && h->nranks > 1) RET(allreduce(h, h->l2_slots + 2 * i, 2, AR_U64_MAX));
      }
      jold = 1 - jold;
    }
    RET(wait_halo(h));
    CK(cudaMemcpyAsync(h->h_l2, h->l2_slots, 2 * (size_t)chunk * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    RET(check_p2p(h));
    cudaFree(scr_j);
    if (scr_c != scr_j) cudaFree(scr_c);
    int executed = chunk, conv = 0, neg = 0;
    for (int i = 0; i < chunk; ++i)
      if (h->h_l2[2 * i + 1]) {
        executed = i + 1;
        neg = 1;
        break;
      }
    if (h->nranks > 1 && check_every != 1) RET(agree_on_negative_step(h, chunk, &executed, &neg));
    for (int i = 0; i < executed; ++i) {
      const long long s = h->t + 1 + i;
      double v = std::numeric_limits<double>::quiet_NaN();
      if (checked(s)) std::memcpy(&v, &h->h_l2[2 * i], sizeof(double));
      if (l2err_hist) l2err_hist[total + i] = v;
      if (neg && i == executed - 1) break;
      if (checked(s) && s > 2 && v <= target_error) {
        executed = i + 1;
        conv = 1;
        break;
      }
    }
    h->t += executed;
    if (executed > 0) {
      h->precollision = false;
      if (executed & 1) {
        h->aa_swapped = !h->aa_swapped;
        h->jc = 1 - h->jc;
      }
      const long long last = h->t;
      if (checked(last) || checked(last + 1)) h->j_valid_step = last;
      h->prev_equals_cur = true;
    }
    total += executed;
    if (steps_done) *steps_done = total;
    if (neg) return fail(h, LBG_ERR_NEGATIVE_POPULATION, lbg_status_string(LBG_ERR_NEGATIVE_POPULATION));
    if (conv) {
      if (converged) *converged = 1;
      break;
    }
  }
  return LBG_OK;
}

int lbg_lb_set_in_place(lbg_handle h, int on) {
  if (!h) return LBG_ERR_INVALID_ARG;
  if (on && h->nranks > 1 && !h->p2p)
    return fail(h, LBG_ERR_UNSUPPORTED, "in-place (AA) mode across slabs needs the peer-to-peer halo transport");
  h->in_place = on != 0;
  h->phase = PH_CREATED;  // takes effect with the next lbg_lb_init / lbg_lb_upload
  return LBG_OK;
}

int lbg_lb_step(lbg_handle h, double tau, int nsteps, int check_every, double target_error, double* l2err_hist,
                int* steps_done, int* converged) {
  if (!h || nsteps < 0 || check_every < 0) return LBG_ERR_INVALID_ARG;
  if (steps_done) *steps_done = 0;
  if (converged) *converged = 0;
  if (h->phase != PH_LB) return fail(h, LBG_ERR_STATE, "lbg_lb_step needs lbg_lb_init/lbg_lb_upload first");
  if (!(tau > 0.0) || tau < 0.5) return fail(h, LBG_ERR_RELAXATION_TIME, lbg_status_string(LBG_ERR_RELAXATION_TIME));
  CK(cudaSetDevice(h->device));
  if (h->in_place) return lb_step_in_place(h, tau, nsteps, check_every, target_error, l2err_hist, steps_done, converged);
  auto checked = [&](long long s) { return check_every > 0 && (s % check_every) == 0; };
  int total = 0;
  while (total < nsteps) {
    const int chunk = (nsteps - total) < SLOT_CAP ? (nsteps - total) : SLOT_CAP;
    CK(cudaMemsetAsync(h->l2_slots, 0, 2 * (size_t)chunk * sizeof(unsigned long long), h->st));
    CK(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->st));
    RET(ensure_collided(h, tau, checked(h->t + 1)));
    ForceSel fs;
    double* scr = nullptr;
    RET(select_force(h, h->fcur, h->fcur, &fs, &scr, &scr));
    int fin = 1 - h->src, jold = h->jc;
    for (int i = 0; i < chunk; ++i) {
      const long long s = h->t + 1 + i;
      StepFlags fl{};
      fl.check = checked(s);
      fl.writej = fl.check || checked(s + 1);
      fl.batch_idx = i;
      fl.prev_checked = (i > 0 && checked(s - 1)) ? 1 : 0;
      fl.prev_may_stop = (s - 1 > 2) ? 1 : 0;
      fl.target = target_error;
      RET(enqueue_lb_kernel(h, fin, tau, fs, jold, fl));
      fin = 1 - fin;
      jold = 1 - jold;
    }
    RET(wait_halo(h));
    CK(cudaMemcpyAsync(h->h_l2, h->l2_slots, 2 * (size_t)chunk * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(h->h_ctrl, h->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    RET(check_p2p(h));
    if (scr) cudaFree(scr);
    int executed = chunk, conv = 0, neg = 0;
    for (int i = 0; i < chunk; ++i)
      if (h->h_l2[2 * i + 1]) {  // on checked steps the flag is global (it travels with l2err); otherwise this slab's
        executed = i + 1;
        neg = 1;
        break;
      }
    if (h->nranks > 1 && check_every != 1) RET(agree_on_negative_step(h, chunk, &executed, &neg));
    for (int i = 0; i < executed; ++i) {
      const long long s = h->t + 1 + i;
      double v = std::numeric_limits<double>::quiet_NaN();
      if (checked(s)) std::memcpy(&v, &h->h_l2[2 * i], sizeof(double));
      if (l2err_hist) l2err_hist[total + i] = v;
      if (neg && i == executed - 1) break;  // the reference stops before l2err on that step
      if (checked(s) && s > 2 && v <= target_error) {  // equilibration.f90:346
        executed = i + 1;
        conv = 1;
        break;
      }
    }
    // bookkeeping: `executed` kernels ran to completion, later ones returned at once
    h->t += executed;
    if (executed > 0) {
      h->precollision = false;
      if (executed & 1) {
        h->src = 1 - h->src;
        h->jc = 1 - h->jc;
      }
      const long long last = h->t;
      if (checked(last) || checked(last + 1)) h->j_valid_step = last;
      h->prev_equals_cur = true;
      h->collided_ok = true;
      h->collided_tau = tau;
      h->collided_force_version = h->force_version;
    }
    total += executed;
    if (steps_done) *steps_done = total;
    if (neg) return fail(h, LBG_ERR_NEGATIVE_POPULATION, lbg_status_string(LBG_ERR_NEGATIVE_POPULATION));
    if (conv) {
      if (converged) *converged = 1;
      break;
    }
  }
  return LBG_OK;
}

static int download_moments(lbg_handle h, double* rho, double* jx, double* jy, double* jz, bool async) {
  if (!h) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(refresh_moments(h, nullptr));
  double* dst[4] = {rho, jx, jy, jz};
  const double* src[4];
  for (int c = 0; c < 4; ++c) src[c] = h->mom + (long long)c * h->geo.nfa;
  RET(wait_halo(h));
  return copy_own_to_host_many(h, dst, src, 4, nullptr, async);
}

int lbg_lb_download_moments(lbg_handle h, double* rho, double* jx, double* jy, double* jz) {
  return download_moments(h, rho, jx, jy, jz, false);
}

int lbg_lb_download_moments_async(lbg_handle h, double* rho, double* jx, double* jy, double* jz) {
  return download_moments(h, rho, jx, jy, jz, true);
}

int lbg_wait_transfers(lbg_handle h) {
  if (!h) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  return wait_transfers(h);
}

int lbg_lb_download_populations(lbg_handle h, double* n) {
  if (!h || !n) return LBG_ERR_INVALID_ARG;
  if (h->phase != PH_LB) return fail(h, LBG_ERR_STATE, "no Lattice-Boltzmann state");
  CK(cudaSetDevice(h->device));
  const double* from = nullptr;
  double* tmp = nullptr;
  bool pull = false;
  if (h->in_place) {
    if (!h->aa_swapped) {
      from = h->f[0];  // layout N(t) is the reference's own state
    } else {
      CK(cudaMalloc(&tmp, 19 * (size_t)h->geo.nfa * sizeof(double)));
      RET(refresh_moments(h, tmp));
      from = tmp;
    }
  } else if (h->precollision) {
    from = h->f[h->src];
  } else {
    // n(t) is rebuilt from n*(t) by the pull rule, one direction at a time, straight into the staging buffer:
    // the stepping state (collided lattice, halo sequence) is not touched, so the call is local to this rank
    from = h->f[h->src];
    pull = true;
  }
  double* dst[19];
  const double* src[19];
  for (int l = 0; l < 19; ++l) {
    dst[l] = n + (size_t)l * h->nown;
    src[l] = from + (long long)l * h->geo.nfa;
  }
  RET(wait_halo(h));
  const int rc = copy_own_to_host_many(h, dst, src, 19, pull ? from : nullptr);
  cudaFree(tmp);
  return rc;
}

int lbg_lb_profiles(lbg_handle h, int axis, int raw, double* out) {
  if (!h || !out || axis < 0 || axis > 2) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(refresh_moments(h, nullptr));
  const int rows = axis == 0 ? h->geo.lx : (axis == 1 ? h->geo.ly : h->geo.nzl);
  RET(wait_transfers(h));
  RET(ensure_stage(h));
  // scratch in the staging buffer when it is large enough (5 values per row), else a one-off allocation
  const bool own_alloc = (size_t)rows * 5 > (size_t)STAGE_SLOTS * (size_t)h->nown;
  double* d = h->stage;
  if (own_alloc) CK(cudaMalloc(&d, (size_t)rows * 5 * sizeof(double)));
  ProfileArgs a{};
  a.geo = h->geo;
  a.mom = h->mom;
  a.axis = axis;
  a.eps = std::numeric_limits<double>::epsilon();
  a.out = d;
  h->launches += launch_profile(a, rows, h->st);
  std::vector<double> tmp((size_t)rows * 5);
  CK(cudaMemcpyAsync(tmp.data(), d, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  if (own_alloc) cudaFree(d);
  for (int p = 0; p < rows; ++p) {
    if (raw) {
      for (int c = 0; c < 5; ++c) out[(size_t)p * 5 + c] = tmp[(size_t)p * 5 + c];
    } else {
      out[(size_t)p * 4 + 0] = tmp[(size_t)p * 5 + 0];
      out[(size_t)p * 4 + 1] = tmp[(size_t)p * 5 + 1];
      out[(size_t)p * 4 + 2] = tmp[(size_t)p * 5 + 2];
      const double cnt = tmp[(size_t)p * 5 + 4];
      out[(size_t)p * 4 + 3] = tmp[(size_t)p * 5 + 3] / (cnt > 1.0 ? cnt : 1.0);
    }
  }
  return LBG_OK;
}

int lbg_lb_total_flux(lbg_handle h, double out[3]) {
  if (!h || !out) return LBG_ERR_INVALID_ARG;
  std::vector<double> prof((size_t)h->geo.nzl * 5);
  RET(lbg_lb_profiles(h, 2, 1, prof.data()));
  out[0] = out[1] = out[2] = 0.0;
  for (int p = 0; p < h->geo.nzl; ++p)
    for (int c = 0; c < 3; ++c) out[c] += prof[(size_t)p * 5 + c];
  return LBG_OK;
}

int lbg_lb_slice(lbg_handle h, int axis, int index, double* rho, double* jx, double* jy, double* jz) {
  if (!h || axis < 0 || axis > 2 || index < 0) return LBG_ERR_INVALID_ARG;
  const int extent = axis == 0 ? h->geo.lx : (axis == 1 ? h->geo.ly : h->geo.nzl);
  if (index >= extent) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(refresh_moments(h, nullptr));
  RET(wait_transfers(h));
  RET(ensure_stage(h));   // a plane is never larger than the staging buffer (lx, ly, nzl >= 1)
  const size_t n = (size_t)(axis == 0 ? h->geo.ly : h->geo.lx) * (size_t)(axis == 2 ? h->geo.ly : h->geo.nzl);
  if (4 * n > (size_t)STAGE_SLOTS * (size_t)h->nown) return fail(h, LBG_ERR_INVALID_ARG, "lbg_lb_slice: plane larger than the lattice");
  h->launches += launch_slice(h->geo, h->mom, axis, index, h->stage, h->st);
  double* dst[4] = {rho, jx, jy, jz};
  for (int c = 0; c < 4; ++c)
    if (dst[c]) CK(cudaMemcpyAsync(dst[c], h->stage + (size_t)c * n, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return LBG_OK;
}

int lbg_lb_probe(lbg_handle h, int i, int j, int k, double out[4]) {
  if (!h || !out || i < 0 || j < 0 || k < 0 || i >= h->geo.lx || j >= h->geo.ly || k >= h->geo.nzl) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(refresh_moments(h, nullptr));
  const long long g = (long long)i + (long long)h->geo.lx * j + (long long)h->geo.plane * (k + 1);
  uint2 w;
  CK(cudaMemcpyAsync(&w, h->words + (g >> 5), sizeof(w), cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  const unsigned bit = (unsigned)(g & 31);
  out[0] = out[1] = out[2] = out[3] = 0.0;  // solid node: everything is 0
  if (!((w.x >> bit) & 1u)) return LBG_OK;
  const long long fid = (long long)w.y + __builtin_popcount(w.x & ((1u << bit) - 1u));
  // out = jx, jy, jz, density
  for (int c = 0; c < 4; ++c)
    CK(cudaMemcpyAsync(&out[c], h->mom + (long long)((c + 1) % 4) * h->geo.nfa + fid, sizeof(double),
                       cudaMemcpyDeviceToHost, h->st));
  CK(cudaStreamSynchronize(h->st));
  return LBG_OK;
}

// --------------------------------------------------------------------------- Phase B
// rho_host != NULL: density and momentum density come from the driver's arrays (lbg_mp_init_from_moments)
// instead of the resident Lattice-Boltzmann state.
static int mp_init_impl(lbg_handle h, double Db, double ka, double kd, const double f_ext[3], double vacf0[3],
                        const double* const host_mom[4]) {
  if (!h || !f_ext) return LBG_ERR_INVALID_ARG;
  if (!host_mom && h->phase != PH_LB)
    return fail(h, LBG_ERR_STATE, "lbg_mp_init needs the Lattice-Boltzmann state (density, momentum)");
  const double eps = std::numeric_limits<double>::epsilon();
  if (Db <= eps) return fail(h, LBG_ERR_TRACER_DB, lbg_status_string(LBG_ERR_TRACER_DB));       // drop_tracers.f90:89
  if (ka < -eps || kd < -eps) return fail(h, LBG_ERR_TRACER_KA_KD, lbg_status_string(LBG_ERR_TRACER_KA_KD));
  CK(cudaSetDevice(h->device));
  PhaseTimer tm("lbg_mp_init");
  if (host_mom) {
    RET(wait_halo(h));
    CK(cudaMemsetAsync(h->mom, 0, 4 * (size_t)h->geo.nfa * sizeof(double), h->st));
    for (int c = 0; c < 4; ++c) RET(copy_own_to_device(h, h->mom + (long long)c * h->geo.nfa, host_mom[c]));
    RET(ring_barrier(h));   // a neighbour's halo push of its moments must not be wiped by my memset
  } else {
    RET(refresh_moments(h, nullptr));
  }
  const Geo& g = h->geo;
  if (h->nranks > 1) {
    const int all4[4] = {0, 1, 2, 3};
    RET(halo_exchange(h, h->mom, all4, 4, all4, 4));
    RET(wait_halo(h));
  }
  // module_moment_propagation.f90:46-56,68,97
  const double K = (std::fabs(kd) <= eps) ? 0.0 : ka / kd;
  h->ads = std::fabs(K) > eps ? 1 : 0;
  h->Db = Db;
  h->ka = ka;
  h->kd = kd;
  long long nf = h->n_fluid, nif = h->n_if_fluid;
  if (h->nranks > 1) {
    unsigned long long c2[2] = {(unsigned long long)nf, (unsigned long long)nif};
    CK(cudaMemcpyAsync(h->counts, c2, sizeof(c2), cudaMemcpyHostToDevice, h->st));
    RET(allreduce(h, h->counts, 2, AR_U64_SUM));
    RET(wait_halo(h));
    RET(read_small_async(h, 0, h->counts, sizeof(c2)));
    CK(cudaStreamSynchronize(h->st));
    std::memcpy(c2, h->h_small, sizeof(c2));
    nf = (long long)c2[0];
    nif = (long long)c2[1];
  }
  const double Pstat = (double)nf + K * (double)nif;
  volatile double one = 1.0, three = 3.0;
  const double kBT = one / three;
  const double lambda = 4.0 * Db / kBT;
  // Phase B aliases the population buffers (and needs the second one even after an in-place Phase A)
  RET(ensure_second_lattice(h));
  const size_t nb = (size_t)g.nfa * sizeof(double);
  h->q = h->f[0];
  h->s = h->f[1];
  h->P[0] = h->f[1] + 4 * g.nfa;
  h->P[1] = h->f[1] + 7 * g.nfa;
  h->A[0] = h->f[1] + 10 * g.nfa;
  h->A[1] = h->f[1] + 13 * g.nfa;
  h->nbt01 = reinterpret_cast<uint32_t*>(h->f[0] + 18 * g.nfa);
  h->nbt27 = reinterpret_cast<uint32_t*>(h->f[1] + 16 * g.nfa);
  // Measured (profiles/variants_r3.txt): the table pays off on porous lattices (-13 % on cfg5w; -7 % on the
  // 256^3 BCC lattice, where 3 of 4 warps hold a seam node); on an all-fluid lattice the rank lookups
  // are perfectly coalesced and cost no more than the table's 28 extra bytes per node.
  const double phi = (double)h->n_fluid / (double)(h->nown > 0 ? h->nown : 1);
  h->mp_use_nbt = (g.lx >= 64 && phi < 0.95) ? 1 : 0;
  if (const char* e = std::getenv("LBG_MP_NBT")) h->mp_use_nbt = std::atoi(e) ? 1 : 0;
  CK(cudaMemsetAsync(h->f[1], 0, 19 * nb, h->st));
  CK(cudaMemsetAsync(h->mp_err, 0, sizeof(int), h->st));
  CK(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->st));
  // compact adsorbed storage over the interfacial fluid nodes of the own planes
  h->a_stride = 0;
  if (h->ads) {
    const long long ngroups = g.nfa >> 5;
    if (!h->awords) CK(cudaMalloc(&h->awords, (size_t)ngroups * sizeof(uint2)));   // normally done in lbg_create*
    h->launches += launch_build_awords(g, own_begin(h), own_end(h), h->awords, h->st);
    CK(cudaMemsetAsync(h->counts, 0, sizeof(unsigned long long), h->st));
    h->launches += launch_scan_ranks(h->awords, ngroups, h->counts, h->st);
    unsigned long long slots = 0;
    RET(read_small_async(h, 0, h->counts, sizeof(slots)));
    CK(cudaStreamSynchronize(h->st));
    std::memcpy(&slots, h->h_small, sizeof(slots));
    tm.lap("moments + awords (first sync)");
    h->a_stride = ((long long)slots + 31) / 32 * 32;
    if (h->a_stride > g.nfa) return fail(h, LBG_ERR_STATE, "adsorbed storage does not fit");  // cannot happen: slots <= nf + 3 nf / 32... guard anyway
  }
  if (!h->mp_use_nbt) {   // rank-lookup path: nodes whose neighbour ids follow by arithmetic skip the lookups
    const size_t ngroups = (size_t)(g.nfa >> 5);
    if (!h->rwords) CK(cudaMalloc(&h->rwords, ngroups * sizeof(uint32_t)));
    CK(cudaMemsetAsync(h->rwords, 0, ngroups * sizeof(uint32_t), h->st));
  }
  MPInitArgs a{};
  a.geo = g;
  a.k = h->k;
  a.mom = h->mom;
  a.rwords = h->mp_use_nbt ? nullptr : h->rwords;
  a.q = h->q;
  a.nbt01 = h->nbt01;
  a.nbt27 = h->nbt27;
  a.s = h->s;
  a.P0 = h->P[0];
  a.fid_begin = own_begin(h);
  a.fid_end = own_end(h);
  for (int d = 0; d < 3; ++d) a.f[d] = f_ext[d];
  for (int i = 0; i < 3; ++i) a.lambda_w[i] = lambda * h->k.a0[i];
  a.bw = 1.0 / Pstat;
  a.ka = ka;
  a.ads = h->ads;
  a.partial = h->partial;
  a.err = h->mp_err;
  const long long nblk = (h->n_fluid + BLOCK - 1) / BLOCK;
  const int grid = (int)(nblk < 1 ? 1 : (nblk < h->grid_mp ? nblk : h->grid_mp));
  h->launches += launch_mp_init(a, grid, h->st);
  std::vector<double> part((size_t)grid * 3);
  int bad = 0;
  RET(read_small_async(h, 0, h->partial, part.size() * sizeof(double)));
  RET(read_small_async(h, SMALL_CAP - 64, h->mp_err, sizeof(int)));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaGetLastError());
  std::memcpy(part.data(), h->h_small, part.size() * sizeof(double));
  std::memcpy(&bad, h->h_small + SMALL_CAP - 64, sizeof(int));
  tm.lap("init kernel + partials (second sync)");
  double v0[3] = {0, 0, 0};
  for (int b = 0; b < grid; ++b)
    for (int d = 0; d < 3; ++d) v0[d] += part[(size_t)b * 3 + d];
  if (h->nranks > 1) {
    CK(cudaMemcpyAsync(h->vacf_slots, v0, sizeof(v0), cudaMemcpyHostToDevice, h->st));
    RET(allreduce(h, h->vacf_slots, 3, AR_F64_SUM));
    RET(wait_halo(h));
    RET(read_small_async(h, 0, h->vacf_slots, sizeof(v0)));
    unsigned long long badsum = bad ? 1ull : 0ull;
    CK(cudaMemcpyAsync(h->counts, &badsum, sizeof(badsum), cudaMemcpyHostToDevice, h->st));
    RET(allreduce(h, h->counts, 1, AR_U64_MAX));
    RET(wait_halo(h));
    RET(read_small_async(h, 64, h->counts, sizeof(badsum)));
    CK(cudaStreamSynchronize(h->st));
    std::memcpy(v0, h->h_small, sizeof(v0));
    std::memcpy(&badsum, h->h_small + 64, sizeof(badsum));
    bad = badsum ? 1 : 0;
    // P(now) halo planes for the first step
    const int all3[3] = {0, 1, 2};
    RET(halo_exchange(h, h->P[0], all3, 3, all3, 3));
  }
  if (vacf0)
    for (int d = 0; d < 3; ++d) vacf0[d] = v0[d];
  // strip order over the planes the big propagate launch covers
  if (h->nranks == 1) RET(build_strips(h, 1, g.nzl, &h->strips));
  else RET(build_strips(h, 2, g.nzl - 1, &h->strips));
  h->mp_bad = bad;
  h->phase = PH_MP;
  h->it = 0;
  h->pc = 0;
  return LBG_OK;
}

int lbg_mp_init(lbg_handle h, double Db, double ka, double kd, const double f_ext[3], double vacf0[3]) {
  return mp_init_impl(h, Db, ka, kd, f_ext, vacf0, nullptr);
}

int lbg_mp_init_from_moments(lbg_handle h, const double* rho, const double* jx, const double* jy, const double* jz,
                             double Db, double ka, double kd, const double f_ext[3], double vacf0[3]) {
  if (!h || !rho || !jx || !jy || !jz) return LBG_ERR_INVALID_ARG;
  const double* const m[4] = {rho, jx, jy, jz};
  return mp_init_impl(h, Db, ka, kd, f_ext, vacf0, m);
}

int lbg_mp_step(lbg_handle h, int nsteps, double* vacf, int* steps_done, int* converged) {
  if (!h || nsteps < 0) return LBG_ERR_INVALID_ARG;
  if (steps_done) *steps_done = 0;
  if (converged) *converged = 0;
  if (h->phase != PH_MP) return fail(h, LBG_ERR_STATE, "lbg_mp_step needs lbg_mp_init first");
  if (nsteps == 0) return LBG_OK;
  // the remaining fraction is static: the reference would stop in its first propagate call (:249,257)
  if (h->mp_bad) return fail(h, LBG_ERR_RESTPART_NEGATIVE, lbg_status_string(LBG_ERR_RESTPART_NEGATIVE));
  CK(cudaSetDevice(h->device));
  const Geo& g = h->geo;
  const std::vector<long long>& ps = h->pstart;
  const int nz = g.nzl;
  const double lim = 1.0 / (2.0 * g.lx * g.ly * h->lz_global / h->Db);
  auto is_conv = [&](long long it, const double* v) {
    return it > 2 && std::fabs(v[0]) < lim && std::fabs(v[1]) < lim && std::fabs(v[2]) < lim &&
           std::fabs(v[0]) < 1.e-12 && std::fabs(v[1]) < 1.e-12 && std::fabs(v[2]) < 1.e-12;
  };
  int total = 0;
  while (total < nsteps) {
    const int chunk = (nsteps - total) < SLOT_CAP ? (nsteps - total) : SLOT_CAP;
    CK(cudaMemsetAsync(h->ctrl, 0, sizeof(Ctrl), h->st));
    CK(cudaMemsetAsync(h->vacf_slots, 0, (size_t)chunk * 3 * sizeof(double), h->st));
    int pc = h->pc;
    // The stop test needs the global vacf of a step.  On one GPU step i looks at step i-1.  Across GPUs
    // the all-reduce of step i is left a whole step to complete: step i looks at step i-2, so at most
    // one step runs past the converged one -- and its input buffers are exactly the wanted state.
    // Measured at N=2 (profiles/multigpu_r1.txt): with peer-to-peer halos the plain blocking all-reduce is
    // fastest (the GPU is otherwise idle at that point); the lagged one only pays off on the NCCL path.
    int lag = (h->nranks > 1 && !h->p2p) ? 2 : 1;
    if (const char* e = std::getenv("LBG_MP_LAG")) lag = (h->nranks > 1 && std::atoi(e) == 2) ? 2 : 1;
    for (int i = 0; i < chunk; ++i) {
      const long long it = h->it + 1 + i;
      MPArgs a{};
      a.geo = g;
      a.geo.tpc = 0;
      a.tpc = h->mp_tpc;
      a.q = h->q;
      a.nbt01 = h->nbt01;
      a.nbt27 = h->nbt27;
      a.use_nbt = h->mp_use_nbt;
      a.s = h->s;
      a.Pnow = h->P[pc];
      a.Pnext = h->P[1 - pc];
      a.Anow = h->A[pc];
      a.Anext = h->A[1 - pc];
      a.awords = h->awords;
      a.a_stride = h->a_stride;
      a.rwords = (h->mp_use_nbt || h->mp_arith == 0) ? nullptr : h->rwords;
      a.ka = h->ka;
      a.kd = h->kd;
      a.one_minus_kd = 1.0 - h->kd;
      a.ads = h->ads;
      a.partial = h->partial;
      a.vacf_slots = h->vacf_slots;
      a.batch_idx = i;
      a.accumulate = 1;  // slots are zeroed per batch; every launch of a step adds its share
      a.check_slot = (i - lag >= 0 && (it - lag) > 2) ? i - lag : -1;
      a.lim = lim;
      a.ctrl = h->ctrl;
      RET(wait_halo(h));
      auto launch = [&](long long b, long long e, bool use_strips = false) {
        a.fid_begin = b;
        a.fid_end = e;
        a.nseg = 0;
        if (use_strips && h->strips.nseg > 0) {
          a.nseg = h->strips.nseg;
          a.ntiles = h->strips.ntiles;
          a.tile_cum = h->strips.tile_cum;
          a.seg_begin = h->strips.seg_begin;
          a.seg_end = h->strips.seg_end;
        }
        h->launches += launch_mp_step(a, h->grid_mp, h->st);
      };
      if (h->nranks == 1) {
        launch(own_begin(h), own_end(h), true);
      } else {
        const int all3[3] = {0, 1, 2};
        if (lag == 2 && i >= 2) CK(cudaStreamWaitEvent(h->st, h->ev_ar[i & 1], 0));  // all-reduce of step i-2
        launch(ps[1], ps[2]);
        if (nz > 1) launch(ps[nz], ps[nz + 1]);
        RET(halo_exchange(h, h->P[1 - pc], all3, 3, all3, 3));
        if (nz > 2) launch(ps[2], ps[nz], true);
        if (lag == 1) {  // blocking variant: the next step waits for this step's global vacf
          RET(wait_halo(h));
          RET(allreduce(h, h->vacf_slots + 3 * i, 3, AR_F64_SUM));
          pc = 1 - pc;
          continue;
        }
        // all-reduce of this step's vacf on the communication stream, not waited for here
        CK(cudaEventRecord(h->ev_ready, h->st));
        CK(cudaStreamWaitEvent(h->st_comm, h->ev_ready, 0));
        NK(g_nccl.AllReduce(h->vacf_slots + 3 * i, h->vacf_slots + 3 * i, 3, ncclDouble, ncclSum, h->comm, h->st_comm));
        CK(cudaEventRecord(h->ev_ar[i & 1], h->st_comm));
      }
      pc = 1 - pc;
    }
    RET(wait_halo(h));
    if (h->nranks > 1 && lag == 2) {
      CK(cudaStreamWaitEvent(h->st, h->ev_ar[0], 0));
      CK(cudaStreamWaitEvent(h->st, h->ev_ar[1], 0));
    }
    CK(cudaMemcpyAsync(h->h_vacf, h->vacf_slots, (size_t)chunk * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    RET(check_p2p(h));
    int executed = chunk, conv = 0, ran = chunk;
    for (int i = 0; i < chunk; ++i) {
      const long long it = h->it + 1 + i;
      if (vacf)
        for (int d = 0; d < 3; ++d) vacf[(size_t)(total + i) * 3 + d] = h->h_vacf[(size_t)i * 3 + d];
      if (is_conv(it, h->h_vacf + (size_t)i * 3)) {
        executed = i + 1;                                  // steps the reference would have made
        ran = (i + lag < chunk) ? i + lag : chunk;         // step kernels that actually ran
        conv = 1;
        break;
      }
    }
    // `ran` kernels flipped the buffers; a kernel that ran past the converged step left its input intact
    if (ran & 1) h->pc = 1 - h->pc;
    if ((ran - executed) & 1) h->pc = 1 - h->pc;
    h->it += executed;
    total += executed;
    if (steps_done) *steps_done = total;
    if (conv) {
      if (converged) *converged = 1;
      break;
    }
  }
  return LBG_OK;
}

int lbg_mp_download(lbg_handle h, double* P, double* Pads) {
  if (!h) return LBG_ERR_INVALID_ARG;
  if (h->phase != PH_MP) return fail(h, LBG_ERR_STATE, "lbg_mp_download needs lbg_mp_init first");
  CK(cudaSetDevice(h->device));
  RET(wait_halo(h));
  RET(wait_transfers(h));
  RET(ensure_stage(h));
  if (P) {
    h->launches += launch_scatter3_to_dense_aos(h->geo, h->P[h->pc], h->stage, h->st);
    CK(cudaMemcpyAsync(P, h->stage, (size_t)h->nown * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
  }
  if (Pads) {
    h->launches += launch_scatter3_compact_to_dense_aos(h->geo, h->ads ? h->awords : nullptr, h->A[h->pc], h->a_stride,
                                                        h->stage, h->st);
    CK(cudaMemcpyAsync(Pads, h->stage, (size_t)h->nown * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
  }
  return LBG_OK;
}

// --------------------------------------------------------------------------- measurement
int lbg_timer_start(lbg_handle h) {
  if (!h) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(wait_halo(h));
  CK(cudaEventRecord(h->ev_t0, h->st));
  return LBG_OK;
}

int lbg_timer_stop(lbg_handle h, float* ms) {
  if (!h || !ms) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  RET(wait_halo(h));
  CK(cudaEventRecord(h->ev_t1, h->st));
  CK(cudaEventSynchronize(h->ev_t1));
  CK(cudaEventElapsedTime(ms, h->ev_t0, h->ev_t1));
  return LBG_OK;
}

int lbg_launch_count(lbg_handle h, int64_t* n) {
  if (!h || !n) return LBG_ERR_INVALID_ARG;
  *n = h->launches;
  return LBG_OK;
}

int lbg_get_info(lbg_handle h, const char* key, int64_t* value) {
  if (!h || !key || !value) return LBG_ERR_INVALID_ARG;
  const std::string k(key);
  if (k == "mp_neighbour_table") *value = h->mp_use_nbt;
  else if (k == "lb_variant") *value = 100 * h->lb_pipe + 10 * (h->lb_tpc > 0 ? 1 : 0) + h->lb_minb;
  else if (k == "in_place") *value = h->in_place ? 1 : 0;
  else if (k == "p2p") *value = h->p2p ? 1 : 0;
  else if (k == "ipc") {
    bool any = h->peer_ipc[0] || h->peer_ipc[1];
    for (size_t r = 0; r < h->peer_mail_ipc.size(); ++r) any = any || h->peer_mail_ipc[r];
    *value = any ? 1 : 0;
  }
  else if (k == "nranks") *value = h->nranks;
  else if (k == "rank") *value = h->rank;
  else if (k == "fluid_nodes_with_halo") *value = h->nf;
  else return fail(h, LBG_ERR_INVALID_ARG, "lbg_get_info: unknown key " + k);
  return LBG_OK;
}

int lbg_sync(lbg_handle h) {
  if (!h) return LBG_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->st));
  CK(cudaStreamSynchronize(h->st_comm));
  CK(cudaStreamSynchronize(h->st_ar));
  return LBG_OK;
}

}  // extern "C"
