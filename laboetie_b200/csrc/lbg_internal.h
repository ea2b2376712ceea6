// lbg_internal.h -- shared between the kernel translation units and api.cu.
//
// Storage: fluid-compacted structure of arrays.  The lattice nodes of a slab (own planes plus one
// halo plane on each z side) are numbered in the reference's memory order, g = x + lx*(y + ly*p);
// the fluid nodes among them are numbered consecutively in that same order (x fastest), fid = 0..NF-1.
// Every field is a set of arrays indexed by fid with a common stride nfa.  Solid nodes own no storage:
// their populations are 0 for ever (init_simu.f90:32-39, the swap rule of equilibration.f90:204-222)
// and nothing ever reads them.  Consequences for HBM traffic: every 32-byte sector a warp touches
// is full of fluid data, stores are whole sectors with all 32 lanes active, and a porous lattice moves
// only its fluid bytes.  A plane's fluid nodes are one contiguous fid range, so the z-halo planes stay
// contiguous runs that NCCL can send without packing.
//
// The map between the two numberings is a rank structure over the dense order, 8 bytes per 32 nodes:
//   words[w] = { bits: fluid bit of nodes 32w..32w+31,  rank: number of fluid nodes before node 32w }
//   fid(g)   = rank + popc(bits & lower_mask(g & 31))            (one 8-byte load, L1/L2 resident)
//   gidx[fid] = g | interfacial << 31                            (4 bytes per fluid node, streamed)
// It replaces the il/jl/kl neighbour tables of equilibration.f90:109-119 and the per-node flags of
// supercell_definition.f90:115-147.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "d3q19.cuh"

namespace lbg {

constexpr int BLOCK = 256;
constexpr uint32_t GIDX_MASK = 0x7fffffffu;
constexpr uint32_t GIDX_INTERFACIAL = 0x80000000u;

struct Geo {
  int lx, ly, plane, nzl, zwrap;
  long long nfa;          // stride of every per-fluid-node array (multiple of 32 elements)
  const uint2* words;     // {bits, rank} per 32 dense nodes over planes 0..nzl+1
  const uint32_t* gidx;   // per fid: dense index, bit 31 = interfacial
  // g / plane and g / lx for 0 <= g < 2^31 as (g * mul) >> sh  (set_div_magic; exact, see there)
  uint32_t mul_plane, mul_lx;
  int sh_plane, sh_lx;
  int zero;               // always 0, unknown to the compiler: lets a kernel make one value depend on others (issue order)
  // Tile schedule of the step kernels (lattice.cuh Tiles): a persistent grid either way; tpc > 0: an atomic
  // counter hands out chunks of tpc consecutive tiles of BLOCK fids (SMs that get more HBM bandwidth take more
  // chunks); tpc == 0: static tile-stride loop.
  int tpc;
};

// Round-up multiplier for an exact unsigned division of 31-bit dividends (Granlund & Montgomery):
// with l = ceil(log2 d) and m = ceil(2^(31+l) / d), floor(g / d) == (g * m) >> (31 + l) for all
// 0 <= g < 2^31, and m fits in 32 bits.
inline void div_magic(int d, uint32_t* mul, int* sh) {
  int l = 0;
  while ((1LL << l) < (long long)d) ++l;
  const unsigned __int128 one = 1;
  const unsigned __int128 m = ((one << (31 + l)) + (unsigned)d - 1) / (unsigned)d;
  *mul = (uint32_t)m;
  *sh = 31 + l;
}

inline void set_div_magic(Geo& g) {
  div_magic(g.plane, &g.mul_plane, &g.sh_plane);
  div_magic(g.lx, &g.mul_lx, &g.sh_lx);
}

// Device-side control block: lets a batch of step kernels stop itself at the
// reference's exit step without a host round trip per step.
struct Ctrl {
  int stop;                 // set by the first kernel that sees the criterion met
  int neg_step_idx;         // 1 + batch index of the first step with a negative population (0 = none)
  int stop_idx;             // 1 + batch index of the converged step
  unsigned int ticket;      // CTAs that have finished: elects the last one (vacf reduction, counter reset)
  unsigned int tile_next;   // dynamic tile schedule: next chunk to hand out (reset by the last CTA of a launch)
};

enum ForceMode { FORCE_NONE = 0, FORCE_UNIFORM = 1, FORCE_FIELD = 2 };

struct LBArgs {
  Geo geo;
  d3q19::Consts k;
  const double* fin;   // 19 arrays, stride geo.nfa: post-collision populations n*(t)
  double* fout;        // n*(t+1)
  long long fid_begin, fid_end;  // fluid nodes to process
  double w1, w2, w3;         // 1-1/tau, 1/tau, 1-1/(2 tau)
  double fj[3];              // uniform force in effect for step t (enters j as f/2)
  double fc[3];              // uniform force of the collision that follows
  const double* fj_field;    // 3 arrays, stride nfa (FORCE_FIELD)
  const double* fc_field;
  const double* jold;        // 3 arrays: j(t-1)
  double* jnew;              // 3 arrays: j(t)
  unsigned long long* l2_slots;  // two per batch step: l2err (bit pattern of a non-negative double), negative flag
  int batch_idx;                 // index of this step in the batch
  int prev_checked;              // step batch_idx-1 was a checked step of this batch
  int prev_may_stop;             // ... and its global t-1 > 2
  double target;
  Ctrl* ctrl;
  int pipe;                      // two-stage software pipeline (lb_step_pipe_kernel) instead of the plain kernel
  // optional strip order of the plain step kernel (see SegTable in api.cu): nseg == 0 means plain fid order
  int nseg, ntiles;
  const int* tile_cum;          // nseg + 1: tiles before segment k (a segment's tiles start on its 32-fid boundary)
  const long long* seg_begin;   // nseg: first fid of segment k
  const long long* seg_end;     // nseg: one past the last fid of segment k
  int neg_flag_local;            // the previous step's negative-population flag is this slab's only (several slabs, unchecked
                                 // step): do not stop on it -- the ranks agree on the first such step at the end of the batch
};

struct CollideArgs {
  Geo geo;
  d3q19::Consts k;
  const double* fin;  // pre-collision n(t)
  double* fout;
  const double* mom;  // rho, jx, jy, jz: 4 arrays, stride nfa
  long long fid_begin, fid_end;
  double w1, w2, w3;
  double fc[3];
  const double* fc_field;
};

struct MomArgs {
  Geo geo;
  const double* fin;  // n*(t)
  double* mom;        // out: rho, jx, jy, jz
  double* pops;       // out (optional): n(t), 19 arrays stride nfa
  long long fid_begin, fid_end;
  double fj[3];
  const double* fj_field;
};

struct MPInitArgs {
  Geo geo;
  d3q19::Consts k;
  const double* mom;   // rho, jx, jy, jz with valid halo ranges (or zwrap)
  double* q;           // 18 arrays: incoming link probabilities q_l(r) = p_{inv l}(r + c_l)
  double* s;           // 4 arrays: remaining fraction (after -ka where adsorbing), u*_x, u*_y, u*_z
  double* P0;          // 3 arrays: Propagated_Quantity(:, now) at t=0
  long long fid_begin, fid_end;
  double f[3];
  double lambda_w[3];  // lambda * w per kind
  double bw;           // 1 / Pstat
  double ka;
  int ads;
  double* partial;     // per block: vacf0 x,y,z
  int* err;            // set if the remaining fraction < eps somewhere
  uint32_t* nbt01;     // neighbour table words 0,1 (stride nfa), see NBT_* below
  uint32_t* nbt27;     // words 2..4 (stride nfa)
  uint32_t* rwords;    // per group of 32 fids: bit i = node 32 k + i is "regular" (see mp_init_kernel); may be NULL
};

// Phase-B neighbour table: the flow AND the geometry are frozen, so the fluid ids of a node's
// neighbours are static too.  Per fluid node, the rank position c of the centre node (x, y+dy, z+dz) of each of
// the 8 neighbouring rows (dy,dz) != (0,0) (its fid if it is fluid) and a "centre is fluid" bit.  The
// x-neighbours of a row follow without a lookup: fid(x+1) = c + centre_fluid, fid(x-1) = c - 1 (if that node is
// solid the index is a harmless in-range one: its link probability q is 0).
// Five 32-bit words (20 bytes) per node -- they replace gidx (4 bytes) and 18 rank lookups per step:
//   word 0: c of row (0,+z), bits 0..29; bit 31 = centre fluid; bit 30 = "slow"
//   word 1: c of row (0,-z), bits 0..29; bit 31 = centre fluid
//   word 2: rows (+y,0) | (-y,0) << 16 as 16-bit deltas from the node's own fid
//   word 3: rows (+y,+z) | (-y,+z) << 16 as 16-bit deltas from word 0's c
//   word 4: rows (+y,-z) | (-y,-z) << 16 as 16-bit deltas from word 1's c
// A 16-bit delta is (d + 16384) in bits 0..14 and the row's centre-fluid bit in bit 15: a neighbouring row
// of the same plane starts at most one row of fluid nodes away.  "Slow" nodes -- on the periodic x or y seam,
// a delta out of range, or an index that would leave the arrays -- resolve their neighbours through the rank
// structure instead; word 2 then holds their dense index g.
constexpr uint32_t NBT_FID_MASK = 0x3fffffffu;
constexpr uint32_t NBT_FLAG = 0x40000000u;       // word 0: slow
constexpr uint32_t NBT_CENTRE_FLUID = 0x80000000u;
constexpr int NBT_DELTA_BIAS = 16384;
__host__ __device__ constexpr bool nbt_delta_fits(long long d) { return d >= -NBT_DELTA_BIAS && d < NBT_DELTA_BIAS; }
__host__ __device__ constexpr uint32_t nbt_enc16(int d, bool fluid) {
  return ((uint32_t)(d + NBT_DELTA_BIAS) & 0x7fffu) | (fluid ? 0x8000u : 0u);
}
__host__ __device__ constexpr int nbt_dec16(uint32_t h) { return (int)(h & 0x7fffu) - NBT_DELTA_BIAS; }
// Adsorbed quantity, stored compactly over the interfacial fluid nodes of the own planes.  For each group of 32
// consecutive fids: awords[fid >> 5] = { interfacial bits of the 32 nodes, first slot of the group }, slot of a
// node = first slot + popc(bits below it).  A group's slots are padded to a multiple of 4 (one 32-byte sector
// per component), so every warp reads and writes whole sectors of its own: 24 + 24 bytes per interfacial node
// and step (plus padding) instead of a sparse pass over an nfa-sized field.
// row index of a direction's (cy, cz); -1 for the node's own row
__host__ __device__ constexpr int nbt_row(int cy, int cz) {
  return cz == 0 ? (cy > 0 ? 0 : (cy < 0 ? 1 : -1))
                 : (cz > 0 ? (cy == 0 ? 2 : (cy > 0 ? 4 : 5)) : (cy == 0 ? 3 : (cy > 0 ? 6 : 7)));
}

struct MPArgs {
  Geo geo;
  const double* q;
  const double* s;
  const double* Pnow;   // 3 arrays
  double* Pnext;
  const double* Anow;   // adsorbed, 3 arrays
  double* Anext;
  long long fid_begin, fid_end;
  double ka, kd, one_minus_kd;
  int ads;
  double* partial;      // per block partial vacf (3 each)
  double* vacf_slots;   // 3 per batch step
  int batch_idx;
  int accumulate;       // add into the slot instead of overwriting (later launch of a split step)
  int check_slot;       // evaluate the convergence criterion on this (complete, global) slot, or -1
  double lim;           // 1/(2 lx ly lz / Db)
  Ctrl* ctrl;
  const uint32_t* nbt01;  // neighbour table (see NBT_*), words 0,1 and 2..4
  const uint32_t* nbt27;
  const uint32_t* rwords; // rank-lookup path: "regular" bits per group of 32 fids (neighbour ids by arithmetic); may be NULL
  const uint2* awords;    // compact adsorbed storage (see above); Anow / Anext: 3 components of stride a_stride
  long long a_stride;
  int use_nbt;            // 0: resolve neighbours through the rank structure (narrow lattices: every warp has seam nodes)
  int tpc;                // > 0: grid over the tiles, tpc consecutive tiles per CTA; 0: persistent grid, tile-stride loop
  // optional strip order (see SegTable): nseg == 0 means plain fid order over [fid_begin, fid_end)
  int nseg, ntiles;
  const int* tile_cum;          // nseg + 1: tiles before segment k
  const long long* seg_begin;   // nseg: first fid of segment k
  const long long* seg_end;     // nseg: one past the last fid of segment k
};

struct ProfileArgs {
  Geo geo;
  const double* mom;
  int axis;
  double eps;
  double* out;  // 5 per row: sum jx, jy, jz, sum rho, count(rho > eps)
};

// Tiles of BLOCK consecutive fids start on a 32-fid boundary whatever fid_begin is (a plane may start at
// any fid): every warp then reads and writes whole, aligned 256-byte runs of each array.  Threads of
// the first tile that fall before fid_begin skip (`ff < fid_begin`).
#ifndef LBG_ALIGN_TILES
#define LBG_ALIGN_TILES 1
#endif
__host__ __device__ __forceinline__ long long tile_base(long long fid_begin) {
#if LBG_ALIGN_TILES
  return fid_begin & ~31LL;
#else
  return fid_begin;
#endif
}

inline int clamp_grid(long long n, int grid) {
  const long long b = (n + 31 + BLOCK - 1) / BLOCK;  // + 31: tiles start on the 32-fid boundary below fid_begin
  return (int)(b < 1 ? 1 : (b < grid ? b : grid));
}


// launchers (geometry.cu / lb_kernels.cu / mp_kernels.cu).  Each returns the number of kernels launched.
int launch_build_bits(int plane, int nzl, const int8_t* nature_halo, uint2* words, long long nwords, cudaStream_t st);
int launch_scan_ranks(uint2* words, long long nwords, unsigned long long* total, cudaStream_t st);
int launch_build_nature(int label, int lx, int ly, int lz, int k0, int nzl, int8_t* nature_halo, cudaStream_t st);
int launch_dense_nature(const Geo& g, int8_t* out_own, cudaStream_t st);
int launch_build_gidx(const Geo& g, long long nwords, uint32_t* gidx, cudaStream_t st);
int launch_rank_at(const Geo& g, const long long* dense_idx, int n, long long ndense, long long total, long long* out,
                   cudaStream_t st);
int launch_count_interfacial(const Geo& g, long long fid_begin, long long fid_end, unsigned long long* count,
                             cudaStream_t st);
int launch_dense_interfacial(const Geo& g, int8_t* out_own, cudaStream_t st);
// dense <-> compact transfers over own planes (dense index relative to the first own plane)
int launch_scatter_to_dense(const Geo& g, const double* arr, double* dense_own, cudaStream_t st);
int launch_gather_from_dense(const Geo& g, const double* dense_own, double* arr, cudaStream_t st);
// one plane of rho, jx, jy, jz (4 arrays of the plane's size in out4), see slice_kernel
int launch_slice(const Geo& g, const double* mom, int axis, int index, double* out4, cudaStream_t st);
// n(t)(., l) pulled from the post-collision populations fin into the dense own-plane order
int launch_pull_to_dense(const Geo& g, const double* fin, int l, double* dense_own, cudaStream_t st);
int launch_scatter3_to_dense_aos(const Geo& g, const double* soa3, double* dense_aos_own, cudaStream_t st);
// compact adsorbed storage: group words {interfacial bits, padded count} for fids [0, nfa), then launch_scan_ranks
// turns the counts into first slots; read-back into the reference's AoS order (awords == NULL: all zero)
int launch_build_awords(const Geo& g, long long fid_begin, long long fid_end, uint2* awords, cudaStream_t st);
int launch_scatter3_compact_to_dense_aos(const Geo& g, const uint2* awords, const double* a3, long long a_stride,
                                         double* dense_aos_own, cudaStream_t st);

int launch_lb_init(const Geo& g, long long fid_begin, long long fid_end, double rho0, const double a0[3], double* f,
                   double* mom, cudaStream_t st);
int launch_collide(const CollideArgs& a, bool tau1, int fmode, int grid, cudaStream_t st);
int launch_lb_step(const LBArgs& a, bool tau1, int fmode, bool check, bool writej, int minb, int grid,
                   cudaStream_t st);
int launch_moments(const MomArgs& a, int fmode, int grid, cudaStream_t st);
// in-place (AA) variant, lb_aa_kernels.cu
int launch_aa_step(const LBArgs& a, bool tau1, int fmode, bool swapped, bool first, const double* mom, int grid,
                   cudaStream_t st);
int launch_aa_moments(const LBArgs& a, int fmode, bool swapped, bool check, bool writej, double* mom, double* pops,
                      int grid, cudaStream_t st);
int occupancy_grid_aa(int sm_count);
int launch_fill_force(const Geo& g, long long nf, const double f[3], double* field, cudaStream_t st);
int launch_profile(const ProfileArgs& a, int rows, cudaStream_t st);
int launch_mp_init(const MPInitArgs& a, int grid, cudaStream_t st);
int launch_mp_step(const MPArgs& a, int grid, cudaStream_t st);
int occupancy_grid_lb(int sm_count, int minb);
int occupancy_grid_mp(int sm_count);

}  // namespace lbg
