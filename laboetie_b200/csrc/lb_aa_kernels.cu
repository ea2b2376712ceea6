// lb_aa_kernels.cu -- Phase A in place (AA pattern): one population buffer instead of two.
//
// The single buffer alternates between two layouts (t = completed reference steps):
//   N(t): slot l of node r holds n(t)(r,l), the reference's own state after step t
//         (post-streaming, pre-collision) -- what lbg_lb_init / lbg_lb_upload provide at t = 0;
//   S(t): slot inv(l) of node r holds n*(t)(r,l), the post-collision population of step t that has
//         not been streamed yet.
// Step t+1 from N(t)  ("even" kernel, purely local):  moments of n(t), collide -> n*(t+1), each value
//   written back into the node's own opposite slot.                              N(t)   -> S(t+1)
// Step t+2 from S(t+1) ("odd" kernel):  pull n(t+1)(r,l) from slot inv(l) of r-c_l (bounce-back: slot l of
//   r itself), moments, collide -> n*(t+2)(r,l), pushed into slot l of r+c_l (bounce-back: slot inv(l) of
//   r).  Every slot is read and written by the same thread, so the update is race-free in place.
//                                                                                S(t+1) -> N(t+2)
// Either way a step is one pass of 19 loads + 19 stores per fluid node.  What the in-place scheme cannot
// give for free is the reference's per-step convergence scalar: the momentum of the state a kernel
// leaves behind is only computed by the *next* kernel, after it has already collided.  On checked steps
// a separate moments pass (aa_moments_kernel: 19 loads, j_old read, j write) therefore evaluates
// rho, j, max|j - j_old| and ANY(n<0) on the new state before the next step is allowed to start; the
// exit step and all results stay exactly the reference's, at 504 instead of 352 bytes per node on
// checked steps.  Arithmetic is the shared lb_node.cuh, so results are bit-identical to the
// two-lattice kernels.
#include "lb_node.cuh"

namespace lbg {
using namespace d3q19;

namespace {

__device__ __forceinline__ bool aa_stop_check(const LBArgs& a) {
  int stop = *(volatile int*)&a.ctrl->stop;
  // the flag of step i-1 comes from a moments pass (checked steps), that of step i-2 from kernel i-1 (aa_flag_negative)
  // (across slabs a flag only counts once it is global: neg_flag_local 1 = neither is, 2 = only that of step i-1)
  if (!stop && ((a.batch_idx > 0 && a.neg_flag_local != 1 &&
                 *(volatile unsigned long long*)&a.l2_slots[2 * (a.batch_idx - 1) + 1] != 0ull) ||
                (a.batch_idx > 1 && a.neg_flag_local == 0 &&
                 *(volatile unsigned long long*)&a.l2_slots[2 * (a.batch_idx - 2) + 1] != 0ull))) {
    a.ctrl->stop = 1;  // equilibration.f90:248
    stop = 1;
  }
  if (!stop && a.prev_checked && a.prev_may_stop) {
    const double prev =
        __longlong_as_double((long long)*(volatile unsigned long long*)&a.l2_slots[2 * (a.batch_idx - 1)]);
    if (prev <= a.target) {  // equilibration.f90:346
      a.ctrl->stop = 1;
      a.ctrl->stop_idx = a.batch_idx;
      stop = 1;
    }
  }
  return stop != 0;
}

// ANY(n<0) of equilibration.f90:248 without a separate pass: the populations a step kernel reads ARE the state
// the previous step left behind, so kernel i raises the flag of step i-1 (the first step of a batch looks at a
// state the previous batch's closing moments pass -- or lbg_lb_init -- has already vetted).  The host then
// reports exactly the reference's stopping step; rare, so a plain atomic per offending thread.
__device__ __forceinline__ void aa_flag_negative(const LBArgs& a) {
  if (a.batch_idx > 0) atomicMax(&a.l2_slots[2 * (a.batch_idx - 1) + 1], 1ull);
}

template <int FMODE>
__device__ __forceinline__ void load_forces(const LBArgs& a, int fid, double (&fj)[3], double (&fc)[3]) {
  const long long nfa = a.geo.nfa;
  fj[0] = fj[1] = fj[2] = fc[0] = fc[1] = fc[2] = 0.0;
  if constexpr (FMODE == FORCE_UNIFORM) {
    for (int d = 0; d < 3; ++d) {
      fj[d] = a.fj[d];
      fc[d] = a.fc[d];
    }
  } else if constexpr (FMODE == FORCE_FIELD) {
    for (int d = 0; d < 3; ++d) {
      fj[d] = a.fj_field[d * nfa + fid];
      fc[d] = a.fc_field[d * nfa + fid];
    }
  }
}

// N(t) -> S(t+1).  FIRST: the step right after init / upload takes density and momentum from the
// driver's arrays (equilibration.f90:59,75-80) instead of the populations.
template <bool TAU1, int FMODE, bool FIRST>
__global__ void __launch_bounds__(BLOCK, 2) aa_even_kernel(const __grid_constant__ LBArgs a, const double* __restrict__ mom) {
  __shared__ int s_stop;
  if (threadIdx.x == 0) s_stop = aa_stop_check(a) ? 1 : 0;
  __syncthreads();
  if (s_stop) return;
  const long long nfa = a.geo.nfa;
  double* f = a.fout;  // in place: fin == fout
  __shared__ unsigned int s_slot[2];
  Tiles ts;
  ts.init(a.geo, a.fid_begin, a.fid_end, &a.ctrl->tile_next, s_slot);
  const long long base = tile_base(a.fid_begin) + threadIdx.x;
  while (ts.tile >= 0) {
    const long long ff = base + (long long)ts.tile * BLOCK;
    ts.advance();
    if (ff < a.fid_begin || ff >= a.fid_end) continue;
    const int fid = (int)ff;
    double n[NV];
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      n[L] = __ldcg(f + (long long)L * nfa + fid);
    });
    double fj[3], fc[3];
    load_forces<FMODE>(a, fid, fj, fc);
    double rho, jx, jy, jz;
    bool neg = false;
    if constexpr (FIRST) {
      rho = mom[fid];
      jx = mom[nfa + fid];
      jy = mom[2 * nfa + fid];
      jz = mom[3 * nfa + fid];
    } else {
      moments(n, fj[0] / 2.0, fj[1] / 2.0, fj[2] / 2.0, rho, jx, jy, jz, neg);
    }
    if (neg) aa_flag_negative(a);
    collide<TAU1, FMODE != FORCE_NONE>(n, a.k, rho, jx, jy, jz, fc[0], fc[1], fc[2], a.w1, a.w2, a.w3);
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      f[(long long)inv(L) * nfa + fid] = n[L];
    });
  }
  if (a.geo.tpc > 0) cta_checks_in_last(a.ctrl);
}

// S(t) -> N(t+1)
template <bool TAU1, int FMODE>
__global__ void __launch_bounds__(BLOCK, 2) aa_odd_kernel(const __grid_constant__ LBArgs a) {
  __shared__ int s_stop;
  if (threadIdx.x == 0) s_stop = aa_stop_check(a) ? 1 : 0;
  __syncthreads();
  if (s_stop) return;
  const Geo& geo = a.geo;
  const long long nfa = geo.nfa;
  double* f = a.fout;
  __shared__ unsigned int s_slot[2];
  Tiles ts;
  ts.init(geo, a.fid_begin, a.fid_end, &a.ctrl->tile_next, s_slot);
  const long long base = tile_base(a.fid_begin) + threadIdx.x;
  long long ff_next = ts.tile >= 0 ? base + (long long)ts.tile * BLOCK : -1;
  uint32_t gi_next = (ff_next >= 0 && ff_next < a.fid_end) ? __ldg(geo.gidx + ff_next) : 0u;
  while (ts.tile >= 0) {
    const long long ff = ff_next;
    const uint32_t gi = gi_next;
    const int tn = ts.next_tile();
    ff_next = tn >= 0 ? base + (long long)tn * BLOCK : -1;
    if (ff_next >= 0 && ff_next < a.fid_end) gi_next = __ldg(geo.gidx + ff_next);
    ts.advance();
    if (ff < a.fid_begin || ff >= a.fid_end) continue;
    const int fid = (int)ff;
    const int g = (int)(gi & GIDX_MASK);
    const Nb nb = neighbours(geo, g);
    // neighbour r + c_L: fluid? and its fluid id (the slot both read for direction inv(L) and written for L)
    int nfid[NV];
    uint32_t fl = 0;
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      int t;
      if (lookup(geo, g + offset_plus<L>(nb), t)) fl |= 1u << L;
      nfid[L] = t;
    });
    double n[NV];
    n[0] = __ldcg(f + fid);
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      // n(r,L) = n*(r - c_L, L), kept in slot inv(L) of r - c_L = r + c_inv(L); bounce-back: slot L of r
      const bool src_fluid = (fl >> inv(L)) & 1u;
      const int idx = src_fluid ? nfid[inv(L)] : fid;
      const int arr = src_fluid ? inv(L) : L;
      n[L] = __ldcg(f + (long long)arr * nfa + idx);
    });
    double fj[3], fc[3];
    load_forces<FMODE>(a, fid, fj, fc);
    double rho, jx, jy, jz;
    bool neg;
    moments(n, fj[0] / 2.0, fj[1] / 2.0, fj[2] / 2.0, rho, jx, jy, jz, neg);
    if (neg) aa_flag_negative(a);
    collide<TAU1, FMODE != FORCE_NONE>(n, a.k, rho, jx, jy, jz, fc[0], fc[1], fc[2], a.w1, a.w2, a.w3);
    f[fid] = n[0];
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      // push n*(r,L) to slot L of r + c_L; bounce-back: slot inv(L) of r
      const bool dst_fluid = (fl >> L) & 1u;
      const int idx = dst_fluid ? nfid[L] : fid;
      const int arr = dst_fluid ? L : inv(L);
      f[(long long)arr * nfa + idx] = n[L];
    });
  }
  if (geo.tpc > 0) cta_checks_in_last(a.ctrl);
}

// Reference state of the current step from either layout: n(t)(r,·), then density, momentum, the
// negativity guard and max|j - j_old| (equilibration.f90:248-300,339-343).  Optional outputs: j for the
// next check, the driver's rho/j arrays, the populations in normal layout.
template <int FMODE, bool SWAPPED>
__global__ void __launch_bounds__(BLOCK, 2) aa_moments_kernel(const __grid_constant__ LBArgs a, int check, int writej,
                                                               double* __restrict__ mom, double* __restrict__ pops) {
  __shared__ double s_red[BLOCK / 32];
  __shared__ int s_neg;
  __shared__ int s_stop;
  if (threadIdx.x == 0) {
    s_neg = 0;
    s_stop = *(volatile int*)&a.ctrl->stop;  // a step that did not run leaves nothing to evaluate
  }
  __syncthreads();
  if (s_stop) return;
  const Geo& geo = a.geo;
  const long long nfa = geo.nfa;
  const double* f = a.fin;
  double dmax = 0.0;
  bool any_neg = false;
  for (long long ff = first_fid(a.fid_begin); ff < a.fid_end; ff += (long long)gridDim.x * BLOCK) {
    if (ff < a.fid_begin) continue;
    const int fid = (int)ff;
    double n[NV];
    if constexpr (!SWAPPED) {
      static_for<0, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        n[L] = __ldcg(f + (long long)L * nfa + fid);
      });
    } else {
      const int g = (int)(geo.gidx[fid] & GIDX_MASK);
      const Nb nb = neighbours(geo, g);
      n[0] = __ldcg(f + fid);
      static_for<1, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        int fsrc;
        const bool src_fluid = lookup(geo, g + offset_plus<inv(L)>(nb), fsrc);
        const int idx = src_fluid ? fsrc : fid;
        const int arr = src_fluid ? inv(L) : L;
        n[L] = __ldcg(f + (long long)arr * nfa + idx);
      });
    }
    double fjx = 0, fjy = 0, fjz = 0;
    if constexpr (FMODE == FORCE_UNIFORM) {
      fjx = a.fj[0]; fjy = a.fj[1]; fjz = a.fj[2];
    } else if constexpr (FMODE == FORCE_FIELD) {
      fjx = a.fj_field[fid]; fjy = a.fj_field[nfa + fid]; fjz = a.fj_field[2 * nfa + fid];
    }
    double rho, jx, jy, jz;
    bool neg;
    moments(n, fjx / 2.0, fjy / 2.0, fjz / 2.0, rho, jx, jy, jz, neg);
    any_neg |= neg;
    if (check) {
      const double ox = a.jold[fid], oy = a.jold[nfa + fid], oz = a.jold[2 * nfa + fid];
      dmax = fmax(dmax, fmax(fabs(jx - ox), fmax(fabs(jy - oy), fabs(jz - oz))));
    }
    if (writej) {
      a.jnew[fid] = jx;
      a.jnew[nfa + fid] = jy;
      a.jnew[2 * nfa + fid] = jz;
    }
    if (mom) {
      mom[fid] = rho;
      mom[nfa + fid] = jx;
      mom[2 * nfa + fid] = jy;
      mom[3 * nfa + fid] = jz;
    }
    if (pops) {
      static_for<0, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        pops[(long long)L * nfa + fid] = n[L];
      });
    }
  }
  if (any_neg) s_neg = 1;
  dmax = warp_max(dmax);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dmax;
  __syncthreads();
  if (threadIdx.x == 0 && a.l2_slots) {
    if (check) {
      double v = s_red[0];
#pragma unroll
      for (int w = 1; w < BLOCK / 32; ++w) v = fmax(v, s_red[w]);
      atomicMax(&a.l2_slots[2 * a.batch_idx], (unsigned long long)__double_as_longlong(v));
    }
    if (s_neg) atomicMax(&a.l2_slots[2 * a.batch_idx + 1], 1ull);
  }
}

}  // namespace

// One in-place step: `swapped` is the layout the buffer is in before the step.
int launch_aa_step(const LBArgs& a, bool tau1, int fmode, bool swapped, bool first, const double* mom, int grid,
                   cudaStream_t st) {
  if (a.fid_end <= a.fid_begin) return 0;
  const int gr = clamp_grid(a.fid_end - a.fid_begin, grid);
#define LBG_AA_EVEN(T, F)                                                          \
  do {                                                                             \
    if (first) aa_even_kernel<T, F, true><<<gr, BLOCK, 0, st>>>(a, mom);           \
    else aa_even_kernel<T, F, false><<<gr, BLOCK, 0, st>>>(a, mom);                \
  } while (0)
#define LBG_AA_DISPATCH(T)                                                         \
  do {                                                                             \
    if (!swapped) {                                                                \
      if (fmode == FORCE_NONE) LBG_AA_EVEN(T, FORCE_NONE);                         \
      else if (fmode == FORCE_UNIFORM) LBG_AA_EVEN(T, FORCE_UNIFORM);              \
      else LBG_AA_EVEN(T, FORCE_FIELD);                                            \
    } else {                                                                       \
      if (fmode == FORCE_NONE) aa_odd_kernel<T, FORCE_NONE><<<gr, BLOCK, 0, st>>>(a);          \
      else if (fmode == FORCE_UNIFORM) aa_odd_kernel<T, FORCE_UNIFORM><<<gr, BLOCK, 0, st>>>(a); \
      else aa_odd_kernel<T, FORCE_FIELD><<<gr, BLOCK, 0, st>>>(a);                 \
    }                                                                              \
  } while (0)
  if (tau1) LBG_AA_DISPATCH(true);
  else LBG_AA_DISPATCH(false);
#undef LBG_AA_DISPATCH
#undef LBG_AA_EVEN
  return 1;
}

int launch_aa_moments(const LBArgs& a, int fmode, bool swapped, bool check, bool writej, double* mom, double* pops,
                      int grid, cudaStream_t st) {
  if (a.fid_end <= a.fid_begin) return 0;
  const int gr = clamp_grid(a.fid_end - a.fid_begin, grid);
#define LBG_AA_MOM(F)                                                                                         \
  do {                                                                                                        \
    if (swapped) aa_moments_kernel<F, true><<<gr, BLOCK, 0, st>>>(a, check ? 1 : 0, writej ? 1 : 0, mom, pops); \
    else aa_moments_kernel<F, false><<<gr, BLOCK, 0, st>>>(a, check ? 1 : 0, writej ? 1 : 0, mom, pops);       \
  } while (0)
  if (fmode == FORCE_NONE) LBG_AA_MOM(FORCE_NONE);
  else if (fmode == FORCE_UNIFORM) LBG_AA_MOM(FORCE_UNIFORM);
  else LBG_AA_MOM(FORCE_FIELD);
#undef LBG_AA_MOM
  return 1;
}

int occupancy_grid_aa(int sm_count) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, aa_odd_kernel<true, FORCE_UNIFORM>, BLOCK, 0);
  if (per_sm < 1) per_sm = 1;
  return sm_count * per_sm;
}

}  // namespace lbg
