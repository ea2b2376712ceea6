// lb_kernels.cu -- Phase-A kernels for sm_100a: initial state, the fused pull stream + moments +
// collide step, moments/population read-back, plane profiles.
//
// Data layout in HBM: fluid-compacted structure of arrays, fp64 (lbg_internal.h).  Population l of
// fluid node fid lives at f[l*nfa + fid]; fids follow the reference's node order (x fastest), so a
// warp's 32 consecutive fids are 32 consecutive fluid nodes of a row and every own-node access of a
// warp is one aligned, fully used 256-byte run.  One thread owns one fluid node.
//
// The step kernel K(t) implements, for fluid node r (SURVEY 8a "A2 o A3"):
//   n(t)(r,l)   = n*(t)(r-c_l, l)   if r-c_l is fluid          [streaming, equilibration.f90:227-243]
//               = n*(t)(r, inv l)    otherwise                  [bounce-back, equilibration.f90:204-222]
//   rho(t), j(t), ANY(n<0), max|j(t)-j(t-1)|                    [equilibration.f90:248-300,339-343]
//   n*(t+1)     = collide(n(t), rho(t), j(t), f)                [module_collision.f90:77-108]
// i.e. the reference's stream of step t fused with its collide of step t+1.
// n* is written to the other buffer (two-lattice), so a step can be redone when
// the driver changes the force after a convergence event.
#include "lb_node.cuh"

#ifndef LBG_PULL_FENCE
#define LBG_PULL_FENCE 0  // measured neutral to -1 % (profiles/variants_r3.txt): the kernel runs at the HBM rate of its access shape either way
#endif
#ifndef LBG_LOADMODE
#define LBG_LOADMODE 1  // population loads: 0 default, 1 .cg (L2 only), 2 .cs (streaming); measured in profiles/
#endif

namespace lbg {
using namespace d3q19;

namespace {

__device__ __forceinline__ double ld_pop(const double* p) {
#if LBG_LOADMODE == 1
  return __ldcg(p);
#elif LBG_LOADMODE == 2
  return __ldcs(p);
#else
  return *p;
#endif
}

// n(t)(r,·) by pull with halfway bounce-back.  The source of direction L is node r - c_L = r + c_inv(L);
// if it is solid the population comes back from the node's own opposite slot.
// Two phases: first all 18 rank lookups are resolved into (array, index) pairs, then the 19 population
// loads are issued back to back.  A warp has few scoreboards: a lookup word that is consumed after the
// first population loads were issued would wait for those DRAM loads too (one exposed latency more).
__device__ __forceinline__ void pull(const Geo& geo, const double* __restrict__ fin, int fid, int g, const Nb& nb,
                                     double (&n)[NV]) {
  const long long nfa = geo.nfa;
  int idx[NV];
  uint32_t fluid = 0;
  static_for<1, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    int fsrc;
    const bool src_fluid = lookup(geo, g + offset_plus<inv(L)>(nb), fsrc);
    idx[L] = src_fluid ? fsrc : fid;
    fluid |= (src_fluid ? 1u : 0u) << L;
  });
#if LBG_PULL_FENCE
  // make every population address depend on every lookup (geo.zero is 0 at run time), so that the
  // scheduler cannot sink a lookup below the first population load
  uint32_t any = fluid;
  static_for<1, NV>([&](auto Lc) { any |= (uint32_t)idx[decltype(Lc)::value]; });
  fin += (any & (uint32_t)geo.zero);
#endif
  n[0] = ld_pop(fin + fid);
  static_for<1, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    const int arr = ((fluid >> L) & 1u) ? L : inv(L);
    n[L] = ld_pop(fin + (long long)arr * nfa + idx[L]);
  });
}

// ---------------------------------------------------------------------------
// MINB = resident blocks per SM the register allocation aims at.
template <bool TAU1, int FMODE, bool CHECK, bool WRITEJ, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) lb_step_kernel(const __grid_constant__ LBArgs a) {
  __shared__ int s_stop;
  __shared__ double s_red[BLOCK / 32];
  __shared__ int s_neg;
  if (threadIdx.x == 0) {
    // slot pair of a step: [2i] = l2err bits, [2i+1] = non-zero if a population went negative
    int stop = *(volatile int*)&a.ctrl->stop;
    if (!stop && a.batch_idx > 0 && !a.neg_flag_local && *(volatile unsigned long long*)&a.l2_slots[2 * (a.batch_idx - 1) + 1] != 0ull) {
      a.ctrl->stop = 1;  // equilibration.f90:248: the reference stops at the step with a negative population
      stop = 1;
    }
    if (!stop && a.prev_checked && a.prev_may_stop) {
      const double prev =
          __longlong_as_double((long long)*(volatile unsigned long long*)&a.l2_slots[2 * (a.batch_idx - 1)]);
      if (prev <= a.target) {  // equilibration.f90:346
        a.ctrl->stop = 1;
        a.ctrl->stop_idx = a.batch_idx;  // 1 + index of the converged step
        stop = 1;
      }
    }
    s_stop = stop;
    s_neg = 0;
  }
  __syncthreads();
  if (s_stop) return;

  const Geo& geo = a.geo;
  const long long nfa = geo.nfa;
  double dmax = 0.0;
  bool any_neg = false;
  // tiles of BLOCK consecutive fluid nodes (schedule: lattice.cuh Tiles); the dense index of the next tile's
  // node is fetched one iteration ahead so that its DRAM latency does not sit in front of the dependent
  // population loads
  // Strip order (a.nseg > 0): the planes are cut into strips of rows and all planes of a strip are visited
  // before the next strip.  A wall node reads slot inv(l) of ITS OWN fid (bounce-back) while the other nodes of
  // that 32-byte sector are read, shifted, from the plane above or below; in plain order the two reads are a whole
  // plane of traffic apart and the sector comes from HBM twice; in strip order they are one (strip, plane)
  // segment apart and the second read hits L2.
  __shared__ unsigned int s_slot[2];
  Tiles ts;
  ts.init(geo, a.fid_begin, a.fid_end, &a.ctrl->tile_next, s_slot, a.nseg > 0 ? a.ntiles : -1);
  const long long base = tile_base(a.fid_begin) + threadIdx.x;
  int seg = 0;
  auto node_of = [&](int tile) -> long long {  // fid of this thread in `tile`, or -1 (tiles come in increasing order)
    if (tile < 0) return -1;
    if (a.nseg == 0) {
      const long long f = base + (long long)tile * BLOCK;
      return (f >= a.fid_begin && f < a.fid_end) ? f : -1;
    }
    while (tile >= a.tile_cum[seg + 1]) ++seg;
    const long long sb = a.seg_begin[seg];
    const long long f = tile_base(sb) + (long long)(tile - a.tile_cum[seg]) * BLOCK + threadIdx.x;
    return (f >= sb && f < a.seg_end[seg]) ? f : -1;
  };
  long long ff_next = node_of(ts.tile);
  uint32_t gi_next = ff_next >= 0 ? __ldg(geo.gidx + ff_next) : 0u;
  while (ts.tile >= 0) {
    const long long ff = ff_next;
    const uint32_t gi = gi_next;
    ff_next = node_of(ts.next_tile());
    if (ff_next >= 0) gi_next = __ldg(geo.gidx + ff_next);
    ts.advance();
    if (ff < 0) continue;
    const int fid = (int)ff;
    const int g = (int)(gi & GIDX_MASK);
    const Nb nb = neighbours(geo, g);
    double n[NV];
    pull(geo, a.fin, fid, g, nb, n);
    double fjx = 0, fjy = 0, fjz = 0, fcx = 0, fcy = 0, fcz = 0;
    if constexpr (FMODE == FORCE_UNIFORM) {
      fjx = a.fj[0]; fjy = a.fj[1]; fjz = a.fj[2];
      fcx = a.fc[0]; fcy = a.fc[1]; fcz = a.fc[2];
    } else if constexpr (FMODE == FORCE_FIELD) {
      fjx = a.fj_field[fid]; fjy = a.fj_field[nfa + fid]; fjz = a.fj_field[2 * nfa + fid];
      fcx = a.fc_field[fid]; fcy = a.fc_field[nfa + fid]; fcz = a.fc_field[2 * nfa + fid];
    }
    double rho, jx, jy, jz;
    bool neg;
    moments(n, fjx / 2.0, fjy / 2.0, fjz / 2.0, rho, jx, jy, jz, neg);
    any_neg |= neg;
    if constexpr (CHECK) {
      const double ox = a.jold[fid], oy = a.jold[nfa + fid], oz = a.jold[2 * nfa + fid];
      dmax = fmax(dmax, fmax(fabs(jx - ox), fmax(fabs(jy - oy), fabs(jz - oz))));
    }
    if constexpr (WRITEJ) {
      a.jnew[fid] = jx;
      a.jnew[nfa + fid] = jy;
      a.jnew[2 * nfa + fid] = jz;
    }
    collide<TAU1, FMODE != FORCE_NONE>(n, a.k, rho, jx, jy, jz, fcx, fcy, fcz, a.w1, a.w2, a.w3);
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      a.fout[(long long)L * nfa + fid] = n[L];
    });
  }
  if (any_neg) s_neg = 1;
  if constexpr (CHECK) {
    dmax = warp_max(dmax);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if constexpr (CHECK) {
      double v = s_red[0];
#pragma unroll
      for (int w = 1; w < BLOCK / 32; ++w) v = fmax(v, s_red[w]);
      // non-negative doubles order like their bit patterns
      atomicMax(&a.l2_slots[2 * a.batch_idx], (unsigned long long)__double_as_longlong(v));
    }
    if (s_neg) atomicMax(&a.l2_slots[2 * a.batch_idx + 1], 1ull);
  }
  if (geo.tpc > 0) cta_checks_in_last(a.ctrl);
}

// ---------------------------------------------------------------------------
// Two-stage software pipeline of the step kernel.  The plain kernel above serialises, per warp and tile,
// [18 rank lookups (an L2 round trip)] -> [19 population loads (a DRAM round trip; two when the register
// allocator interleaves late lookups with the first loads, which then share a scoreboard)] -> arithmetic ->
// stores; ncu's source view puts 35 % of its stall samples on the lookup words and 22 % on the populations
// (profiles/r4f_lb_source_stalls.txt).  Here the lookups of tile i+1 are resolved while the population loads
// of tile i are in flight, and handed to the next iteration through 18 packed words per thread in shared
// memory (source fid | source-is-fluid << 31; each thread reads only its own column: no barrier).  An
// iteration then starts with 18 LDS and issues all 22 loads of its tile back to back: one exposed round trip.
// The dense index (gidx) is fetched two tiles ahead.  Static tile-stride schedule.
constexpr uint32_t SRC_FLUID = 0x80000000u;

template <bool TAU1, int FMODE, bool CHECK, bool WRITEJ, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) lb_step_pipe_kernel(const __grid_constant__ LBArgs a) {
  __shared__ uint32_t s_src[NV - 1][BLOCK];
  __shared__ int s_stop;
  __shared__ double s_red[BLOCK / 32];
  __shared__ int s_neg;
  if (threadIdx.x == 0) {
    int stop = *(volatile int*)&a.ctrl->stop;
    if (!stop && a.batch_idx > 0 && !a.neg_flag_local && *(volatile unsigned long long*)&a.l2_slots[2 * (a.batch_idx - 1) + 1] != 0ull) {
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
    s_stop = stop;
    s_neg = 0;
  }
  __syncthreads();
  if (s_stop) return;

  const Geo& geo = a.geo;
  const long long nfa = geo.nfa;
  double dmax = 0.0;
  bool any_neg = false;
  const long long stride = (long long)gridDim.x * BLOCK;
  const long long first = first_fid(a.fid_begin);
  auto valid = [&](long long ff) { return ff >= a.fid_begin && ff < a.fid_end; };
  // resolve the 18 pull sources of node (fid, g) and park them in this thread's column of s_src
  auto resolve = [&](int fid, int g) {
    const Nb nb = neighbours(geo, g);
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      int fsrc;
      const bool src_fluid = lookup(geo, g + offset_plus<inv(L)>(nb), fsrc);
      s_src[L - 1][threadIdx.x] = src_fluid ? ((uint32_t)fsrc | SRC_FLUID) : (uint32_t)fid;
    });
  };
  // prologue: sources of the first tile, dense index of the second
  if (valid(first)) resolve((int)first, (int)(__ldg(geo.gidx + first) & GIDX_MASK));
  uint32_t gi1 = valid(first + stride) ? __ldg(geo.gidx + first + stride) : 0u;  // tile i+1
  for (long long ff = first; ff < a.fid_end; ff += stride) {
    const bool ok = valid(ff);
    const int fid = (int)ff;
    double n[NV];
    double ox = 0, oy = 0, oz = 0;
    if (ok) {
      uint32_t src[NV - 1];
      static_for<1, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        src[L - 1] = s_src[L - 1][threadIdx.x];
      });
      n[0] = ld_pop(a.fin + fid);
      static_for<1, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        const uint32_t e = src[L - 1];
        const int arr = (e & SRC_FLUID) ? L : inv(L);
        n[L] = ld_pop(a.fin + (long long)arr * nfa + (long long)(e & ~SRC_FLUID));
      });
      if constexpr (CHECK) {
        ox = a.jold[fid];
        oy = a.jold[nfa + fid];
        oz = a.jold[2 * nfa + fid];
      }
    }
    // while those loads are in flight: dense index of tile i+2, pull sources of tile i+1
    const uint32_t gi = gi1;
    if (valid(ff + 2 * stride)) gi1 = __ldg(geo.gidx + ff + 2 * stride);
    if (valid(ff + stride)) resolve((int)(ff + stride), (int)(gi & GIDX_MASK));
    if (!ok) continue;
    double fjx = 0, fjy = 0, fjz = 0, fcx = 0, fcy = 0, fcz = 0;
    if constexpr (FMODE == FORCE_UNIFORM) {
      fjx = a.fj[0]; fjy = a.fj[1]; fjz = a.fj[2];
      fcx = a.fc[0]; fcy = a.fc[1]; fcz = a.fc[2];
    } else if constexpr (FMODE == FORCE_FIELD) {
      fjx = a.fj_field[fid]; fjy = a.fj_field[nfa + fid]; fjz = a.fj_field[2 * nfa + fid];
      fcx = a.fc_field[fid]; fcy = a.fc_field[nfa + fid]; fcz = a.fc_field[2 * nfa + fid];
    }
    double rho, jx, jy, jz;
    bool neg;
    moments(n, fjx / 2.0, fjy / 2.0, fjz / 2.0, rho, jx, jy, jz, neg);
    any_neg |= neg;
    if constexpr (CHECK) dmax = fmax(dmax, fmax(fabs(jx - ox), fmax(fabs(jy - oy), fabs(jz - oz))));
    if constexpr (WRITEJ) {
      a.jnew[fid] = jx;
      a.jnew[nfa + fid] = jy;
      a.jnew[2 * nfa + fid] = jz;
    }
    collide<TAU1, FMODE != FORCE_NONE>(n, a.k, rho, jx, jy, jz, fcx, fcy, fcz, a.w1, a.w2, a.w3);
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      a.fout[(long long)L * nfa + fid] = n[L];
    });
  }
  if (any_neg) s_neg = 1;
  if constexpr (CHECK) {
    dmax = warp_max(dmax);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if constexpr (CHECK) {
      double v = s_red[0];
#pragma unroll
      for (int w = 1; w < BLOCK / 32; ++w) v = fmax(v, s_red[w]);
      atomicMax(&a.l2_slots[2 * a.batch_idx], (unsigned long long)__double_as_longlong(v));
    }
    if (s_neg) atomicMax(&a.l2_slots[2 * a.batch_idx + 1], 1ull);
  }
}

template <bool TAU1, int FMODE>
__global__ void __launch_bounds__(BLOCK) collide_kernel(const __grid_constant__ CollideArgs a) {
  const long long nfa = a.geo.nfa;
  for (long long ff = first_fid(a.fid_begin); ff < a.fid_end; ff += (long long)gridDim.x * BLOCK) {
    if (ff < a.fid_begin) continue;
    const int fid = (int)ff;
    double n[NV];
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      n[L] = a.fin[(long long)L * nfa + fid];
    });
    double fcx = 0, fcy = 0, fcz = 0;
    if constexpr (FMODE == FORCE_UNIFORM) {
      fcx = a.fc[0]; fcy = a.fc[1]; fcz = a.fc[2];
    } else if constexpr (FMODE == FORCE_FIELD) {
      fcx = a.fc_field[fid]; fcy = a.fc_field[nfa + fid]; fcz = a.fc_field[2 * nfa + fid];
    }
    collide<TAU1, FMODE != FORCE_NONE>(n, a.k, a.mom[fid], a.mom[nfa + fid], a.mom[2 * nfa + fid], a.mom[3 * nfa + fid],
                                       fcx, fcy, fcz, a.w1, a.w2, a.w3);
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      a.fout[(long long)L * nfa + fid] = n[L];
    });
  }
}

template <int FMODE>
__global__ void __launch_bounds__(BLOCK) moments_kernel(const __grid_constant__ MomArgs a) {
  const long long nfa = a.geo.nfa;
  for (long long ff = first_fid(a.fid_begin); ff < a.fid_end; ff += (long long)gridDim.x * BLOCK) {
    if (ff < a.fid_begin) continue;
    const int fid = (int)ff;
    const int g = (int)(a.geo.gidx[fid] & GIDX_MASK);
    const Nb nb = neighbours(a.geo, g);
    double n[NV];
    pull(a.geo, a.fin, fid, g, nb, n);
    double fjx = 0, fjy = 0, fjz = 0;
    if constexpr (FMODE == FORCE_UNIFORM) {
      fjx = a.fj[0]; fjy = a.fj[1]; fjz = a.fj[2];
    } else if constexpr (FMODE == FORCE_FIELD) {
      fjx = a.fj_field[fid]; fjy = a.fj_field[nfa + fid]; fjz = a.fj_field[2 * nfa + fid];
    }
    double rho, jx, jy, jz;
    bool neg;
    moments(n, fjx / 2.0, fjy / 2.0, fjz / 2.0, rho, jx, jy, jz, neg);
    if (a.mom) {
      a.mom[fid] = rho;
      a.mom[nfa + fid] = jx;
      a.mom[2 * nfa + fid] = jy;
      a.mom[3 * nfa + fid] = jz;
    }
    if (a.pops) {
      static_for<0, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        a.pops[(long long)L * nfa + fid] = n[L];
      });
    }
  }
}

// init_simu.f90:24-39 and equilibration.f90:59,75-80 (solid nodes hold 0 and own no storage)
__global__ void __launch_bounds__(BLOCK) lb_init_kernel(long long nfa, long long fid_begin, long long fid_end, double rho0,
                                                        double a00, double a01, double a02, double* __restrict__ f,
                                                        double* __restrict__ mom) {
  const double a0[3] = {a00, a01, a02};
  for (long long fid = fid_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; fid < fid_end;
       fid += (long long)gridDim.x * BLOCK) {
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      f[(long long)L * nfa + fid] = rho0 * a0[kind(L)];
    });
    mom[fid] = rho0;
    mom[nfa + fid] = 0.0;
    mom[2 * nfa + fid] = 0.0;
    mom[3 * nfa + fid] = 0.0;
  }
}

// equilibration.f90:381-386 materialised as a field (used when a uniform force
// and a force field meet across a force change).
__global__ void __launch_bounds__(BLOCK) fill_force_kernel(long long nfa, long long nf, double fx, double fy, double fz,
                                                           double* __restrict__ field) {
  for (long long fid = (long long)blockIdx.x * BLOCK + threadIdx.x; fid < nf; fid += (long long)gridDim.x * BLOCK) {
    field[fid] = fx;
    field[nfa + fid] = fy;
    field[2 * nfa + fid] = fz;
  }
}

// equilibration.f90:161-172: one block per row along `axis`; fixed-order tree, so deterministic.
__global__ void __launch_bounds__(BLOCK) profile_kernel(const __grid_constant__ ProfileArgs a) {
  const Geo& geo = a.geo;
  const int p = blockIdx.x;
  const int n1 = a.axis == 0 ? geo.ly : geo.lx;
  const int n2 = a.axis == 2 ? geo.ly : geo.nzl;
  double sx = 0, sy = 0, sz = 0, sd = 0, cnt = 0;
  for (long long q = threadIdx.x; q < (long long)n1 * n2; q += BLOCK) {
    const int b = (int)(q / n1), c = (int)(q - (long long)b * n1);
    const int i = a.axis == 0 ? p : c;
    const int j = a.axis == 0 ? c : (a.axis == 1 ? p : b);
    const int kk = a.axis == 2 ? p : b;
    const int g = i + geo.lx * j + geo.plane * (kk + 1);
    int fid;
    if (!lookup(geo, g, fid)) continue;  // solid: density and momentum are 0
    const double d = a.mom[fid];
    sd += d;
    sx += a.mom[geo.nfa + fid];
    sy += a.mom[2 * geo.nfa + fid];
    sz += a.mom[3 * geo.nfa + fid];
    cnt += (d > a.eps) ? 1.0 : 0.0;
  }
  __shared__ double red[5][BLOCK];
  red[0][threadIdx.x] = sx; red[1][threadIdx.x] = sy; red[2][threadIdx.x] = sz;
  red[3][threadIdx.x] = sd; red[4][threadIdx.x] = cnt;
  __syncthreads();
  for (int s = BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int v = 0; v < 5; ++v) red[v][threadIdx.x] += red[v][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 5) a.out[(long long)p * 5 + threadIdx.x] = red[threadIdx.x][0];
}

inline int small_grid(long long n) {
  long long b = (n + BLOCK - 1) / BLOCK;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

template <bool TAU1, int FMODE, int MINB>
void launch_step_cw(const LBArgs& a, bool check, bool writej, int grid, cudaStream_t st) {
  if (a.pipe) {
    if (check) lb_step_pipe_kernel<TAU1, FMODE, true, true, MINB><<<grid, BLOCK, 0, st>>>(a);
    else if (writej) lb_step_pipe_kernel<TAU1, FMODE, false, true, MINB><<<grid, BLOCK, 0, st>>>(a);
    else lb_step_pipe_kernel<TAU1, FMODE, false, false, MINB><<<grid, BLOCK, 0, st>>>(a);
    return;
  }
  if (check) lb_step_kernel<TAU1, FMODE, true, true, MINB><<<grid, BLOCK, 0, st>>>(a);
  else if (writej) lb_step_kernel<TAU1, FMODE, false, true, MINB><<<grid, BLOCK, 0, st>>>(a);
  else lb_step_kernel<TAU1, FMODE, false, false, MINB><<<grid, BLOCK, 0, st>>>(a);
}
template <bool TAU1, int MINB>
void launch_step_f(const LBArgs& a, int fmode, bool check, bool writej, int grid, cudaStream_t st) {
  if (fmode == FORCE_NONE) launch_step_cw<TAU1, FORCE_NONE, MINB>(a, check, writej, grid, st);
  else if (fmode == FORCE_UNIFORM) launch_step_cw<TAU1, FORCE_UNIFORM, MINB>(a, check, writej, grid, st);
  else launch_step_cw<TAU1, FORCE_FIELD, MINB>(a, check, writej, grid, st);
}
template <bool TAU1>
void launch_collide_f(const CollideArgs& a, int fmode, int grid, cudaStream_t st) {
  if (fmode == FORCE_NONE) collide_kernel<TAU1, FORCE_NONE><<<grid, BLOCK, 0, st>>>(a);
  else if (fmode == FORCE_UNIFORM) collide_kernel<TAU1, FORCE_UNIFORM><<<grid, BLOCK, 0, st>>>(a);
  else collide_kernel<TAU1, FORCE_FIELD><<<grid, BLOCK, 0, st>>>(a);
}

}  // namespace

int launch_lb_init(const Geo& g, long long fid_begin, long long fid_end, double rho0, const double a0[3], double* f,
                   double* mom, cudaStream_t st) {
  if (fid_end <= fid_begin) return 0;
  lb_init_kernel<<<small_grid(fid_end - fid_begin), BLOCK, 0, st>>>(g.nfa, fid_begin, fid_end, rho0, a0[0], a0[1], a0[2],
                                                                    f, mom);
  return 1;
}

int launch_fill_force(const Geo& g, long long nf, const double f[3], double* field, cudaStream_t st) {
  if (nf <= 0) return 0;
  fill_force_kernel<<<small_grid(nf), BLOCK, 0, st>>>(g.nfa, nf, f[0], f[1], f[2], field);
  return 1;
}

int launch_profile(const ProfileArgs& a, int rows, cudaStream_t st) {
  profile_kernel<<<rows, BLOCK, 0, st>>>(a);
  return 1;
}

int launch_lb_step(const LBArgs& a, bool tau1, int fmode, bool check, bool writej, int minb, int grid,
                   cudaStream_t st) {
  if (a.fid_end <= a.fid_begin) return 0;
  const int gr = clamp_grid(a.fid_end - a.fid_begin, grid);
  if (minb >= 3) {
    if (tau1) launch_step_f<true, 3>(a, fmode, check, writej, gr, st);
    else launch_step_f<false, 3>(a, fmode, check, writej, gr, st);
  } else {
    if (tau1) launch_step_f<true, 2>(a, fmode, check, writej, gr, st);
    else launch_step_f<false, 2>(a, fmode, check, writej, gr, st);
  }
  return 1;
}

int launch_collide(const CollideArgs& a, bool tau1, int fmode, int grid, cudaStream_t st) {
  if (a.fid_end <= a.fid_begin) return 0;
  const int gr = clamp_grid(a.fid_end - a.fid_begin, grid);
  if (tau1) launch_collide_f<true>(a, fmode, gr, st);
  else launch_collide_f<false>(a, fmode, gr, st);
  return 1;
}

int launch_moments(const MomArgs& a, int fmode, int grid, cudaStream_t st) {
  if (a.fid_end <= a.fid_begin) return 0;
  const int gr = clamp_grid(a.fid_end - a.fid_begin, grid);
  if (fmode == FORCE_NONE) moments_kernel<FORCE_NONE><<<gr, BLOCK, 0, st>>>(a);
  else if (fmode == FORCE_UNIFORM) moments_kernel<FORCE_UNIFORM><<<gr, BLOCK, 0, st>>>(a);
  else moments_kernel<FORCE_FIELD><<<gr, BLOCK, 0, st>>>(a);
  return 1;
}

int occupancy_grid_lb(int sm_count, int minb) {
  int per_sm = 0;
  if (minb >= 3)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lb_step_pipe_kernel<true, FORCE_UNIFORM, true, true, 3>, BLOCK, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lb_step_pipe_kernel<true, FORCE_UNIFORM, true, true, 2>, BLOCK, 0);
  if (per_sm < 1) per_sm = 1;
  return sm_count * per_sm;
}

}  // namespace lbg
