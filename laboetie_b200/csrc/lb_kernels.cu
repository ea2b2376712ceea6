// lb_kernels.cu -- Phase-A kernels for sm_100a: link masks, initial state,
// the fused pull stream + moments + collide step, moments/population read-back,
// plane profiles.
//
// Data layout in HBM: structure of arrays, fp64, x fastest.  Population l of
// node g lives at f[l*nalloc + g]; g = x + lx*(y + ly*p), p = plane index in the
// slab including one halo plane on each z side.  One thread owns one node; a
// warp covers 32 consecutive x, so every population access of a warp is one
// contiguous 256-byte run (the +-x neighbours are the same run shifted by 8 B).
//
// The step kernel K(t) implements, for fluid node r (SURVEY 8a "A2 o A3"):
//   n(t)(r,l)   = n*(t)(r-c_l, l)   if r-c_l is fluid          [streaming, equilibration.f90:227-243]
//               = n*(t)(r, inv l)    otherwise                  [bounce-back, equilibration.f90:204-222]
//   rho(t), j(t), ANY(n<0), max|j(t)-j(t-1)|                    [equilibration.f90:248-300,339-343]
//   n*(t+1)     = collide(n(t), rho(t), j(t), f)                [module_collision.f90:77-108]
// i.e. the reference's stream of step t fused with its collide of step t+1.
// n* is written to the other buffer (two-lattice), so a step can be redone when
// the driver changes the force after a convergence event.
#include <type_traits>

#include "lbg_internal.h"

#ifndef LBG_PREFETCH_MASK
#define LBG_PREFETCH_MASK 1
#endif
#ifndef LBG_BLOCKED
#define LBG_BLOCKED 0
#endif
// how the population loads go through the cache hierarchy: 0 = default (L1 allocating), 1 = .cg (L2 only),
// 2 = .cs (streaming)
#ifndef LBG_LOADMODE
#define LBG_LOADMODE 1
#endif

namespace lbg {
using namespace d3q19;

namespace {

template <int L, int END, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (L < END) {
    f(std::integral_constant<int, L>{});
    static_for<L + 1, END>(f);
  }
}

// neighbour offsets of one node in the linear alloc index
struct Nb {
  int oxm, oxp, oym, oyp, ozm, ozp;
};

__device__ __forceinline__ Nb neighbours(const Geo& geo, int g) {
  const int p = g / geo.plane;
  const int rem = g - p * geo.plane;
  const int y = rem / geo.lx;
  const int x = rem - y * geo.lx;
  Nb nb;
  nb.oxm = (x == 0) ? (geo.lx - 1) : -1;
  nb.oxp = (x == geo.lx - 1) ? -(geo.lx - 1) : 1;
  nb.oym = (y == 0) ? (geo.ly - 1) * geo.lx : -geo.lx;
  nb.oyp = (y == geo.ly - 1) ? -(geo.ly - 1) * geo.lx : geo.lx;
  nb.ozm = (geo.zwrap && p == 1) ? (geo.nzl - 1) * geo.plane : -geo.plane;
  nb.ozp = (geo.zwrap && p == geo.nzl) ? -(geo.nzl - 1) * geo.plane : geo.plane;
  return nb;
}

// offset of node r + s*c_L (s = +1 or -1)
template <int L, int S>
__device__ __forceinline__ int offset(const Nb& nb) {
  constexpr int X = S * cx(L), Y = S * cy(L), Z = S * cz(L);
  int o = 0;
  if constexpr (X > 0) o += nb.oxp;
  if constexpr (X < 0) o += nb.oxm;
  if constexpr (Y > 0) o += nb.oyp;
  if constexpr (Y < 0) o += nb.oym;
  if constexpr (Z > 0) o += nb.ozp;
  if constexpr (Z < 0) o += nb.ozm;
  return o;
}

__device__ __forceinline__ double ld_pop(const double* p) {
#if LBG_LOADMODE == 1
  return __ldcg(p);
#elif LBG_LOADMODE == 2
  return __ldcs(p);
#else
  return *p;
#endif
}

// n(t)(r,·) by pull with halfway bounce-back.
__device__ __forceinline__ void pull(const double* __restrict__ fin, long long nalloc, int g, uint32_t m, const Nb& nb,
                                     double (&n)[NV]) {
  static_for<0, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    if constexpr (L == 0) {
      n[0] = ld_pop(fin + g);
    } else {
      const bool src_fluid = (m >> inv(L)) & 1u;  // r - c_L == r + c_inv(L)
      const int idx = src_fluid ? g + offset<L, -1>(nb) : g;
      const int arr = src_fluid ? L : inv(L);
      n[L] = ld_pop(fin + (long long)arr * nalloc + idx);
    }
  });
}

// equilibration.f90:254 and :293-300, sequential in l.
__device__ __forceinline__ void moments(const double (&n)[NV], double fjx_half, double fjy_half, double fjz_half,
                                        double& rho, double& jx, double& jy, double& jz, bool& negative) {
  rho = n[0];
  jx = fjx_half;
  jy = fjy_half;
  jz = fjz_half;
  negative = n[0] < 0;
  static_for<1, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    rho = rho + n[L];
    negative = negative || (n[L] < 0);
    if constexpr (cx(L) > 0) jx = jx + n[L];
    if constexpr (cx(L) < 0) jx = jx - n[L];
    if constexpr (cy(L) > 0) jy = jy + n[L];
    if constexpr (cy(L) < 0) jy = jy - n[L];
    if constexpr (cz(L) > 0) jz = jz + n[L];
    if constexpr (cz(L) < 0) jz = jz - n[L];
  });
}

// module_collision.f90:77-108 on one fluid node, in the reference's association order.
template <bool TAU1, bool FORCED>
__device__ __forceinline__ void collide(double (&n)[NV], const Consts& k, double rho, double jx, double jy, double jz,
                                        double fx, double fy, double fz, double w1, double w2, double w3) {
  const double ux = jx / rho, uy = jy / rho, uz = jz / rho;
  const double pxx = jx * ux, pxy = jx * uy, pxz = jx * uz;
  const double pyx = jy * ux, pyy = jy * uy, pyz = jy * uz;
  const double pzx = jz * ux, pzy = jz * uy, pzz = jz * uz;
  const double qx1 = pxx * k.c1, qx0 = pxx * k.mcsq;
  const double qy1 = pyy * k.c1, qy0 = pyy * k.mcsq;
  const double qz1 = pzz * k.c1, qz0 = pzz * k.mcsq;
  const double a0rho[3] = {k.a0[0] * rho, k.a0[1] * rho, k.a0[2] * rho};
  // (c - u) * f for c in {-1, 0, +1}
  double gx[3], gy[3], gz[3];
  if constexpr (FORCED) {
    gx[0] = (-1.0 - ux) * fx; gx[1] = (0.0 - ux) * fx; gx[2] = (1.0 - ux) * fx;
    gy[0] = (-1.0 - uy) * fy; gy[1] = (0.0 - uy) * fy; gy[2] = (1.0 - uy) * fy;
    gz[0] = (-1.0 - uz) * fz; gz[1] = (0.0 - uz) * fz; gz[2] = (1.0 - uz) * fz;
  }
  static_for<0, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    constexpr int X = cx(L), Y = cy(L), Z = cz(L), K = kind(L);
    const double cj = cdot<L>(jx, jy, jz);
    double br = X ? qx1 : qx0;
    if constexpr (X && Y) br = br + (X * Y > 0 ? pxy : -pxy);
    if constexpr (X && Z) br = br + (X * Z > 0 ? pxz : -pxz);
    if constexpr (Y && X) br = br + (Y * X > 0 ? pyx : -pyx);
    br = br + (Y ? qy1 : qy0);
    if constexpr (Y && Z) br = br + (Y * Z > 0 ? pyz : -pyz);
    if constexpr (Z && X) br = br + (Z * X > 0 ? pzx : -pzx);
    if constexpr (Z && Y) br = br + (Z * Y > 0 ? pzy : -pzy);
    br = br + (Z ? qz1 : qz0);
    const double neq = (a0rho[K] + k.a1[K] * cj) + k.a2[K] * br;
    double v;
    if constexpr (TAU1) v = neq;  // w1 == 0, w2 == 1: 0*n + 1*neq == neq
    else v = w1 * n[L] + w2 * neq;
    if constexpr (FORCED) {
      const double g1 = (gx[X + 1] + gy[Y + 1]) + gz[Z + 1];
      const double cu = cdot<L>(ux, uy, uz);
      const double cf = cdot<L>(fx, fy, fz);
      const double force = k.a1[K] * g1 + (k.two_a2[K] * cu) * cf;
      v = v + w3 * force;
    }
    n[L] = v;
  });
}

// True when the 32-byte sector (4 consecutive nodes, all arrays are 256-byte aligned and nalloc is a
// multiple of 32) that holds node g contains a node with `bit` set.  Threads of such a sector all
// store (zeros on solid nodes, which is what those nodes hold anyway): a fully written sector needs no
// read-fill from HBM, a partially written one costs a DRAM read on top of the write.
#ifndef LBG_FILL
#define LBG_FILL 4  // nodes per fill group: 4 = one 32-byte sector
#endif
__device__ __forceinline__ bool sector_has(const uint32_t* __restrict__ mask, int g, uint32_t bit) {
  const uint4* p = reinterpret_cast<const uint4*>(mask + (g & ~(LBG_FILL - 1)));
  uint32_t any = 0;
#pragma unroll
  for (int i = 0; i < LBG_FILL / 4; ++i) {
    const uint4 mm = __ldg(p + i);
    any |= mm.x | mm.y | mm.z | mm.w;
  }
  return (any & bit) != 0;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One fluid node of K(t): pull, moments, convergence / negativity bookkeeping, collide, store.
template <bool TAU1, int FMODE, bool CHECK, bool WRITEJ>
__device__ __forceinline__ void lb_fluid_node(const LBArgs& a, int g, uint32_t m, double& dmax, bool& any_neg) {
  const long long nalloc = a.geo.nalloc;
  const Nb nb = neighbours(a.geo, g);
  double n[NV];
  pull(a.fin, nalloc, g, m, nb, n);
  double fjx = 0, fjy = 0, fjz = 0, fcx = 0, fcy = 0, fcz = 0;
  if constexpr (FMODE == FORCE_UNIFORM) {
    fjx = a.fj[0]; fjy = a.fj[1]; fjz = a.fj[2];
    fcx = a.fc[0]; fcy = a.fc[1]; fcz = a.fc[2];
  } else if constexpr (FMODE == FORCE_FIELD) {
    fjx = a.fj_field[g]; fjy = a.fj_field[nalloc + g]; fjz = a.fj_field[2 * nalloc + g];
    fcx = a.fc_field[g]; fcy = a.fc_field[nalloc + g]; fcz = a.fc_field[2 * nalloc + g];
  }
  double rho, jx, jy, jz;
  bool neg;
  moments(n, fjx / 2.0, fjy / 2.0, fjz / 2.0, rho, jx, jy, jz, neg);
  any_neg |= neg;
  if constexpr (CHECK) {
    const double ox = a.jold[g], oy = a.jold[nalloc + g], oz = a.jold[2 * nalloc + g];
    dmax = fmax(dmax, fmax(fabs(jx - ox), fmax(fabs(jy - oy), fabs(jz - oz))));
  }
  if constexpr (WRITEJ) {
    a.jnew[g] = jx;
    a.jnew[nalloc + g] = jy;
    a.jnew[2 * nalloc + g] = jz;
  }
  collide<TAU1, FMODE != FORCE_NONE>(n, a.k, rho, jx, jy, jz, fcx, fcy, fcz, a.w1, a.w2, a.w3);
  static_for<0, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    a.fout[(long long)L * nalloc + g] = n[L];
  });
}

// A solid node that shares a 32-byte sector with a fluid node stores its zeros (populations and j are 0
// on solid nodes, init_simu.f90:32-39), so the sector is written whole and needs no read-fill from HBM.
template <bool WRITEJ>
__device__ __forceinline__ void lb_fill_node(const LBArgs& a, int g) {
  const long long nalloc = a.geo.nalloc;
  if constexpr (WRITEJ) {
    a.jnew[g] = 0.0;
    a.jnew[nalloc + g] = 0.0;
    a.jnew[2 * nalloc + g] = 0.0;
  }
  static_for<0, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    a.fout[(long long)L * nalloc + g] = 0.0;
  });
}

// ---------------------------------------------------------------------------
// MINB = resident blocks per SM the register allocation aims at: 3 (80 registers) keeps more loads in
// flight on mostly-fluid lattices, 2 (124 registers, no spills) is faster on porous ones, where solid
// lanes idle and instruction issue, not memory latency, is the co-limiter.
template <bool TAU1, int FMODE, bool CHECK, bool WRITEJ, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) lb_step_kernel(const __grid_constant__ LBArgs a) {
  __shared__ int s_stop;
  __shared__ double s_red[BLOCK / 32];
  __shared__ int s_neg;
  if (threadIdx.x == 0) {
    int stop = *(volatile int*)&a.ctrl->stop | *(volatile int*)&a.ctrl->neg_step_idx;
    if (!stop && a.prev_checked && a.prev_may_stop) {
      const double prev = __longlong_as_double((long long)*(volatile unsigned long long*)&a.l2_slots[a.batch_idx - 1]);
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
  const long long nalloc = geo.nalloc;
  double dmax = 0.0;
  bool any_neg = false;
  // tiles of BLOCK consecutive nodes; the mask word of the next tile is fetched one iteration ahead so
  // that its DRAM latency does not sit in front of the 22 dependent population loads
  const int ntiles = (int)((a.g_end - a.g_begin + BLOCK - 1) / BLOCK);
#if LBG_BLOCKED
  const int tpb = (ntiles + gridDim.x - 1) / gridDim.x;
  const int t_first = blockIdx.x * tpb, t_last = min(ntiles, t_first + tpb), t_step = 1;
#else
  const int t_first = blockIdx.x, t_last = ntiles, t_step = gridDim.x;
#endif
  auto node_of = [&](int tile) { return a.g_begin + (long long)tile * BLOCK + threadIdx.x; };
#if LBG_PREFETCH_MASK
  uint32_t m_next = 0;
  if (t_first < t_last && node_of(t_first) < a.g_end) m_next = __ldg(a.mask + node_of(t_first));
#endif
  for (int tile = t_first; tile < t_last; tile += t_step) {
    const long long gg = node_of(tile);
#if LBG_PREFETCH_MASK
    const uint32_t m = m_next;
    {
      const int tn = tile + t_step;
      m_next = (tn < t_last && node_of(tn) < a.g_end) ? __ldg(a.mask + node_of(tn)) : 0u;
    }
    if (gg >= a.g_end) continue;
    const int g = (int)gg;
#else
    if (gg >= a.g_end) continue;
    const int g = (int)gg;
    const uint32_t m = __ldg(a.mask + g);
#endif
    if (m & MASK_FLUID) lb_fluid_node<TAU1, FMODE, CHECK, WRITEJ>(a, g, m, dmax, any_neg);
    else if (sector_has(a.mask, g, MASK_FLUID)) lb_fill_node<WRITEJ>(a, g);
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
      atomicMax(&a.l2_slots[a.batch_idx], (unsigned long long)__double_as_longlong(v));
    }
    if (s_neg) atomicCAS(&a.ctrl->neg_step_idx, 0, a.batch_idx + 1);
  }
}

// Porous lattices: the same step with block-level compaction.  A block takes a super-tile of
// CT_NODES consecutive nodes, turns its mask words into a list of fluid nodes (ballot-free prefix
// sums over 8-node groups) and a list of fill nodes in shared memory, and then every lane works on a
// fluid node.  Without this a warp on a 60 %-fluid lattice runs with ~19 of 32 lanes active and the
// kernel is bound by instruction issue and exposed latency rather than by HBM.  Per-node arithmetic
// is the shared lb_fluid_node(), so results are identical to lb_step_kernel.
constexpr int CT_PER_THREAD = 8;
constexpr int CT_NODES = BLOCK * CT_PER_THREAD;

template <bool TAU1, int FMODE, bool CHECK, bool WRITEJ>
__global__ void __launch_bounds__(BLOCK, 2) lb_step_compact_kernel(const __grid_constant__ LBArgs a) {
  __shared__ int s_stop;
  __shared__ double s_red[BLOCK / 32];
  __shared__ int s_neg;
  __shared__ uint32_t s_mask[CT_NODES];
  __shared__ unsigned short s_fluid[CT_NODES];
  __shared__ unsigned short s_fill[CT_NODES];
  __shared__ int s_wsum[2][BLOCK / 32];
  __shared__ int s_tot[2];
  if (threadIdx.x == 0) {
    int stop = *(volatile int*)&a.ctrl->stop | *(volatile int*)&a.ctrl->neg_step_idx;
    if (!stop && a.prev_checked && a.prev_may_stop) {
      const double prev = __longlong_as_double((long long)*(volatile unsigned long long*)&a.l2_slots[a.batch_idx - 1]);
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

  double dmax = 0.0;
  bool any_neg = false;
  const long long g_lo = a.g_begin & ~7LL;  // super-tiles start on an 8-node boundary (uint4 mask loads)
  const int ntiles = (int)((a.g_end - g_lo + CT_NODES - 1) / CT_NODES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  auto load_masks = [&](int tile, uint32_t (&mk)[CT_PER_THREAD]) {
    const long long g0 = g_lo + (long long)tile * CT_NODES + threadIdx.x * CT_PER_THREAD;
#pragma unroll
    for (int i = 0; i < CT_PER_THREAD; ++i) mk[i] = 0;
    if (tile < ntiles && g0 < a.g_end) {  // nalloc is padded, so the vector loads stay inside the array
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(a.mask + g0));
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.mask + g0 + 4));
      mk[0] = u.x; mk[1] = u.y; mk[2] = u.z; mk[3] = u.w;
      mk[4] = v.x; mk[5] = v.y; mk[6] = v.z; mk[7] = v.w;
#pragma unroll
      for (int i = 0; i < CT_PER_THREAD; ++i)
        if (g0 + i < a.g_begin || g0 + i >= a.g_end) mk[i] = 0;  // outside this launch's range
    }
  };

  uint32_t mk[CT_PER_THREAD], mk_next[CT_PER_THREAD];
  load_masks(blockIdx.x, mk_next);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll
    for (int i = 0; i < CT_PER_THREAD; ++i) mk[i] = mk_next[i];
    load_masks(tile + gridDim.x, mk_next);  // one super-tile ahead
    // --- phase 1: lists of fluid nodes and of fill nodes (solid, in a sector with a fluid node)
    int nf = 0, nz = 0;
    const bool g0_has = ((mk[0] | mk[1] | mk[2] | mk[3]) & MASK_FLUID) != 0;
    const bool g1_has = ((mk[4] | mk[5] | mk[6] | mk[7]) & MASK_FLUID) != 0;
#pragma unroll
    for (int i = 0; i < CT_PER_THREAD; ++i) {
      const bool fl = mk[i] & MASK_FLUID;
      nf += fl ? 1 : 0;
      nz += (!fl && (i < 4 ? g0_has : g1_has)) ? 1 : 0;
    }
    // the part of the range outside [g_begin, g_end) must not be filled either
    {
      const long long g0 = g_lo + (long long)tile * CT_NODES + threadIdx.x * CT_PER_THREAD;
      int nz2 = 0;
#pragma unroll
      for (int i = 0; i < CT_PER_THREAD; ++i) {
        const bool fl = mk[i] & MASK_FLUID;
        const bool inside = (g0 + i >= a.g_begin) && (g0 + i < a.g_end);
        nz2 += (!fl && inside && (i < 4 ? g0_has : g1_has)) ? 1 : 0;
      }
      nz = nz2;
    }
    int pf = nf, pz = nz;  // inclusive warp scans
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tf = __shfl_up_sync(0xffffffffu, pf, o), tz = __shfl_up_sync(0xffffffffu, pz, o);
      if (lane >= o) {
        pf += tf;
        pz += tz;
      }
    }
    if (lane == 31) {
      s_wsum[0][warp] = pf;
      s_wsum[1][warp] = pz;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      int acc = 0;
      for (int w = 0; w < BLOCK / 32; ++w) {
        const int v = s_wsum[threadIdx.x][w];
        s_wsum[threadIdx.x][w] = acc;
        acc += v;
      }
      s_tot[threadIdx.x] = acc;
    }
    __syncthreads();
    {
      int of = s_wsum[0][warp] + pf - nf, oz = s_wsum[1][warp] + pz - nz;
      const long long g0 = g_lo + (long long)tile * CT_NODES + threadIdx.x * CT_PER_THREAD;
#pragma unroll
      for (int i = 0; i < CT_PER_THREAD; ++i) {
        const int off = threadIdx.x * CT_PER_THREAD + i;
        s_mask[off] = mk[i];
        const bool fl = mk[i] & MASK_FLUID;
        const bool inside = (g0 + i >= a.g_begin) && (g0 + i < a.g_end);
        if (fl) s_fluid[of++] = (unsigned short)off;
        else if (inside && (i < 4 ? g0_has : g1_has)) s_fill[oz++] = (unsigned short)off;
      }
    }
    __syncthreads();
    // --- phase 2: every lane on a fluid node
    const int tot_f = s_tot[0], tot_z = s_tot[1];
    const long long base = g_lo + (long long)tile * CT_NODES;
    for (int i = threadIdx.x; i < tot_f; i += BLOCK) {
      const int off = s_fluid[i];
      lb_fluid_node<TAU1, FMODE, CHECK, WRITEJ>(a, (int)(base + off), s_mask[off], dmax, any_neg);
    }
    for (int i = threadIdx.x; i < tot_z; i += BLOCK) lb_fill_node<WRITEJ>(a, (int)(base + s_fill[i]));
    __syncthreads();  // lists are rebuilt by the next super-tile
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
      atomicMax(&a.l2_slots[a.batch_idx], (unsigned long long)__double_as_longlong(v));
    }
    if (s_neg) atomicCAS(&a.ctrl->neg_step_idx, 0, a.batch_idx + 1);
  }
}

template <bool TAU1, int FMODE>
__global__ void __launch_bounds__(BLOCK) collide_kernel(const __grid_constant__ CollideArgs a) {
  const long long nalloc = a.geo.nalloc;
  for (long long gg = a.g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; gg < a.g_end;
       gg += (long long)gridDim.x * BLOCK) {
    const int g = (int)gg;
    const uint32_t m = __ldg(a.mask + g);
    double n[NV];
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      n[L] = a.fin[(long long)L * nalloc + g];
    });
    if (m & MASK_FLUID) {
      double fcx = 0, fcy = 0, fcz = 0;
      if constexpr (FMODE == FORCE_UNIFORM) {
        fcx = a.fc[0]; fcy = a.fc[1]; fcz = a.fc[2];
      } else if constexpr (FMODE == FORCE_FIELD) {
        fcx = a.fc_field[g]; fcy = a.fc_field[nalloc + g]; fcz = a.fc_field[2 * nalloc + g];
      }
      collide<TAU1, FMODE != FORCE_NONE>(n, a.k, a.mom[g], a.mom[nalloc + g], a.mom[2 * nalloc + g],
                                         a.mom[3 * nalloc + g], fcx, fcy, fcz, a.w1, a.w2, a.w3);
    }
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      a.fout[(long long)L * nalloc + g] = n[L];
    });
  }
}

template <int FMODE>
__global__ void __launch_bounds__(BLOCK) moments_kernel(const __grid_constant__ MomArgs a) {
  const long long nalloc = a.geo.nalloc;
  for (long long gg = a.g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; gg < a.g_end;
       gg += (long long)gridDim.x * BLOCK) {
    const int g = (int)gg;
    const uint32_t m = __ldg(a.mask + g);
    double n[NV];
    double rho = 0, jx = 0, jy = 0, jz = 0;
    if (m & MASK_FLUID) {
      const Nb nb = neighbours(a.geo, g);
      pull(a.fin, nalloc, g, m, nb, n);
      double fjx = 0, fjy = 0, fjz = 0;
      if constexpr (FMODE == FORCE_UNIFORM) {
        fjx = a.fj[0]; fjy = a.fj[1]; fjz = a.fj[2];
      } else if constexpr (FMODE == FORCE_FIELD) {
        fjx = a.fj_field[g]; fjy = a.fj_field[nalloc + g]; fjz = a.fj_field[2 * nalloc + g];
      }
      bool neg;
      moments(n, fjx / 2.0, fjy / 2.0, fjz / 2.0, rho, jx, jy, jz, neg);
    } else {
      static_for<0, NV>([&](auto Lc) { n[decltype(Lc)::value] = 0.0; });
    }
    if (a.mom) {
      a.mom[g] = rho;
      a.mom[nalloc + g] = jx;
      a.mom[2 * nalloc + g] = jy;
      a.mom[3 * nalloc + g] = jz;
    }
    if (a.pops) {
      static_for<0, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        a.pops[(long long)L * nalloc + g] = n[L];
      });
    }
  }
}

// supercell_definition.f90:115-147 + the neighbour tables of equilibration.f90:109-119,
// folded into one word per node.  nature has valid halo planes.
__global__ void __launch_bounds__(BLOCK) build_mask_kernel(Geo geo, const int8_t* __restrict__ nat,
                                                           uint32_t* __restrict__ mask) {
  const long long g_begin = geo.plane, g_end = (long long)geo.plane * (geo.nzl + 1);
  for (long long gg = g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; gg < g_end;
       gg += (long long)gridDim.x * BLOCK) {
    const int g = (int)gg;
    Geo gz = geo;
    gz.zwrap = 0;  // nature carries real halo planes
    const Nb nb = neighbours(gz, g);
    const int8_t me = nat[g];
    uint32_t m = (me == 0) ? MASK_FLUID : 0u;
    bool interfacial = false;
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      const int8_t other = nat[g + offset<L, +1>(nb)];
      if (other == 0) m |= (1u << L);
      interfacial = interfacial || (other != me);
    });
    if (interfacial) m |= MASK_INTERFACIAL;
    mask[g] = m;
  }
}

// init_simu.f90:24-39 and equilibration.f90:59,75-80
__global__ void __launch_bounds__(BLOCK) lb_init_kernel(Geo geo, const uint32_t* __restrict__ mask, double rho0,
                                                        double a00, double a01, double a02, double* __restrict__ f,
                                                        double* __restrict__ mom) {
  const long long g_begin = geo.plane, g_end = (long long)geo.plane * (geo.nzl + 1);
  const double a0[3] = {a00, a01, a02};
  for (long long g = g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; g < g_end;
       g += (long long)gridDim.x * BLOCK) {
    const double dens = (mask[g] & MASK_FLUID) ? rho0 : 0.0;
    static_for<0, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      f[(long long)L * geo.nalloc + g] = dens * a0[kind(L)];
    });
    mom[g] = dens;
    mom[geo.nalloc + g] = 0.0;
    mom[2 * geo.nalloc + g] = 0.0;
    mom[3 * geo.nalloc + g] = 0.0;
  }
}

// equilibration.f90:381-386 materialised as a field (used when a uniform force
// and a force field meet across a force change).
__global__ void __launch_bounds__(BLOCK) fill_force_kernel(Geo geo, const uint32_t* __restrict__ mask, double fx,
                                                           double fy, double fz, double* __restrict__ field) {
  const long long g_begin = geo.plane, g_end = (long long)geo.plane * (geo.nzl + 1);
  for (long long g = g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; g < g_end;
       g += (long long)gridDim.x * BLOCK) {
    const bool fl = mask[g] & MASK_FLUID;
    field[g] = fl ? fx : 0.0;
    field[geo.nalloc + g] = fl ? fy : 0.0;
    field[2 * geo.nalloc + g] = fl ? fz : 0.0;
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
    const long long g = (long long)i + (long long)geo.lx * j + (long long)geo.plane * (kk + 1);
    const double d = a.mom[g];
    sd += d;
    sx += a.mom[geo.nalloc + g];
    sy += a.mom[2 * geo.nalloc + g];
    sz += a.mom[3 * geo.nalloc + g];
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

__global__ void __launch_bounds__(BLOCK) count_flags_kernel(Geo geo, const uint32_t* __restrict__ mask,
                                                            unsigned long long* counts) {
  const long long g_begin = geo.plane, g_end = (long long)geo.plane * (geo.nzl + 1);
  unsigned int nf = 0, nif = 0;
  for (long long g = g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; g < g_end;
       g += (long long)gridDim.x * BLOCK) {
    const uint32_t m = mask[g];
    if (m & MASK_FLUID) {
      ++nf;
      if (m & MASK_INTERFACIAL) ++nif;
    }
  }
  nf = __reduce_add_sync(0xffffffffu, nf);
  nif = __reduce_add_sync(0xffffffffu, nif);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counts[0], (unsigned long long)nf);
    atomicAdd(&counts[1], (unsigned long long)nif);
  }
}

__global__ void __launch_bounds__(BLOCK) extract_flag_kernel(Geo geo, const uint32_t* __restrict__ mask, uint32_t bit,
                                                             int8_t* __restrict__ out) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK)
    out[q] = (mask[q + geo.plane] & bit) ? 1 : 0;
}

inline int small_grid(long long n) {
  long long b = (n + BLOCK - 1) / BLOCK;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace

int launch_build_mask(const Geo& g, const int8_t* nature_halo, uint32_t* mask, cudaStream_t st) {
  build_mask_kernel<<<small_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, nature_halo, mask);
  return 1;
}

int launch_lb_init(const Geo& g, const uint32_t* mask, double rho0, const double a0[3], double* f, double* mom,
                   cudaStream_t st) {
  lb_init_kernel<<<small_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, mask, rho0, a0[0], a0[1], a0[2], f, mom);
  return 1;
}

int launch_fill_force(const Geo& g, const uint32_t* mask, const double f[3], double* field, cudaStream_t st) {
  fill_force_kernel<<<small_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, mask, f[0], f[1], f[2], field);
  return 1;
}

int launch_count_flags(const Geo& g, const uint32_t* mask, unsigned long long* counts2, cudaStream_t st) {
  count_flags_kernel<<<small_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, mask, counts2);
  return 1;
}

int launch_extract_flag(const Geo& g, const uint32_t* mask, uint32_t bit, int8_t* out_own, cudaStream_t st) {
  extract_flag_kernel<<<small_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, mask, bit, out_own);
  return 1;
}

int launch_profile(const ProfileArgs& a, int rows, cudaStream_t st) {
  profile_kernel<<<rows, BLOCK, 0, st>>>(a);
  return 1;
}

namespace {
template <bool TAU1, int FMODE, int MINB>
void launch_step_cw(const LBArgs& a, bool check, bool writej, int grid, cudaStream_t st) {
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
int clamp_grid(long long n, int grid) {
  const long long b = (n + BLOCK - 1) / BLOCK;
  return (int)(b < 1 ? 1 : (b < grid ? b : grid));
}
}  // namespace

namespace {
template <bool TAU1, int FMODE>
void launch_compact_cw(const LBArgs& a, bool check, bool writej, int grid, cudaStream_t st) {
  if (check) lb_step_compact_kernel<TAU1, FMODE, true, true><<<grid, BLOCK, 0, st>>>(a);
  else if (writej) lb_step_compact_kernel<TAU1, FMODE, false, true><<<grid, BLOCK, 0, st>>>(a);
  else lb_step_compact_kernel<TAU1, FMODE, false, false><<<grid, BLOCK, 0, st>>>(a);
}
template <bool TAU1>
void launch_compact_f(const LBArgs& a, int fmode, bool check, bool writej, int grid, cudaStream_t st) {
  if (fmode == FORCE_NONE) launch_compact_cw<TAU1, FORCE_NONE>(a, check, writej, grid, st);
  else if (fmode == FORCE_UNIFORM) launch_compact_cw<TAU1, FORCE_UNIFORM>(a, check, writej, grid, st);
  else launch_compact_cw<TAU1, FORCE_FIELD>(a, check, writej, grid, st);
}
}  // namespace

int launch_lb_step(const LBArgs& a, bool tau1, int fmode, bool check, bool writej, int minb, int grid,
                   cudaStream_t st) {
  if (minb == 0) {  // block-compacting kernel for porous lattices
    const long long nt = (a.g_end - (a.g_begin & ~7LL) + CT_NODES - 1) / CT_NODES;
    const int gr = (int)(nt < 1 ? 1 : (nt < grid ? nt : grid));
    if (tau1) launch_compact_f<true>(a, fmode, check, writej, gr, st);
    else launch_compact_f<false>(a, fmode, check, writej, gr, st);
    return 1;
  }
  const int gr = clamp_grid(a.g_end - a.g_begin, grid);
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
  const int gr = clamp_grid(a.g_end - a.g_begin, grid);
  if (tau1) launch_collide_f<true>(a, fmode, gr, st);
  else launch_collide_f<false>(a, fmode, gr, st);
  return 1;
}

int launch_moments(const MomArgs& a, int fmode, int grid, cudaStream_t st) {
  const int gr = clamp_grid(a.g_end - a.g_begin, grid);
  if (fmode == FORCE_NONE) moments_kernel<FORCE_NONE><<<gr, BLOCK, 0, st>>>(a);
  else if (fmode == FORCE_UNIFORM) moments_kernel<FORCE_UNIFORM><<<gr, BLOCK, 0, st>>>(a);
  else moments_kernel<FORCE_FIELD><<<gr, BLOCK, 0, st>>>(a);
  return 1;
}

int occupancy_grid_lb(int sm_count, int minb) {
  int per_sm = 0;
  if (minb == 0)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lb_step_compact_kernel<true, FORCE_UNIFORM, true, true>, BLOCK, 0);
  else if (minb >= 3)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lb_step_kernel<true, FORCE_UNIFORM, true, true, 3>, BLOCK, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lb_step_kernel<true, FORCE_UNIFORM, true, true, 2>, BLOCK, 0);
  if (per_sm < 1) per_sm = 1;
  return sm_count * per_sm;
}

}  // namespace lbg
