// mp_kernels.cu -- Phase-B kernels for sm_100a: tracer moment propagation.
//
// The flow is frozen during Phase B (drop_tracers.f90:41-49), so everything the
// reference recomputes per step from n(l,i,j,k) and density is static.  mp_init
// evaluates it once, with the reference's own expressions and association order,
// and stores per fluid node (SoA, x fastest, fp64):
//   q_l(r)  l=1..18   incoming link probability p_{inv l}(r+c_l)  (scattprop_p, :228)
//   s_0(r)            fractionOfParticleRemaining after the neighbour loop, and
//                     after "- ka" on adsorbing interfacial nodes (:215-225,240)
//   s_1..3(r)         u_star (:226)
// A propagate step (:207-253) is then a pure gather-multiply-accumulate with no
// division: acc = sum_l P(r+c_l) q_l(r) in the reference's l order, the
// adsorption branch, and the vacf reduction.  Results are bit-identical per node
// to recomputing the probabilities every step, because the same fp64 operations
// are applied to the same operands in the same order.
//
// Propagated_Quantity is kept as three SoA arrays per time level and the
// now/next array copies (:262-267) become a pointer swap: "next" is fully
// overwritten on fluid nodes and stays 0 on solid nodes.
#include <type_traits>

#include "lbg_internal.h"

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

template <int L>
__device__ __forceinline__ int offset_plus(const Nb& nb) {
  constexpr int X = cx(L), Y = cy(L), Z = cz(L);
  int o = 0;
  if constexpr (X > 0) o += nb.oxp;
  if constexpr (X < 0) o += nb.oxm;
  if constexpr (Y > 0) o += nb.oyp;
  if constexpr (Y < 0) o += nb.oym;
  if constexpr (Z > 0) o += nb.ozp;
  if constexpr (Z < 0) o += nb.ozm;
  return o;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sums of three values in a fixed order; result valid on thread 0
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double (*sh)[BLOCK / 32]) {
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    sh[0][w] = a;
    sh[1][w] = b;
    sh[2][w] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = sh[0][0];
    b = sh[1][0];
    c = sh[2][0];
    for (int i = 1; i < BLOCK / 32; ++i) {
      a += sh[0][i];
      b += sh[1][i];
      c += sh[2][i];
    }
  }
}

// update_tracer_population (drop_tracers.f90:97-105) for direction L at a fluid
// node with t = j + f_ext:  a0*rho + a1*sum(c*t);  then calc_scattprop
// (module_moment_propagation.f90:341-346) with fermi = 1/2 (neutral tracer).
template <int L>
__device__ __forceinline__ double scattprop(const Consts& k, const double (&lambda_w)[3], double rho, double tx,
                                            double ty, double tz) {
  constexpr int K = kind(L);
  const double n = k.a0[K] * rho + k.a1[K] * cdot<L>(tx, ty, tz);
  return (n / rho - k.a0[K]) + lambda_w[K] * 0.5;
}

// drop_tracers.f90:63-105 + module_moment_propagation.f90:96-137, plus the
// static part of propagate (:213-228,240,249).
__global__ void __launch_bounds__(BLOCK) mp_init_kernel(const __grid_constant__ MPInitArgs a) {
  __shared__ double sh[3][BLOCK / 32];
  const Geo& geo = a.geo;
  const long long nalloc = geo.nalloc;
  const double eps = 2.220446049250313e-16;  // epsilon(1._dp)
  double v0x = 0, v0y = 0, v0z = 0;
  bool bad = false;
  for (long long gg = a.g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; gg < a.g_end;
       gg += (long long)gridDim.x * BLOCK) {
    const int g = (int)gg;
    const uint32_t m = __ldg(a.mask + g);
    if (!(m & MASK_FLUID)) continue;
    const Nb nb = neighbours(geo, g);
    const double rho = a.mom[g];
    const double tx = a.mom[nalloc + g] + a.f[0];
    const double ty = a.mom[2 * nalloc + g] + a.f[1];
    const double tz = a.mom[3 * nalloc + g] + a.f[2];
    double frac = 1.0, usx = 0.0, usy = 0.0, usz = 0.0;
    double px = 0.0, py = 0.0, pz = 0.0;
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      constexpr int LI = inv(L);
      double q = 0.0;
      if ((m >> L) & 1u) {  // neighbour r + c_L is fluid
        const double sp = scattprop<L>(a.k, a.lambda_w, rho, tx, ty, tz);
        frac = frac - sp;
        if constexpr (cx(L) > 0) usx = usx + sp;
        if constexpr (cx(L) < 0) usx = usx - sp;
        if constexpr (cy(L) > 0) usy = usy + sp;
        if constexpr (cy(L) < 0) usy = usy - sp;
        if constexpr (cz(L) > 0) usz = usz + sp;
        if constexpr (cz(L) < 0) usz = usz - sp;
        // vacf(:,tini) += boltz_weight*scattprop*c**2   (:129)
        const double bs = a.bw * sp;
        if constexpr (cx(L) != 0) v0x += bs;
        if constexpr (cy(L) != 0) v0y += bs;
        if constexpr (cz(L) != 0) v0z += bs;
        const int gp = g + offset_plus<L>(nb);
        const double rhop = a.mom[gp];
        const double txp = a.mom[nalloc + gp] + a.f[0];
        const double typ = a.mom[2 * nalloc + gp] + a.f[1];
        const double tzp = a.mom[3 * nalloc + gp] + a.f[2];
        q = scattprop<LI>(a.k, a.lambda_w, rhop, txp, typ, tzp);
        // P(:,r,now) += exp_min_dphi*scattprop_p*c_inv(:)*boltz_weight   (:134-135)
        if constexpr (cx(LI) > 0) px = px + q * a.bw;
        if constexpr (cx(LI) < 0) px = px + (-q) * a.bw;
        if constexpr (cy(LI) > 0) py = py + q * a.bw;
        if constexpr (cy(LI) < 0) py = py + (-q) * a.bw;
        if constexpr (cz(LI) > 0) pz = pz + q * a.bw;
        if constexpr (cz(LI) < 0) pz = pz + (-q) * a.bw;
      }
      a.q[(long long)(L - 1) * nalloc + g] = q;
    });
    if (a.ads && (m & MASK_INTERFACIAL)) frac = frac - a.ka;  // :240
    if (frac < eps) bad = true;                              // :249
    a.s[g] = frac;
    a.s[nalloc + g] = usx;
    a.s[2 * nalloc + g] = usy;
    a.s[3 * nalloc + g] = usz;
    a.P0[g] = px;
    a.P0[nalloc + g] = py;
    a.P0[2 * nalloc + g] = pz;
  }
  if (bad) *a.err = 1;
  block_sum3(v0x, v0y, v0z, sh);
  if (threadIdx.x == 0) {
    a.partial[3 * blockIdx.x + 0] = v0x;
    a.partial[3 * blockIdx.x + 1] = v0y;
    a.partial[3 * blockIdx.x + 2] = v0z;
  }
}

// Tile order.  A tile is BLOCK consecutive nodes of one plane.  Tiles are walked strip by strip:
// within a strip of STRIP tiles of the plane's linear order, all planes are visited before moving to
// the next strip.  The three time-level-"now" planes a tile gathers from (z-1, z, z+1) and its y+-1
// rows were then touched a few thousand tiles ago at most, so they are still in L2 (126 MB) and
// Propagated_Quantity is read from HBM once per step instead of three times.
constexpr int STRIP = 256;
#ifndef LBG_MP_LOADMODE
#define LBG_MP_LOADMODE 2
#endif
__device__ __forceinline__ double ld_stream(const double* p) {
#if LBG_MP_LOADMODE == 1
  return __ldcg(p);
#elif LBG_MP_LOADMODE == 2
  return __ldcs(p);
#else
  return *p;
#endif
}

__device__ __forceinline__ bool tile_to_node(const Geo& geo, int tile, int p_begin, int np, int chunks_per_plane,
                                             int& g) {
  const int per_strip = STRIP * np;
  const int strip = tile / per_strip;
  const int rem = tile - strip * per_strip;
  const int c0 = strip * STRIP;
  const int cs = min(STRIP, chunks_per_plane - c0);
  const int p = rem / cs;
  const int c = rem - p * cs;
  const int in_plane = (c0 + c) * BLOCK + threadIdx.x;
  g = (p_begin + p) * geo.plane + in_plane;
  return in_plane < geo.plane;
}

// module_moment_propagation.f90:207-253 (one propagate call), see the header comment.
// Branch-free per node: the 18 link probabilities, the remaining fraction, u* and the node's own
// P are streamed in first (25 independent loads in flight), then the 18 neighbour gathers are
// issued without conditions -- a solid neighbour is replaced by the node itself and its q is 0
// (mp_init stores 0 there), so it adds exactly 0.
#ifndef LBG_MP_MINB
#define LBG_MP_MINB 2
#endif
__global__ void __launch_bounds__(BLOCK, LBG_MP_MINB) mp_step_kernel(const __grid_constant__ MPArgs a) {
  __shared__ double sh[3][BLOCK / 32];
  __shared__ int s_flag;
  if (threadIdx.x == 0) {
    int stop = *(volatile int*)&a.ctrl->stop;
    if (!stop && a.check_prev) {
      const volatile double* v = a.vacf_slots + 3 * (a.batch_idx - 1);
      const double ax = fabs(v[0]), ay = fabs(v[1]), az = fabs(v[2]);
      if (ax < a.lim && ay < a.lim && az < a.lim && ax < 1.e-12 && ay < 1.e-12 && az < 1.e-12) {  // :284
        a.ctrl->stop = 1;
        a.ctrl->stop_idx = a.batch_idx;
        stop = 1;
      }
    }
    s_flag = stop;
  }
  __syncthreads();
  if (s_flag) return;

  const Geo& geo = a.geo;
  const long long nalloc = geo.nalloc;
  const int np = a.p_end - a.p_begin;
  const int chunks = (geo.plane + BLOCK - 1) / BLOCK;
  const int ntiles = chunks * np;
  const uint32_t ADS = a.ads ? MASK_INTERFACIAL : 0u;  // adsorbing node <=> fluid && interfacial && ads
  double vx = 0, vy = 0, vz = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int g;
    if (!tile_to_node(geo, tile, a.p_begin, np, chunks, g)) continue;
    const uint4 mm = __ldg(reinterpret_cast<const uint4*>(a.mask + (g & ~3)));
    const int sub = g & 3;
    const uint32_t m = sub == 0 ? mm.x : (sub == 1 ? mm.y : (sub == 2 ? mm.z : mm.w));
    if (!((mm.x | mm.y | mm.z | mm.w) & MASK_FLUID)) continue;  // nothing to write in this 32-byte sector
    const bool fluid = m & MASK_FLUID;
    const bool adsorbing = fluid && (m & ADS);
    // does any fluid node of the sector adsorb?  (then the whole sector of the adsorbed field is written)
    const bool sector_ads = ADS && (((mm.x & MASK_FLUID) && (mm.x & ADS)) || ((mm.y & MASK_FLUID) && (mm.y & ADS)) ||
                                    ((mm.z & MASK_FLUID) && (mm.z & ADS)) || ((mm.w & MASK_FLUID) && (mm.w & ADS)));
    double nx = 0.0, ny = 0.0, nz = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
    if (fluid) {
      double q[NV - 1];
      static_for<1, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        q[L - 1] = ld_stream(a.q + (long long)(L - 1) * nalloc + g);
      });
      const double frac = ld_stream(a.s + g);
      const double usx = ld_stream(a.s + nalloc + g), usy = ld_stream(a.s + 2 * nalloc + g), usz = ld_stream(a.s + 3 * nalloc + g);
      const double px = a.Pnow[g], py = a.Pnow[nalloc + g], pz = a.Pnow[2 * nalloc + g];
      double sx = 0.0, sy = 0.0, sz = 0.0;
      if (adsorbing) {
        sx = a.Anow[g];
        sy = a.Anow[nalloc + g];
        sz = a.Anow[2 * nalloc + g];
      }
      const Nb nb = neighbours(geo, g);
      double ax = 0.0, ay = 0.0, az = 0.0;  // Propagated_Quantity(:,r,next) is always 0 on entry
      static_for<1, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        const int gp = ((m >> L) & 1u) ? g + offset_plus<L>(nb) : g;
        ax = ax + a.Pnow[gp] * q[L - 1];
        ay = ay + a.Pnow[nalloc + gp] * q[L - 1];
        az = az + a.Pnow[2 * nalloc + gp] * q[L - 1];
      });
      vx += px * usx;  // vacf(:,now) += P(:,r,now)*u_star   (:232)
      vy += py * usy;
      vz += pz * usz;
      if (!adsorbing) {  // :235-238
        nx = ax + frac * px;
        ny = ay + frac * py;
        nz = az + frac * pz;
      } else {  // :239-247
        nx = (ax + frac * px) + sx * a.kd;
        ny = (ay + frac * py) + sy * a.kd;
        nz = (az + frac * pz) + sz * a.kd;
        bx = sx * a.one_minus_kd + px * a.ka;
        by = sy * a.one_minus_kd + py * a.ka;
        bz = sz * a.one_minus_kd + pz * a.ka;
      }
    }
    // whole sectors are written: solid nodes hold 0, non-adsorbing nodes hold 0 in the adsorbed field
    __stcs(a.Pnext + g, nx);
    __stcs(a.Pnext + nalloc + g, ny);
    __stcs(a.Pnext + 2 * nalloc + g, nz);
    if (sector_ads) {
      __stcs(a.Anext + g, bx);
      __stcs(a.Anext + nalloc + g, by);
      __stcs(a.Anext + 2 * nalloc + g, bz);
    }
  }
  // vacf: block partials, then the last block to finish adds them in block order
  block_sum3(vx, vy, vz, sh);
  if (threadIdx.x == 0) {
    a.partial[3 * blockIdx.x + 0] = vx;
    a.partial[3 * blockIdx.x + 1] = vy;
    a.partial[3 * blockIdx.x + 2] = vz;
    __threadfence();
    const unsigned int done = atomicAdd(&a.ctrl->ticket, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      double tx = 0, ty = 0, tz = 0;
      for (unsigned int b = 0; b < gridDim.x; ++b) {
        tx += ((volatile double*)a.partial)[3 * b + 0];
        ty += ((volatile double*)a.partial)[3 * b + 1];
        tz += ((volatile double*)a.partial)[3 * b + 2];
      }
      double* slot = a.vacf_slots + 3 * a.batch_idx;
      if (a.accumulate) {
        slot[0] += tx;
        slot[1] += ty;
        slot[2] += tz;
      } else {
        slot[0] = tx;
        slot[1] = ty;
        slot[2] = tz;
      }
      a.ctrl->ticket = 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Bulk-async (TMA) pipelined variant of the propagate step.
//
// The per-node operands that are streamed exactly once per step -- 18 link probabilities, the
// remaining fraction, u*, the node's own P and its mask word: 204 of the ~230 bytes a fluid node
// moves -- are contiguous runs of BLOCK elements per array.  One elected thread copies them
// global -> shared with cp.async.bulk (UBLKCP), completion counted on an mbarrier, one tile ahead
// of the tile the CTA is working on (2 stages x 51 KB, 2 CTAs per SM).  The memory pipeline is then
// kept full by ~100 KB of copies in flight per SM, independent of registers and occupancy, and the
// threads only issue the 54 neighbour gathers of P (served by L1/L2 thanks to the strip order).
// Arithmetic is unchanged, so results are bit-identical to mp_step_kernel.
// Needs plane % 4 == 0 (16-byte alignment of every run); other lattices use mp_step_kernel.
constexpr int MP_STAGES = 2;
constexpr int MP_STAGE_DOUBLES = 25 * BLOCK;                                   // q[18], s[4], P[3]
constexpr int MP_STAGE_BYTES = MP_STAGE_DOUBLES * 8 + BLOCK * 4;               // + mask
constexpr int MP_SMEM_BYTES = MP_STAGES * MP_STAGE_BYTES + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LBG_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra LBG_DONE;\n\t"
      "bra LBG_WAIT;\n\t"
      "LBG_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// first node and length of a tile (uniform over the CTA)
__device__ __forceinline__ void tile_span(const Geo& geo, int tile, int p_begin, int np, int chunks_per_plane, int& g0,
                                          int& len) {
  const int per_strip = STRIP * np;
  const int strip = tile / per_strip;
  const int rem = tile - strip * per_strip;
  const int c0 = strip * STRIP;
  const int cs = min(STRIP, chunks_per_plane - c0);
  const int p = rem / cs;
  const int c = rem - p * cs;
  const int in_plane = (c0 + c) * BLOCK;
  g0 = (p_begin + p) * geo.plane + in_plane;
  len = min(BLOCK, geo.plane - in_plane);
}

__global__ void __launch_bounds__(BLOCK, 2) mp_step_tma_kernel(const __grid_constant__ MPArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double sh[3][BLOCK / 32];
  __shared__ int s_flag;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // MP_STAGES barriers in the first 64 bytes
  unsigned char* stage0 = smem_raw + 64;
  if (threadIdx.x == 0) {
    int stop = *(volatile int*)&a.ctrl->stop;
    if (!stop && a.check_prev) {
      const volatile double* v = a.vacf_slots + 3 * (a.batch_idx - 1);
      const double ax = fabs(v[0]), ay = fabs(v[1]), az = fabs(v[2]);
      if (ax < a.lim && ay < a.lim && az < a.lim && ax < 1.e-12 && ay < 1.e-12 && az < 1.e-12) {  // :284
        a.ctrl->stop = 1;
        a.ctrl->stop_idx = a.batch_idx;
        stop = 1;
      }
    }
    s_flag = stop;
    for (int st = 0; st < MP_STAGES; ++st) mbar_init(&full[st], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (s_flag) return;

  const Geo& geo = a.geo;
  const long long nalloc = geo.nalloc;
  const int np = a.p_end - a.p_begin;
  const int chunks = (geo.plane + BLOCK - 1) / BLOCK;
  const int ntiles = chunks * np;
  const uint32_t ADS = a.ads ? MASK_INTERFACIAL : 0u;

  auto issue = [&](int tile, int st) {  // one thread
    int g0, len;
    tile_span(geo, tile, a.p_begin, np, chunks, g0, len);
    double* sd = reinterpret_cast<double*>(stage0 + (size_t)st * MP_STAGE_BYTES);
    uint32_t* sm = reinterpret_cast<uint32_t*>(sd + MP_STAGE_DOUBLES);
    mbar_expect_tx(&full[st], (uint32_t)len * (25 * 8 + 4));
#pragma unroll 1
    for (int i = 0; i < 18; ++i) bulk_g2s(sd + i * BLOCK, a.q + (long long)i * nalloc + g0, len * 8, &full[st]);
#pragma unroll 1
    for (int i = 0; i < 4; ++i) bulk_g2s(sd + (18 + i) * BLOCK, a.s + (long long)i * nalloc + g0, len * 8, &full[st]);
#pragma unroll 1
    for (int i = 0; i < 3; ++i) bulk_g2s(sd + (22 + i) * BLOCK, a.Pnow + (long long)i * nalloc + g0, len * 8, &full[st]);
    bulk_g2s(sm, a.mask + g0, len * 4, &full[st]);
  };

  if (threadIdx.x == 0) {
    for (int st = 0; st < MP_STAGES; ++st) {
      const int tile = blockIdx.x + st * gridDim.x;
      if (tile < ntiles) issue(tile, st);
    }
  }

  double vx = 0, vy = 0, vz = 0;
  int k = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
    const int st = k % MP_STAGES;
    const uint32_t parity = (k / MP_STAGES) & 1;
    int g0, len;
    tile_span(geo, tile, a.p_begin, np, chunks, g0, len);
    const double* sd = reinterpret_cast<const double*>(stage0 + (size_t)st * MP_STAGE_BYTES);
    const uint32_t* sm = reinterpret_cast<const uint32_t*>(sd + MP_STAGE_DOUBLES);
    mbar_wait(&full[st], parity);
    const int tid = threadIdx.x;
    if (tid < len) {
      const int g = g0 + tid;
      const uint4 mm = *reinterpret_cast<const uint4*>(sm + (tid & ~3));
      const uint32_t m = sm[tid];
      if ((mm.x | mm.y | mm.z | mm.w) & MASK_FLUID) {
        const bool fluid = m & MASK_FLUID;
        const bool adsorbing = fluid && (m & ADS);
        const bool sector_ads = ADS && (((mm.x & MASK_FLUID) && (mm.x & ADS)) || ((mm.y & MASK_FLUID) && (mm.y & ADS)) ||
                                        ((mm.z & MASK_FLUID) && (mm.z & ADS)) || ((mm.w & MASK_FLUID) && (mm.w & ADS)));
        double nx = 0.0, ny = 0.0, nz = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
        if (fluid) {
          double sx = 0.0, sy = 0.0, sz = 0.0;
          if (adsorbing) {
            sx = a.Anow[g];
            sy = a.Anow[nalloc + g];
            sz = a.Anow[2 * nalloc + g];
          }
          const Nb nb = neighbours(geo, g);
          double ax = 0.0, ay = 0.0, az = 0.0;  // Propagated_Quantity(:,r,next) is always 0 on entry
          static_for<1, NV>([&](auto Lc) {
            constexpr int L = decltype(Lc)::value;
            const int gp = ((m >> L) & 1u) ? g + offset_plus<L>(nb) : g;
            const double q = sd[(L - 1) * BLOCK + tid];
            ax = ax + a.Pnow[gp] * q;
            ay = ay + a.Pnow[nalloc + gp] * q;
            az = az + a.Pnow[2 * nalloc + gp] * q;
          });
          const double frac = sd[18 * BLOCK + tid];
          const double px = sd[22 * BLOCK + tid], py = sd[23 * BLOCK + tid], pz = sd[24 * BLOCK + tid];
          vx += px * sd[19 * BLOCK + tid];  // vacf(:,now) += P(:,r,now)*u_star   (:232)
          vy += py * sd[20 * BLOCK + tid];
          vz += pz * sd[21 * BLOCK + tid];
          if (!adsorbing) {  // :235-238
            nx = ax + frac * px;
            ny = ay + frac * py;
            nz = az + frac * pz;
          } else {  // :239-247
            nx = (ax + frac * px) + sx * a.kd;
            ny = (ay + frac * py) + sy * a.kd;
            nz = (az + frac * pz) + sz * a.kd;
            bx = sx * a.one_minus_kd + px * a.ka;
            by = sy * a.one_minus_kd + py * a.ka;
            bz = sz * a.one_minus_kd + pz * a.ka;
          }
        }
        __stcs(a.Pnext + g, nx);
        __stcs(a.Pnext + nalloc + g, ny);
        __stcs(a.Pnext + 2 * nalloc + g, nz);
        if (sector_ads) {
          __stcs(a.Anext + g, bx);
          __stcs(a.Anext + nalloc + g, by);
          __stcs(a.Anext + 2 * nalloc + g, bz);
        }
      }
    }
    __syncthreads();  // every thread is done with this stage: refill it with the tile MP_STAGES ahead
    if (threadIdx.x == 0) {
      const int next = tile + MP_STAGES * gridDim.x;
      if (next < ntiles) issue(next, st);
    }
  }
  // vacf: block partials, then the last block to finish adds them in block order
  block_sum3(vx, vy, vz, sh);
  if (threadIdx.x == 0) {
    a.partial[3 * blockIdx.x + 0] = vx;
    a.partial[3 * blockIdx.x + 1] = vy;
    a.partial[3 * blockIdx.x + 2] = vz;
    __threadfence();
    const unsigned int done = atomicAdd(&a.ctrl->ticket, 1u);
    if (done == gridDim.x - 1) {
      __threadfence();
      double tx = 0, ty = 0, tz = 0;
      for (unsigned int b = 0; b < gridDim.x; ++b) {
        tx += ((volatile double*)a.partial)[3 * b + 0];
        ty += ((volatile double*)a.partial)[3 * b + 1];
        tz += ((volatile double*)a.partial)[3 * b + 2];
      }
      double* slot = a.vacf_slots + 3 * a.batch_idx;
      if (a.accumulate) {
        slot[0] += tx;
        slot[1] += ty;
        slot[2] += tz;
      } else {
        slot[0] = tx;
        slot[1] = ty;
        slot[2] = tz;
      }
      a.ctrl->ticket = 0;
    }
  }
}

// SoA (3 arrays, stride nalloc, with halos) -> reference AoS (x:z,i,j,k) over own planes
__global__ void __launch_bounds__(BLOCK) soa_to_aos3_kernel(Geo geo, const double* __restrict__ soa,
                                                            double* __restrict__ aos) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    const long long g = q + geo.plane;
    aos[3 * q + 0] = soa[g];
    aos[3 * q + 1] = soa[geo.nalloc + g];
    aos[3 * q + 2] = soa[2 * geo.nalloc + g];
  }
}

int clamp_grid(long long n, int grid) {
  const long long b = (n + BLOCK - 1) / BLOCK;
  return (int)(b < 1 ? 1 : (b < grid ? b : grid));
}

}  // namespace

int launch_mp_init(const MPInitArgs& a, int grid, cudaStream_t st) {
  mp_init_kernel<<<grid, BLOCK, 0, st>>>(a);
  return 1;
}

int launch_mp_step(const MPArgs& a, int variant, int grid, cudaStream_t st) {
  const long long ntiles = (long long)((a.geo.plane + BLOCK - 1) / BLOCK) * (a.p_end - a.p_begin);
  const int gr = (int)(ntiles < grid ? ntiles : grid);
  if (variant == 1) mp_step_tma_kernel<<<gr, BLOCK, MP_SMEM_BYTES, st>>>(a);
  else mp_step_kernel<<<gr, BLOCK, 0, st>>>(a);
  return 1;
}

int launch_soa_to_aos3(const Geo& g, const double* soa, double* aos_own, cudaStream_t st) {
  soa_to_aos3_kernel<<<clamp_grid((long long)g.plane * g.nzl, 148 * 8), BLOCK, 0, st>>>(g, soa, aos_own);
  return 1;
}

int occupancy_grid_mp(int sm_count, int variant) {
  int per_sm = 0;
  if (variant == 1) {
    cudaFuncSetAttribute(mp_step_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MP_SMEM_BYTES);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mp_step_tma_kernel, BLOCK, MP_SMEM_BYTES);
  } else {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mp_step_kernel, BLOCK, 0);
  }
  if (per_sm < 1) per_sm = 1;
  return sm_count * per_sm;
}

}  // namespace lbg
