// mp_kernels.cu -- Phase-B kernels for sm_100a: tracer moment propagation.
//
// The flow is frozen during Phase B (drop_tracers.f90:41-49), so everything the
// reference recomputes per step from n(l,i,j,k) and density is static.  mp_init
// evaluates it once, with the reference's own expressions and association order,
// and stores per fluid node (fluid-compacted SoA, fp64):
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
// overwritten every step (solid nodes own no storage).
#include "lattice.cuh"

#ifndef LBG_MP_FENCE
#define LBG_MP_FENCE 1
#endif
#ifndef LBG_MP_MINB
#define LBG_MP_MINB 2
#endif

namespace lbg {
using namespace d3q19;

namespace {

// Streamed operands (read once per step: link probabilities, remaining fraction, u*, table words, adsorbed
// quantity).  LBG_MP_LD selects the cache hint: 0 = ld.global.cs (evict first), 1 = ld.global.cg (L2 only),
// 2 = L1::no_allocate (keeps L1 for the 54 gathers), 3 = L1::no_allocate + L2 evict-first policy.
#ifndef LBG_MP_LD
#define LBG_MP_LD 0
#endif
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
#if LBG_MP_LD == 0
  return __ldcs(p);
#elif LBG_MP_LD == 1
  return __ldcg(p);
#else
  T v;
  if constexpr (sizeof(T) == 8) {
    unsigned long long r;
#if LBG_MP_LD == 2
    asm volatile("ld.global.L1::no_allocate.b64 %0, [%1];" : "=l"(r) : "l"(p));
#else
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.b64 %0, [%1];" : "=l"(r) : "l"(p));
#endif
    v = *reinterpret_cast<T*>(&r);
  } else {
    unsigned int r;
#if LBG_MP_LD == 2
    asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
#else
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.b32 %0, [%1];" : "=r"(r) : "l"(p));
#endif
    v = *reinterpret_cast<T*>(&r);
  }
  return v;
#endif
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
  const long long nfa = geo.nfa;
  const double eps = 2.220446049250313e-16;  // epsilon(1._dp)
  double v0x = 0, v0y = 0, v0z = 0;
  bool bad = false;
  for (long long ff = first_fid(a.fid_begin); ff < a.fid_end; ff += (long long)gridDim.x * BLOCK) {
    if (ff < a.fid_begin) continue;
    const int fid = (int)ff;
    const uint32_t gi = geo.gidx[fid];
    const int g = (int)(gi & GIDX_MASK);
    const Nb nb = neighbours(geo, g);
    const double rho = a.mom[fid];
    const double tx = a.mom[nfa + fid] + a.f[0];
    const double ty = a.mom[2 * nfa + fid] + a.f[1];
    const double tz = a.mom[3 * nfa + fid] + a.f[2];
    double frac = 1.0, usx = 0.0, usy = 0.0, usz = 0.0;
    double px = 0.0, py = 0.0, pz = 0.0;
    uint32_t nbw[8];
    // slow: on the periodic x seam, or a derived index (c - 1, c + 1, fid +- 1) would leave [0, nfa)
    bool slow = (nb.oxm != -1) || (nb.oxp != 1) || fid < 1 || fid + 1 >= nfa;
    // "regular": every fluid neighbour's id is this node's id plus the neighbour's DENSE offset (all nodes in
    // between are fluid -- whole rows and planes of an open geometry), and a solid neighbour's fid + offset stays
    // inside the arrays.  Such a node needs neither rank lookups nor the table in the propagate kernel.
    bool regular = true;
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      constexpr int LI = inv(L);
      double q = 0.0;
      int fp;
      const int off = offset_plus<L>(nb);
      const bool fluid_nb = lookup(geo, g + off, fp);
      const long long guess = (long long)fid + off;
      regular = regular && (fluid_nb ? (long long)fp == guess : (guess >= 0 && guess < nfa));
      if constexpr (cx(L) == 0) {  // centre node of a neighbouring row
        constexpr int R = nbt_row(cy(L), cz(L));
        nbw[R] = (uint32_t)fp | (fluid_nb ? NBT_CENTRE_FLUID : 0u);
        slow = slow || fp < 1 || (long long)fp + 1 >= nfa;
      }
      if (fluid_nb) {  // neighbour r + c_L is fluid
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
        const double rhop = a.mom[fp];
        const double txp = a.mom[nfa + fp] + a.f[0];
        const double typ = a.mom[2 * nfa + fp] + a.f[1];
        const double tzp = a.mom[3 * nfa + fp] + a.f[2];
        q = scattprop<LI>(a.k, a.lambda_w, rhop, txp, typ, tzp);
        // P(:,r,now) += exp_min_dphi*scattprop_p*c_inv(:)*boltz_weight   (:134-135)
        if constexpr (cx(LI) > 0) px = px + q * a.bw;
        if constexpr (cx(LI) < 0) px = px + (-q) * a.bw;
        if constexpr (cy(LI) > 0) py = py + q * a.bw;
        if constexpr (cy(LI) < 0) py = py + (-q) * a.bw;
        if constexpr (cz(LI) > 0) pz = pz + q * a.bw;
        if constexpr (cz(LI) < 0) pz = pz + (-q) * a.bw;
      }
      a.q[(long long)(L - 1) * nfa + fid] = q;
    });
    if (a.ads && (gi & GIDX_INTERFACIAL)) frac = frac - a.ka;  // :240
    if (frac < eps) bad = true;                               // :249
    // pack: rows (0,+-z) absolute, the others as 16-bit deltas (lbg_internal.h NBT_*)
    auto cof = [&](int r) { return (long long)(nbw[r] & NBT_FID_MASK); };
    auto flu = [&](int r) { return (nbw[r] & NBT_CENTRE_FLUID) != 0; };
    const long long d0 = cof(0) - fid, d1 = cof(1) - fid, d4 = cof(4) - cof(2), d5 = cof(5) - cof(2),
                    d6 = cof(6) - cof(3), d7 = cof(7) - cof(3);
    slow = slow || !nbt_delta_fits(d0) || !nbt_delta_fits(d1) || !nbt_delta_fits(d4) || !nbt_delta_fits(d5) ||
           !nbt_delta_fits(d6) || !nbt_delta_fits(d7);
    uint32_t pk[5];
    pk[0] = nbw[2];
    pk[1] = nbw[3];
    pk[2] = nbt_enc16((int)d0, flu(0)) | (nbt_enc16((int)d1, flu(1)) << 16);
    pk[3] = nbt_enc16((int)d4, flu(4)) | (nbt_enc16((int)d5, flu(5)) << 16);
    pk[4] = nbt_enc16((int)d6, flu(6)) | (nbt_enc16((int)d7, flu(7)) << 16);
    if (slow) {  // such a node resolves all its neighbours through the rank structure: it needs g, not the row centres
      pk[0] = NBT_FLAG;
      pk[2] = (uint32_t)g;
    }
    if (a.rwords && regular) atomicOr(a.rwords + (fid >> 5), 1u << ((uint32_t)fid & 31u));
    a.nbt01[fid] = pk[0];
    a.nbt01[nfa + fid] = pk[1];
    for (int r = 2; r < 5; ++r) a.nbt27[(long long)(r - 2) * nfa + fid] = pk[r];
    a.s[fid] = frac;
    a.s[nfa + fid] = usx;
    a.s[2 * nfa + fid] = usy;
    a.s[3 * nfa + fid] = usz;
    a.P0[fid] = px;
    a.P0[nfa + fid] = py;
    a.P0[2 * nfa + fid] = pz;
  }
  if (bad) *a.err = 1;
  block_sum3(v0x, v0y, v0z, sh);
  if (threadIdx.x == 0) {
    a.partial[3 * blockIdx.x + 0] = v0x;
    a.partial[3 * blockIdx.x + 1] = v0y;
    a.partial[3 * blockIdx.x + 2] = v0z;
  }
}

// module_moment_propagation.f90:207-253 (one propagate call), see the header comment.
// Branch-free per node: the 18 link probabilities, the remaining fraction, u* and the node's own
// P are streamed in first (25 independent loads in flight), then the 18 neighbour gathers are
// issued without conditions -- a solid neighbour is replaced by the node itself and its q is 0
// (mp_init stores 0 there), so it adds exactly 0.
// NBT: neighbour fluid ids from the static table (lbg_internal.h NBT_*) instead of 18 rank lookups per node.
template <bool NBT>
__global__ void __launch_bounds__(BLOCK, LBG_MP_MINB) mp_step_kernel(const __grid_constant__ MPArgs a) {
  __shared__ double sh[3][BLOCK / 32];
  __shared__ int s_flag;
  if (threadIdx.x == 0) {
    int stop = *(volatile int*)&a.ctrl->stop;
    if (!stop && a.check_slot >= 0) {
      const volatile double* v = a.vacf_slots + 3 * a.check_slot;
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
  const long long nfa = geo.nfa;
  double vx = 0, vy = 0, vz = 0;
  // Tile order.  Plain: tiles of BLOCK consecutive fids.  Strip order (nseg > 0): the own planes are cut
  // into strips of rows and all planes of a strip are visited before the next strip, so the three
  // planes a node gathers P from were touched a few MB of traffic ago and are still in L2: P is then
  // read from HBM once per step instead of up to three times.
  const long long base = tile_base(a.fid_begin);  // tiles start on a 32-fid boundary (see first_fid)
  const int ntiles = a.nseg > 0 ? a.ntiles : (int)((a.fid_end - base + BLOCK - 1) / BLOCK);
  int seg = 0;
  auto node_of = [&](int tile, int& k) -> long long {  // fid of this thread in `tile`, or -1
    if (a.nseg == 0) {
      const long long f = base + (long long)tile * BLOCK + threadIdx.x;
      return (f >= a.fid_begin && f < a.fid_end) ? f : -1;
    }
    while (tile >= a.tile_cum[k + 1]) ++k;  // tiles are visited in increasing order
    const long long f = a.seg_begin[k] + (long long)(tile - a.tile_cum[k]) * BLOCK + threadIdx.x;
    return f < a.seg_end[k] ? f : -1;
  };
  // Tile schedule.  tpc == 0: persistent grid, CTA b visits tiles b, b + gridDim.x, ...  tpc > 0: the grid covers
  // the tiles, CTA b owns the tpc consecutive tiles from b * tpc -- the hardware block scheduler then balances the
  // SMs (they do not all get the same share of HBM bandwidth; profiles/streams_r4a.txt) and each stream a CTA
  // reads is one contiguous run of tpc * 2 KB.
  const int tile0 = a.tpc > 0 ? (int)blockIdx.x * a.tpc : (int)blockIdx.x;
  const int tstep = a.tpc > 0 ? 1 : (int)gridDim.x;
  const int tend = a.tpc > 0 ? (tile0 + a.tpc < ntiles ? tile0 + a.tpc : ntiles) : ntiles;
  long long f_next = tile0 < tend ? node_of(tile0, seg) : -1;
  // neighbour-table words (NBT) or the dense index (gidx) of the next tile are fetched one iteration
  // ahead: the gathers depend on them
  constexpr int NW = 5;
  uint32_t w_next[NW] = {0, 0, 0, 0, 0};
  uint2 aw_next = make_uint2(0u, 0u);
  auto load_words = [&](long long f) {
    if (a.ads) aw_next = __ldg(a.awords + (f >> 5));
    if constexpr (NBT) {
      w_next[0] = ld_stream(a.nbt01 + f);
      w_next[1] = ld_stream(a.nbt01 + nfa + f);
#pragma unroll
      for (int r = 2; r < NW; ++r) w_next[r] = ld_stream(a.nbt27 + (long long)(r - 2) * nfa + f);
    } else {
      w_next[0] = __ldg(geo.gidx + f);
      w_next[1] = a.rwords ? __ldg(a.rwords + (f >> 5)) : 0u;
    }
  };
  if (f_next >= 0) load_words(f_next);
  for (int tile = tile0; tile < tend; tile += tstep) {
    const long long ff = f_next;
    const int tn = tile + tstep;
    uint32_t w[NW];
#pragma unroll
    for (int r = 0; r < NW; ++r) w[r] = w_next[r];
    const uint2 aw = aw_next;
    f_next = tn < tend ? node_of(tn, seg) : -1;
    if (f_next >= 0) load_words(f_next);
    if (ff < 0) continue;
    const int fid = (int)ff;
    // adsorbing nodes and their slot in the compact adsorbed arrays: one group word per warp (lbg_internal.h)
    bool adsorbing = false;
    long long aslot = 0;
    int apad = 0;   // padding slots this lane zeroes (whole-sector stores)
    if (a.ads) {
      const uint32_t bit = (uint32_t)fid & 31u;
      adsorbing = (aw.x >> bit) & 1u;
      const int cnt = __popc(aw.x), below = __popc(aw.x & ((1u << bit) - 1u));
      if (adsorbing) {
        aslot = (long long)aw.y + below;
      } else {  // the k-th non-adsorbing lane of the group zeroes the k-th padding slot
        const int k = (int)bit - below;
        apad = k < ((cnt + 3) & ~3) - cnt ? 1 : 0;
        aslot = (long long)aw.y + cnt + k;
      }
    }
    // fluid ids of the 18 neighbours (a solid one -> any in-range id: its q is 0 and adds exactly 0)
    int gp[NV];
    const double* Pn = a.Pnow;
    auto resolve_by_lookup = [&](int g) {
      // Two phases, as in the LB pull: resolve all 18 rank lookups first, then gather.  geo.zero is 0
      // at run time and makes every gather address depend on every lookup, which pins that order in
      // the schedule (a lookup word consumed after the first gathers were issued would share a
      // scoreboard with them and wait for their latency as well).
      const Nb nb = neighbours(geo, g);
      uint32_t any = 0;
      static_for<1, NV>([&](auto Lc) {
        constexpr int L = decltype(Lc)::value;
        int fp;
        const bool fl = lookup(geo, g + offset_plus<L>(nb), fp);
        gp[L] = fl ? fp : fid;
        any |= (uint32_t)gp[L];
      });
#if LBG_MP_FENCE
      Pn = a.Pnow + (any & (uint32_t)geo.zero);
#endif
    };
    if constexpr (NBT) {
      if (w[0] & NBT_FLAG) {  // periodic x seam (few nodes): through the rank structure, word 2 holds g
        resolve_by_lookup((int)(w[2] & GIDX_MASK));
      } else {
        // row centres and their "centre is fluid" bits from the packed words
        int rc[8], rf[8];
        rc[2] = (int)(w[0] & NBT_FID_MASK);
        rf[2] = (int)(w[0] >> 31);
        rc[3] = (int)(w[1] & NBT_FID_MASK);
        rf[3] = (int)(w[1] >> 31);
        rc[0] = fid + nbt_dec16(w[2]);
        rf[0] = (int)((w[2] >> 15) & 1u);
        rc[1] = fid + nbt_dec16(w[2] >> 16);
        rf[1] = (int)(w[2] >> 31);
        rc[4] = rc[2] + nbt_dec16(w[3]);
        rf[4] = (int)((w[3] >> 15) & 1u);
        rc[5] = rc[2] + nbt_dec16(w[3] >> 16);
        rf[5] = (int)(w[3] >> 31);
        rc[6] = rc[3] + nbt_dec16(w[4]);
        rf[6] = (int)((w[4] >> 15) & 1u);
        rc[7] = rc[3] + nbt_dec16(w[4] >> 16);
        rf[7] = (int)(w[4] >> 31);
        static_for<1, NV>([&](auto Lc) {
          constexpr int L = decltype(Lc)::value;
          constexpr int R = nbt_row(cy(L), cz(L));
          if constexpr (R < 0) {
            gp[L] = fid + cx(L);
          } else {
            if constexpr (cx(L) == 0) gp[L] = rc[R];
            else if constexpr (cx(L) > 0) gp[L] = rc[R] + rf[R];
            else gp[L] = rc[R] - 1;
          }
        });
      }
    } else {
      const int g = (int)(w[0] & GIDX_MASK);
      if ((w[1] >> ((uint32_t)fid & 31u)) & 1u) {
        // regular node (mp_init): neighbour ids by arithmetic, no lookups -- on open geometries (slit, bulk,
        // wide channels) that is every warp except those next to a wall plane
        const Nb nb = neighbours(geo, g);
        static_for<1, NV>([&](auto Lc) {
          constexpr int L = decltype(Lc)::value;
          gp[L] = fid + offset_plus<L>(nb);
        });
      } else {
        resolve_by_lookup(g);
      }
    }
    double q[NV - 1];
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      q[L - 1] = ld_stream(a.q + (long long)(L - 1) * nfa + fid);
    });
    const double frac = ld_stream(a.s + fid);
    const double usx = ld_stream(a.s + nfa + fid), usy = ld_stream(a.s + 2 * nfa + fid), usz = ld_stream(a.s + 3 * nfa + fid);
    const double px = a.Pnow[fid], py = a.Pnow[nfa + fid], pz = a.Pnow[2 * nfa + fid];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    if (adsorbing) {
      sx = ld_stream(a.Anow + aslot);
      sy = ld_stream(a.Anow + a.a_stride + aslot);
      sz = ld_stream(a.Anow + 2 * a.a_stride + aslot);
    }
    double ax = 0.0, ay = 0.0, az = 0.0;  // Propagated_Quantity(:,r,next) is always 0 on entry
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      const double* p = Pn + gp[L];
      ax = ax + p[0] * q[L - 1];
      ay = ay + p[nfa] * q[L - 1];
      az = az + p[2 * nfa] * q[L - 1];
    });
    vx += px * usx;  // vacf(:,now) += P(:,r,now)*u_star   (:232)
    vy += py * usy;
    vz += pz * usz;
    // the adsorbed quantity lives in the group's own whole sectors: the adsorbing lanes store their slots, up
    // to three other lanes store 0 into the padding, so no sector is written partially (no read-fill)
    const bool write_ads = adsorbing || apad;
    double nx, ny, nz, bx = 0.0, by = 0.0, bz = 0.0;
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
    __stcs(a.Pnext + fid, nx);
    __stcs(a.Pnext + nfa + fid, ny);
    __stcs(a.Pnext + 2 * nfa + fid, nz);
    if (write_ads) {
      __stcs(a.Anext + aslot, bx);
      __stcs(a.Anext + a.a_stride + aslot, by);
      __stcs(a.Anext + 2 * a.a_stride + aslot, bz);
    }
  }
  // vacf: block partials, then the last CTA to finish adds them in a fixed order with all its threads (thread t
  // takes the partials b = t (mod BLOCK) in increasing order, then the BLOCK sums are combined as in block_sum3):
  // deterministic run to run, and no single thread walks 296 x 3 values (visible on 70 us launches).
  block_sum3(vx, vy, vz, sh);
  if (threadIdx.x == 0) {
    a.partial[3 * (size_t)blockIdx.x + 0] = vx;
    a.partial[3 * (size_t)blockIdx.x + 1] = vy;
    a.partial[3 * (size_t)blockIdx.x + 2] = vz;
    __threadfence();
    const unsigned int done = atomicAdd(&a.ctrl->ticket, 1u);
    s_flag = (done == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  double tx = 0, ty = 0, tz = 0;
  for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLOCK) {
    tx += __ldcg(a.partial + 3 * (size_t)b + 0);
    ty += __ldcg(a.partial + 3 * (size_t)b + 1);
    tz += __ldcg(a.partial + 3 * (size_t)b + 2);
  }
  block_sum3(tx, ty, tz, sh);
  if (threadIdx.x == 0) {
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

}  // namespace

int launch_mp_init(const MPInitArgs& a, int grid, cudaStream_t st) {
  mp_init_kernel<<<grid, BLOCK, 0, st>>>(a);
  return 1;
}

int launch_mp_step(const MPArgs& a, int grid, cudaStream_t st) {
  int gr;
  if (a.nseg > 0) {
    if (a.ntiles <= 0) return 0;
    gr = a.ntiles < grid ? a.ntiles : grid;
    if (a.tpc > 0) gr = (a.ntiles + a.tpc - 1) / a.tpc;
  } else {
    if (a.fid_end <= a.fid_begin) return 0;
    gr = clamp_grid(a.fid_end - a.fid_begin, grid);
    if (a.tpc > 0) {
      const long long nt = (a.fid_end - tile_base(a.fid_begin) + BLOCK - 1) / BLOCK;
      gr = (int)((nt + a.tpc - 1) / a.tpc);
    }
  }
  if (a.use_nbt) mp_step_kernel<true><<<gr, BLOCK, 0, st>>>(a);
  else mp_step_kernel<false><<<gr, BLOCK, 0, st>>>(a);
  return 1;
}

int occupancy_grid_mp(int sm_count) {
  int per_sm = 0, per_sm2 = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mp_step_kernel<true>, BLOCK, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, mp_step_kernel<false>, BLOCK, 0);
  if (per_sm2 < per_sm) per_sm = per_sm2;
  if (per_sm < 1) per_sm = 1;
  return sm_count * per_sm;
}

}  // namespace lbg
