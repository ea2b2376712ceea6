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

// module_moment_propagation.f90:207-253 (one propagate call), see the header comment.
__global__ void __launch_bounds__(BLOCK) mp_step_kernel(const __grid_constant__ MPArgs a) {
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
  double vx = 0, vy = 0, vz = 0;
  for (long long gg = a.g_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; gg < a.g_end;
       gg += (long long)gridDim.x * BLOCK) {
    const int g = (int)gg;
    const uint32_t m = __ldg(a.mask + g);
    if (!(m & MASK_FLUID)) continue;
    const Nb nb = neighbours(geo, g);
    double ax = 0.0, ay = 0.0, az = 0.0;  // Propagated_Quantity(:,r,next) is always 0 on entry
    static_for<1, NV>([&](auto Lc) {
      constexpr int L = decltype(Lc)::value;
      if ((m >> L) & 1u) {
        const int gp = g + offset_plus<L>(nb);
        const double q = a.q[(long long)(L - 1) * nalloc + g];
        ax = ax + a.Pnow[gp] * q;
        ay = ay + a.Pnow[nalloc + gp] * q;
        az = az + a.Pnow[2 * nalloc + gp] * q;
      }
    });
    const double frac = a.s[g];
    const double px = a.Pnow[g], py = a.Pnow[nalloc + g], pz = a.Pnow[2 * nalloc + g];
    vx += px * a.s[nalloc + g];  // vacf(:,now) += P(:,r,now)*u_star   (:232)
    vy += py * a.s[2 * nalloc + g];
    vz += pz * a.s[3 * nalloc + g];
    if (!(a.ads && (m & MASK_INTERFACIAL))) {  // :235-238
      a.Pnext[g] = ax + frac * px;
      a.Pnext[nalloc + g] = ay + frac * py;
      a.Pnext[2 * nalloc + g] = az + frac * pz;
    } else {  // :239-247
      const double sx = a.Anow[g], sy = a.Anow[nalloc + g], sz = a.Anow[2 * nalloc + g];
      a.Pnext[g] = (ax + frac * px) + sx * a.kd;
      a.Pnext[nalloc + g] = (ay + frac * py) + sy * a.kd;
      a.Pnext[2 * nalloc + g] = (az + frac * pz) + sz * a.kd;
      a.Anext[g] = sx * a.one_minus_kd + px * a.ka;
      a.Anext[nalloc + g] = sy * a.one_minus_kd + py * a.ka;
      a.Anext[2 * nalloc + g] = sz * a.one_minus_kd + pz * a.ka;
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

int launch_mp_step(const MPArgs& a, int grid, cudaStream_t st) {
  mp_step_kernel<<<grid, BLOCK, 0, st>>>(a);
  return 1;
}

int launch_soa_to_aos3(const Geo& g, const double* soa, double* aos_own, cudaStream_t st) {
  soa_to_aos3_kernel<<<clamp_grid((long long)g.plane * g.nzl, 148 * 8), BLOCK, 0, st>>>(g, soa, aos_own);
  return 1;
}

int occupancy_grid_mp(int sm_count) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mp_step_kernel, BLOCK, 0);
  if (per_sm < 1) per_sm = 1;
  return sm_count * per_sm;
}

}  // namespace lbg
