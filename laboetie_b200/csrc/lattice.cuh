// lattice.cuh -- device helpers shared by the kernels: compile-time loops, periodic neighbour
// offsets in the dense node order (module_system.f90:99-112 `pbc`), and the dense -> fluid-id lookup.
#pragma once
#include <type_traits>

#include "lbg_internal.h"

namespace lbg {

template <int L, int END, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (L < END) {
    f(std::integral_constant<int, L>{});
    static_for<L + 1, END>(f);
  }
}

// neighbour offsets of one node in the dense linear index (periodic in x and y; in z either periodic
// inside the slab (zwrap, single slab) or plain, the halo planes standing in for the neighbours)
struct Nb {
  int oxm, oxp, oym, oyp, ozm, ozp;
};

// floor(g / d) for 0 <= g < 2^31 with the multiplier of div_magic (lbg_internal.h)
__device__ __forceinline__ int fast_div(int g, uint32_t mul, int sh) {
  return (int)(((unsigned long long)(uint32_t)g * mul) >> sh);
}

__device__ __forceinline__ Nb neighbours(const Geo& geo, int g) {
  const int p = fast_div(g, geo.mul_plane, geo.sh_plane);
  const int rem = g - p * geo.plane;
  const int y = fast_div(rem, geo.mul_lx, geo.sh_lx);
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

// dense offset of node r + c_L
template <int L>
__device__ __forceinline__ int offset_plus(const Nb& nb) {
  constexpr int X = d3q19::cx(L), Y = d3q19::cy(L), Z = d3q19::cz(L);
  int o = 0;
  if constexpr (X > 0) o += nb.oxp;
  if constexpr (X < 0) o += nb.oxm;
  if constexpr (Y > 0) o += nb.oyp;
  if constexpr (Y < 0) o += nb.oym;
  if constexpr (Z > 0) o += nb.ozp;
  if constexpr (Z < 0) o += nb.ozm;
  return o;
}

// is dense node g fluid, and which fluid id does it have
__device__ __forceinline__ bool lookup(const Geo& geo, int g, int& fid) {
  const uint2 w = __ldg(geo.words + (g >> 5));
  const uint32_t bit = (uint32_t)g & 31u;
  fid = (int)(w.y + __popc(w.x & ((1u << bit) - 1u)));
  return (w.x >> bit) & 1u;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

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
__device__ __forceinline__ long long first_fid(long long fid_begin) {
  return tile_base(fid_begin) + (long long)blockIdx.x * BLOCK + threadIdx.x;
}

inline int clamp_grid(long long n, int grid) {
  const long long b = (n + 31 + BLOCK - 1) / BLOCK;  // + 31: tiles start on the 32-fid boundary below fid_begin
  return (int)(b < 1 ? 1 : (b < grid ? b : grid));
}

}  // namespace lbg
