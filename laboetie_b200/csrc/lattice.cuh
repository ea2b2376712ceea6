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

__device__ __forceinline__ long long first_fid(long long fid_begin) {
  return tile_base(fid_begin) + (long long)blockIdx.x * BLOCK + threadIdx.x;
}

// Tile schedule of a persistent CTA.  A tile is BLOCK consecutive fids starting on a 32-fid boundary
// (tile_base); thread t of the CTA owns fid tile_base + tile*BLOCK + t.  Two modes (Geo::tpc):
//   static  (tpc == 0): CTA b visits tiles b, b + gridDim.x, ...; its "chunk" is all of them (id = b).
//   dynamic (tpc  > 0): chunks of tpc consecutive tiles are handed out by an atomic counter, one chunk
//                       ahead, so that the counter's round trip hides behind the current chunk.
// Measured (profiles/streams_r4a.txt, streams_r4c.txt): a 19 -> 19 fp64 stream moves 5.2-5.8 TB/s with the static
// schedule and 6.4-6.8 TB/s with the dynamic one on the same persistent grid -- the SMs do not all get the same
// share of HBM bandwidth, and a static partition runs at the pace of the slowest.
// All threads of the CTA must call init / advance together (they contain __syncthreads in dynamic mode).
struct Tiles {
  int tile;    // current tile, -1 when the CTA has no more work
  int chunk;   // id of the chunk the current tile belongs to (slot of a per-chunk partial result)
  int k;       // tiles of the current chunk visited before this one
  int cn;      // dynamic: the chunk after this one, already fetched
  int ntiles, nchunks, tpc, par;
  unsigned int* counter;
  unsigned int* slot;  // two words of shared memory

  __device__ __forceinline__ int grab() {
    if (threadIdx.x == 0) slot[par] = atomicAdd(counter, 1u);
    __syncthreads();
    const unsigned int v = slot[par];
    par ^= 1;
    return v < (unsigned int)nchunks ? (int)v : nchunks;
  }
  __device__ __forceinline__ void init(const Geo& geo, long long fid_begin, long long fid_end, unsigned int* counter_,
                                       unsigned int* slot_, int ntiles_override = -1) {
    ntiles = ntiles_override >= 0 ? ntiles_override : (int)((fid_end - tile_base(fid_begin) + BLOCK - 1) / BLOCK);
    tpc = geo.tpc;
    counter = counter_;
    slot = slot_;
    par = 0;
    k = 0;
    if (tpc > 0) {
      nchunks = (ntiles + tpc - 1) / tpc;
      chunk = grab();
      cn = chunk < nchunks ? grab() : nchunks;
      tile = chunk < nchunks ? chunk * tpc : -1;
    } else {
      nchunks = (int)gridDim.x;
      chunk = (int)blockIdx.x;
      cn = nchunks;
      tile = (int)blockIdx.x < ntiles ? (int)blockIdx.x : -1;
    }
  }
  // the tile advance() will move to, -1 if none (lets the caller prefetch for it)
  __device__ __forceinline__ int next_tile() const {
    if (tpc > 0) {
      if (k + 1 < tpc && tile + 1 < ntiles) return tile + 1;
      return cn < nchunks ? cn * tpc : -1;
    }
    const int t = tile + (int)gridDim.x;
    return t < ntiles ? t : -1;
  }
  // move to the next tile; true if the current tile was the last one of its chunk
  __device__ __forceinline__ bool advance() {
    if (tpc > 0) {
      if (k + 1 < tpc && tile + 1 < ntiles) {
        ++k;
        ++tile;
        return false;
      }
      chunk = cn;
      k = 0;
      if (chunk < nchunks) {
        tile = chunk * tpc;
        cn = grab();
      } else {
        tile = -1;
      }
      return true;
    }
    tile += (int)gridDim.x;
    if (tile < ntiles) return false;
    tile = -1;
    return true;
  }
};

// after its last tile every CTA checks in; the last one of the launch resets the counters for the next launch
__device__ __forceinline__ bool cta_checks_in_last(Ctrl* ctrl) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(&ctrl->ticket, 1u);
    s_last = (done == gridDim.x - 1) ? 1 : 0;
    if (s_last) {
      ctrl->tile_next = 0;
      ctrl->ticket = 0;
    }
  }
  __syncthreads();
  return s_last != 0;
}

}  // namespace lbg
