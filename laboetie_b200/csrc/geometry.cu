// geometry.cu -- builds the fluid-compacted numbering from the driver's `nature` array, and moves
// fields between the driver's dense (i,j,k) arrays and the compact storage.
//
// Replaces, on the device: detectInterfacialNodes (supercell_definition.f90:115-147) and the il/jl/kl
// neighbour tables of equilibration.f90:109-119.
#include <cub/block/block_scan.cuh>

#include "lattice.cuh"

namespace lbg {
using namespace d3q19;

namespace {

// one warp per 32 dense nodes: fluid bits by ballot, count in .y (turned into a rank by the scan)
__global__ void __launch_bounds__(BLOCK) build_bits_kernel(long long ndense, const int8_t* __restrict__ nat,
                                                           uint2* __restrict__ words, long long nwords) {
  const long long warp0 = ((long long)blockIdx.x * BLOCK + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * BLOCK) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long w = warp0; w < nwords; w += nwarps) {
    const long long g = w * 32 + lane;
    const bool fluid = (g < ndense) && (nat[g] == 0);  // module_system.f90:32: fluid = 0
    const uint32_t b = __ballot_sync(0xffffffffu, fluid);
    if (lane == 0) words[w] = make_uint2(b, (uint32_t)__popc(b));
  }
}

constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = BLOCK * SCAN_ITEMS;

// pass 1: per-tile totals of the counts; pass 3: exclusive scan inside each tile plus the tile offset
__global__ void __launch_bounds__(BLOCK) scan_tile_sums_kernel(const uint2* __restrict__ words, long long nwords,
                                                               unsigned long long* __restrict__ tile_sums) {
  using BS = cub::BlockScan<unsigned int, BLOCK>;
  __shared__ typename BS::TempStorage tmp;
  const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  unsigned int v = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i)
    if (base + i < nwords) v += words[base + i].y;
  unsigned int incl, total;
  BS(tmp).InclusiveSum(v, incl, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(BLOCK) scan_tile_offsets_kernel(unsigned long long* __restrict__ tile_sums, int ntiles,
                                                                  unsigned long long* __restrict__ total) {
  // one block: serial over chunks of BLOCK tiles (ntiles is a few thousand at most)
  using BS = cub::BlockScan<unsigned long long, BLOCK>;
  __shared__ typename BS::TempStorage tmp;
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < ntiles; c0 += BLOCK) {
    const int i = c0 + threadIdx.x;
    const unsigned long long v = i < ntiles ? tile_sums[i] : 0ull;
    unsigned long long excl, tot;
    BS(tmp).ExclusiveSum(v, excl, tot);
    if (i < ntiles) tile_sums[i] = carry + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(BLOCK) scan_apply_kernel(uint2* __restrict__ words, long long nwords,
                                                           const unsigned long long* __restrict__ tile_offsets) {
  using BS = cub::BlockScan<unsigned int, BLOCK>;
  __shared__ typename BS::TempStorage tmp;
  const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  unsigned int c[SCAN_ITEMS], v = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    c[i] = (base + i < nwords) ? words[base + i].y : 0u;
    v += c[i];
  }
  unsigned int excl;
  BS(tmp).ExclusiveSum(v, excl);
  unsigned int run = excl + (unsigned int)tile_offsets[blockIdx.x];
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < nwords) words[base + i].y = run;
    run += c[i];
  }
}

// interfacial flag of a dense node of an own plane: one of its 18 neighbours has the other nature.
// The halo planes hold the true periodic neighbours, so z is not wrapped here.
__device__ __forceinline__ bool node_interfacial(const Geo& geo, int g, bool me_fluid) {
  Geo gz = geo;
  gz.zwrap = 0;
  const Nb nb = neighbours(gz, g);
  bool itf = false;
  static_for<1, NV>([&](auto Lc) {
    constexpr int L = decltype(Lc)::value;
    int fidn;
    const bool other = lookup(geo, g + offset_plus<L>(nb), fidn);
    itf = itf || (other != me_fluid);
  });
  return itf;
}

__global__ void __launch_bounds__(BLOCK) build_gidx_kernel(Geo geo, long long ndense, uint32_t* __restrict__ gidx) {
  for (long long gg = (long long)blockIdx.x * BLOCK + threadIdx.x; gg < ndense; gg += (long long)gridDim.x * BLOCK) {
    const int g = (int)gg;
    int fid;
    if (!lookup(geo, g, fid)) continue;
    const int p = fast_div(g, geo.mul_plane, geo.sh_plane);
    uint32_t v = (uint32_t)g;
    if (p >= 1 && p <= geo.nzl && node_interfacial(geo, g, true)) v |= GIDX_INTERFACIAL;
    gidx[fid] = v;
  }
}

// number of fluid nodes before dense node idx[i] (idx[i] == ndense allowed)
__global__ void __launch_bounds__(BLOCK) rank_at_kernel(Geo geo, const long long* __restrict__ idx, int n, long long ndense,
                                                        long long total, long long* __restrict__ out) {
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= n) return;
  const long long g = idx[i];
  if (g >= ndense) {
    out[i] = total;
    return;
  }
  const uint2 w = geo.words[g >> 5];
  const unsigned bit = (unsigned)(g & 31);
  out[i] = (long long)w.y + __popc(w.x & ((1u << bit) - 1u));
}

__global__ void __launch_bounds__(BLOCK) count_interfacial_kernel(Geo geo, long long fid_begin, long long fid_end,
                                                                  unsigned long long* count) {
  unsigned int n = 0;
  for (long long f = fid_begin + (long long)blockIdx.x * BLOCK + threadIdx.x; f < fid_end; f += (long long)gridDim.x * BLOCK)
    n += (geo.gidx[f] & GIDX_INTERFACIAL) ? 1u : 0u;
  n = __reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, (unsigned long long)n);
}

// supercell_definition.f90:115-147 for every node of the own planes (solid ones included)
__global__ void __launch_bounds__(BLOCK) dense_interfacial_kernel(Geo geo, int8_t* __restrict__ out) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    const int g = (int)(q + geo.plane);
    int fid;
    const bool fl = lookup(geo, g, fid);
    out[q] = node_interfacial(geo, g, fl) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(BLOCK) scatter_to_dense_kernel(Geo geo, const double* __restrict__ arr,
                                                                 double* __restrict__ dense) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    int fid;
    const bool fl = lookup(geo, (int)(q + geo.plane), fid);
    dense[q] = fl ? arr[fid] : 0.0;
  }
}

__global__ void __launch_bounds__(BLOCK) gather_from_dense_kernel(Geo geo, const double* __restrict__ dense,
                                                                  double* __restrict__ arr) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    int fid;
    if (lookup(geo, (int)(q + geo.plane), fid)) arr[fid] = dense[q];
  }
}

// n(t)(r, l) of the reference rebuilt from the post-collision populations n*(t) by the pull rule
// (equilibration.f90:204-243 in closed form, see lb_kernels.cu), written straight into the driver's dense
// (i,j,k) order for one direction l: a read-back that needs no population-sized scratch and leaves the
// stepping state untouched.
__global__ void __launch_bounds__(BLOCK) pull_to_dense_kernel(Geo geo, const double* __restrict__ fin, int l,
                                                              double* __restrict__ dense) {
  const long long nown = (long long)geo.plane * geo.nzl;
  const int X = d3q19::cx(l), Y = d3q19::cy(l), Z = d3q19::cz(l), li = d3q19::inv(l);
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    const int g = (int)(q + geo.plane);
    int fid;
    double v = 0.0;
    if (lookup(geo, g, fid)) {
      const Nb nb = neighbours(geo, g);
      // source node r - c_l
      const int o = (X > 0 ? nb.oxm : (X < 0 ? nb.oxp : 0)) + (Y > 0 ? nb.oym : (Y < 0 ? nb.oyp : 0)) +
                    (Z > 0 ? nb.ozm : (Z < 0 ? nb.ozp : 0));
      int fsrc;
      const bool src_fluid = lookup(geo, g + o, fsrc);
      v = src_fluid ? fin[(long long)l * geo.nfa + fsrc] : fin[(long long)li * geo.nfa + fid];
    }
    dense[q] = v;
  }
}

// one plane of the moments (equilibration.f90:526-548: the 2-D field outputs): axis 0 -> x = index, out(j,k);
// axis 1 -> y = index, out(i,k); axis 2 -> own plane k = index, out(i,j); first index fastest, 0 on solid nodes.
__global__ void __launch_bounds__(BLOCK) slice_kernel(Geo geo, const double* __restrict__ mom, int axis, int index,
                                                      double* __restrict__ out, int ncomp) {
  const int n1 = axis == 0 ? geo.ly : geo.lx;
  const int n2 = axis == 2 ? geo.ly : geo.nzl;
  const long long n = (long long)n1 * n2;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < n; q += (long long)gridDim.x * BLOCK) {
    const int b = (int)(q / n1), a = (int)(q - (long long)b * n1);
    int i, j, k;
    if (axis == 0) { i = index; j = a; k = b; }
    else if (axis == 1) { i = a; j = index; k = b; }
    else { i = a; j = b; k = index; }
    const long long g = (long long)i + (long long)geo.lx * j + (long long)geo.plane * (k + 1);
    int fid;
    const bool fl = lookup(geo, (int)g, fid);
    for (int c = 0; c < ncomp; ++c) out[(long long)c * n + q] = fl ? mom[(long long)c * geo.nfa + fid] : 0.0;
  }
}

// three SoA arrays -> the reference's AoS (x:z,i,j,k) over the own planes
__global__ void __launch_bounds__(BLOCK) scatter3_aos_kernel(Geo geo, const double* __restrict__ soa,
                                                             double* __restrict__ aos) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    int fid;
    const bool fl = lookup(geo, (int)(q + geo.plane), fid);
    aos[3 * q + 0] = fl ? soa[fid] : 0.0;
    aos[3 * q + 1] = fl ? soa[geo.nfa + fid] : 0.0;
    aos[3 * q + 2] = fl ? soa[2 * geo.nfa + fid] : 0.0;
  }
}

// compact adsorbed storage (lbg_internal.h): one warp per group of 32 fids
__global__ void __launch_bounds__(BLOCK) build_awords_kernel(Geo geo, long long fid_begin, long long fid_end,
                                                             uint2* __restrict__ awords) {
  const long long ngroups = geo.nfa >> 5;
  const long long warp0 = ((long long)blockIdx.x * BLOCK + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * BLOCK) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long w = warp0; w < ngroups; w += nwarps) {
    const long long f = w * 32 + lane;
    const bool itf = f >= fid_begin && f < fid_end && (geo.gidx[f] & GIDX_INTERFACIAL);
    const uint32_t b = __ballot_sync(0xffffffffu, itf);
    if (lane == 0) awords[w] = make_uint2(b, ((uint32_t)__popc(b) + 3u) & ~3u);
  }
}

__global__ void __launch_bounds__(BLOCK) scatter3_compact_aos_kernel(Geo geo, const uint2* __restrict__ awords,
                                                                     const double* __restrict__ a3, long long a_stride,
                                                                     double* __restrict__ aos) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    int fid;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    if (awords && lookup(geo, (int)(q + geo.plane), fid)) {
      const uint2 w = awords[fid >> 5];
      const uint32_t bit = (uint32_t)fid & 31u;
      if ((w.x >> bit) & 1u) {
        const long long slot = (long long)w.y + __popc(w.x & ((1u << bit) - 1u));
        v0 = a3[slot];
        v1 = a3[a_stride + slot];
        v2 = a3[2 * a_stride + slot];
      }
    }
    aos[3 * q + 0] = v0;
    aos[3 * q + 1] = v1;
    aos[3 * q + 2] = v2;
  }
}

// supercell_definition.f90:50-59 on the device, for the labels whose thresholds are exact in integer
// arithmetic: -1 bulk, 1 slit (module_geometry.f90:158-166), 2 cylinder along z (:253-277: solid iff
// |r - (l+1)/2| >= (lx-1)/2), 3 BCC spheres (:206-245: solid iff the distance to a cube corner or to the
// centre is <= (lx-1)*sqrt(3)/4  <=>  16 d^2 <= 3 (lx-1)^2).  Writes planes k0-1 .. k0+nzl (periodic).
__global__ void __launch_bounds__(BLOCK) build_nature_kernel(int label, int lx, int ly, int lz, int k0, int nzl,
                                                             int8_t* __restrict__ nat) {
  const long long plane = (long long)lx * ly, n = plane * (nzl + 2);
  for (long long g = (long long)blockIdx.x * BLOCK + threadIdx.x; g < n; g += (long long)gridDim.x * BLOCK) {
    const int p = (int)(g / plane);
    const int rem = (int)(g - (long long)p * plane);
    const int j = rem / lx, i = rem - j * lx;
    int k = (k0 - 1 + p) % lz;
    if (k < 0) k += lz;
    bool solid = false;
    if (label == 1) {
      solid = (k == 0) || (k == lz - 1);
    } else if (label == 2) {
      const long long dx = 2LL * (i + 1) - (lx + 1), dy = 2LL * (j + 1) - (ly + 1);
      solid = dx * dx + dy * dy >= (long long)(lx - 1) * (lx - 1);
    } else if (label == 3) {
      const long long thr = 3LL * (lx - 1) * (lx - 1);
      const long long X = 2LL * (i + 1), Y = 2LL * (j + 1), Z = 2LL * (k + 1);
      const long long cx2[2] = {2, 2LL * lx}, cy2[2] = {2, 2LL * ly}, cz2[2] = {2, 2LL * lz};
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
          for (int c = 0; c < 2; ++c) {
            const long long d2 = (X - cx2[a]) * (X - cx2[a]) + (Y - cy2[b]) * (Y - cy2[b]) + (Z - cz2[c]) * (Z - cz2[c]);
            solid = solid || (4 * d2 <= thr);
          }
      const long long mx = lx + 1, my = ly + 1, mz = lz + 1;
      solid = solid || (4 * ((X - mx) * (X - mx) + (Y - my) * (Y - my) + (Z - mz) * (Z - mz)) <= thr);
    }
    nat[g] = solid ? 1 : 0;
  }
}

// node%nature of the own planes back from the rank structure
__global__ void __launch_bounds__(BLOCK) dense_nature_kernel(Geo geo, int8_t* __restrict__ out) {
  const long long nown = (long long)geo.plane * geo.nzl;
  for (long long q = (long long)blockIdx.x * BLOCK + threadIdx.x; q < nown; q += (long long)gridDim.x * BLOCK) {
    int fid;
    out[q] = lookup(geo, (int)(q + geo.plane), fid) ? 0 : 1;
  }
}

inline int big_grid(long long n) {
  const long long b = (n + BLOCK - 1) / BLOCK;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace

int launch_build_bits(int plane, int nzl, const int8_t* nature_halo, uint2* words, long long nwords, cudaStream_t st) {
  const long long ndense = (long long)plane * (nzl + 2);
  build_bits_kernel<<<big_grid(nwords * 32), BLOCK, 0, st>>>(ndense, nature_halo, words, nwords);
  return 1;
}

int launch_build_nature(int label, int lx, int ly, int lz, int k0, int nzl, int8_t* nature_halo, cudaStream_t st) {
  build_nature_kernel<<<big_grid((long long)lx * ly * (nzl + 2)), BLOCK, 0, st>>>(label, lx, ly, lz, k0, nzl, nature_halo);
  return 1;
}

int launch_dense_nature(const Geo& g, int8_t* out_own, cudaStream_t st) {
  dense_nature_kernel<<<big_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, out_own);
  return 1;
}

int launch_scan_ranks(uint2* words, long long nwords, unsigned long long* total, cudaStream_t st) {
  const int ntiles = (int)((nwords + SCAN_TILE - 1) / SCAN_TILE);
  unsigned long long* tile_sums = nullptr;
  cudaMallocAsync(&tile_sums, (size_t)ntiles * sizeof(unsigned long long), st);
  scan_tile_sums_kernel<<<ntiles, BLOCK, 0, st>>>(words, nwords, tile_sums);
  scan_tile_offsets_kernel<<<1, BLOCK, 0, st>>>(tile_sums, ntiles, total);
  scan_apply_kernel<<<ntiles, BLOCK, 0, st>>>(words, nwords, tile_sums);
  cudaFreeAsync(tile_sums, st);
  return 3;
}

int launch_build_gidx(const Geo& g, long long nwords, uint32_t* gidx, cudaStream_t st) {
  const long long ndense = (long long)g.plane * (g.nzl + 2);
  (void)nwords;
  build_gidx_kernel<<<big_grid(ndense), BLOCK, 0, st>>>(g, ndense, gidx);
  return 1;
}

int launch_rank_at(const Geo& g, const long long* dense_idx, int n, long long ndense, long long total, long long* out,
                   cudaStream_t st) {
  rank_at_kernel<<<(n + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(g, dense_idx, n, ndense, total, out);
  return 1;
}

int launch_count_interfacial(const Geo& g, long long fid_begin, long long fid_end, unsigned long long* count,
                             cudaStream_t st) {
  count_interfacial_kernel<<<big_grid(fid_end - fid_begin), BLOCK, 0, st>>>(g, fid_begin, fid_end, count);
  return 1;
}

int launch_dense_interfacial(const Geo& g, int8_t* out_own, cudaStream_t st) {
  dense_interfacial_kernel<<<big_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, out_own);
  return 1;
}

int launch_scatter_to_dense(const Geo& g, const double* arr, double* dense_own, cudaStream_t st) {
  scatter_to_dense_kernel<<<big_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, arr, dense_own);
  return 1;
}

int launch_pull_to_dense(const Geo& g, const double* fin, int l, double* dense_own, cudaStream_t st) {
  pull_to_dense_kernel<<<big_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, fin, l, dense_own);
  return 1;
}

int launch_slice(const Geo& g, const double* mom, int axis, int index, double* out4, cudaStream_t st) {
  const long long n = (long long)(axis == 0 ? g.ly : g.lx) * (axis == 2 ? g.ly : g.nzl);
  slice_kernel<<<big_grid(n), BLOCK, 0, st>>>(g, mom, axis, index, out4, 4);
  return 1;
}

int launch_gather_from_dense(const Geo& g, const double* dense_own, double* arr, cudaStream_t st) {
  gather_from_dense_kernel<<<big_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, dense_own, arr);
  return 1;
}

int launch_build_awords(const Geo& g, long long fid_begin, long long fid_end, uint2* awords, cudaStream_t st) {
  build_awords_kernel<<<big_grid(g.nfa), BLOCK, 0, st>>>(g, fid_begin, fid_end, awords);
  return 1;
}

int launch_scatter3_compact_to_dense_aos(const Geo& g, const uint2* awords, const double* a3, long long a_stride,
                                         double* dense_aos_own, cudaStream_t st) {
  scatter3_compact_aos_kernel<<<big_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, awords, a3, a_stride, dense_aos_own);
  return 1;
}

int launch_scatter3_to_dense_aos(const Geo& g, const double* soa3, double* dense_aos_own, cudaStream_t st) {
  scatter3_aos_kernel<<<big_grid((long long)g.plane * g.nzl), BLOCK, 0, st>>>(g, soa3, dense_aos_own);
  return 1;
}

}  // namespace lbg
