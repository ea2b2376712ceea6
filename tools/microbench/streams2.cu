// streams2.cu -- skeleton search for the step kernels: which launch shape / access width / staging moves a
// 19 -> 19 fp64 structure-of-arrays stream (aligned copy, and D3Q19-shifted "pull") closest to the copy peak?
// Not product code: the yardstick the kernels in laboetie_b200/csrc are tuned against (profiles/streams_r4*.txt).
//
// Variants (template parameters of sk<>):
//   VEC      doubles per thread per array access (1: LDG/STG.64, 2: LDG/STG.128 on the aligned streams)
//   THREADS  CTA size;  MINB  resident CTAs per SM the register allocation aims at
//   PERSIST  persistent grid (148 x resident CTAs, tile-stride loop) or one tile per CTA
//   LDP/STP  cache policy of loads (0 default, 1 .cg, 2 .cs) / stores (0 default, 1 .cs)
//   SHIFT    read array l at i + off[l] (misaligned runs: the pull pattern); loads stay 64-bit, stores VEC wide
// plus a cp.async (LDGSTS) double-buffered pull: the next tile's 19 x TILE doubles land in shared memory while
// the current tile is consumed.
#include <cuda_pipeline.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e = (x);                                                      \
    if (e != cudaSuccess) {                                                   \
      printf("%s: %s\n", #x, cudaGetErrorString(e));                          \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

struct Offs {
  int o[32];
};

template <int P>
__device__ __forceinline__ double ld1(const double* p) {
  if constexpr (P == 1) return __ldcg(p);
  else if constexpr (P == 2) return __ldcs(p);
  else return *p;
}
template <int P>
__device__ __forceinline__ double2 ld2(const double* p) {
  if constexpr (P == 1) return __ldcg(reinterpret_cast<const double2*>(p));
  else if constexpr (P == 2) return __ldcs(reinterpret_cast<const double2*>(p));
  else return *reinterpret_cast<const double2*>(p);
}
template <int P>
__device__ __forceinline__ void st1(double* p, double v) {
  if constexpr (P == 1) __stcs(p, v);
  else *p = v;
}
template <int P>
__device__ __forceinline__ void st2(double* p, double2 v) {
  if constexpr (P == 1) __stcs(reinterpret_cast<double2*>(p), v);
  else *reinterpret_cast<double2*>(p) = v;
}

template <int NR, int NW, int VEC, int THREADS, int MINB, bool PERSIST, int LDP, int STP, bool SHIFT>
__global__ void __launch_bounds__(THREADS, MINB) sk(const double* __restrict__ in, double* __restrict__ out, long long n,
                                                    long long stride, Offs offs) {
  constexpr int TILE = THREADS * VEC;
  const long long step = PERSIST ? (long long)gridDim.x * TILE : n;
  for (long long i = (long long)blockIdx.x * TILE + (long long)threadIdx.x * VEC; i < n; i += step) {
    double v[NR][VEC];
#pragma unroll
    for (int l = 0; l < NR; ++l) {
      if constexpr (SHIFT) {
        long long j = i + offs.o[l];
        if (j < 0) j += n;
        if (j + VEC > n) j -= n;
        if (j < 0) j = 0;
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[l][e] = ld1<LDP>(in + l * stride + j + e);
      } else if constexpr (VEC == 2) {
        const double2 t = ld2<LDP>(in + l * stride + i);
        v[l][0] = t.x;
        v[l][1] = t.y;
      } else {
        v[l][0] = ld1<LDP>(in + l * stride + i);
      }
    }
    if constexpr (NW == NR) {
#pragma unroll
      for (int l = 0; l < NW; ++l) {
        if constexpr (VEC == 2) st2<STP>(out + l * stride + i, make_double2(v[l][0] + v[NR - 1 - l][0], v[l][1] + v[NR - 1 - l][1]));
        else st1<STP>(out + l * stride + i, v[l][0] + v[NR - 1 - l][0]);
      }
    } else {
      double s[NW][VEC];
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int e = 0; e < VEC; ++e) s[w][e] = 0.0;
#pragma unroll
      for (int l = 0; l < NR; ++l)
#pragma unroll
        for (int e = 0; e < VEC; ++e) s[l % NW][e] += v[l][e];
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        if constexpr (VEC == 2) st2<STP>(out + w * stride + i, make_double2(s[w][0], s[w][1]));
        else st1<STP>(out + w * stride + i, s[w][0]);
      }
    }
  }
}

// cp.async (LDGSTS) double-buffered shifted pull, 8-byte copies, one node per thread, 2 stages in shared memory.
template <int NR, int THREADS, int MINB, int STP>
__global__ void __launch_bounds__(THREADS, MINB) sk_async(const double* __restrict__ in, double* __restrict__ out,
                                                          long long n, long long stride, Offs offs) {
  extern __shared__ double sm[];  // [2][NR][THREADS]
  const long long step = (long long)gridDim.x * THREADS;
  long long i = (long long)blockIdx.x * THREADS + threadIdx.x;
  auto issue = [&](long long ii, int stage) {
    if (ii < n) {
#pragma unroll
      for (int l = 0; l < NR; ++l) {
        long long j = ii + offs.o[l];
        if (j < 0) j += n;
        if (j >= n) j -= n;
        __pipeline_memcpy_async(&sm[((size_t)stage * NR + l) * THREADS + threadIdx.x], in + l * stride + j, 8);
      }
    }
    __pipeline_commit();
  };
  int stage = 0;
  issue(i, 0);
  for (; i < n; i += step) {
    issue(i + step, stage ^ 1);
    __pipeline_wait_prior(1);
    double v[NR];
#pragma unroll
    for (int l = 0; l < NR; ++l) v[l] = sm[((size_t)stage * NR + l) * THREADS + threadIdx.x];
#pragma unroll
    for (int l = 0; l < NR; ++l) st1<STP>(out + l * stride + i, v[l] + v[NR - 1 - l]);
    stage ^= 1;
  }
}


// Persistent grid with DYNAMIC tile scheduling: an atomic counter hands out chunks of CHUNK consecutive tiles
// (DYN), or a flat grid where CTA b owns chunk b (DYN == false).  One node per thread, 64-bit accesses.
template <int NR, int THREADS, int MINB, int CHUNK, bool DYN, bool SHIFT>
__global__ void __launch_bounds__(THREADS, MINB) sk_chunk(const double* __restrict__ in, double* __restrict__ out, long long n,
                                                          long long stride, Offs offs, unsigned int* counter) {
  __shared__ unsigned int s_chunk;
  const long long nchunks = (n + (long long)THREADS * CHUNK - 1) / ((long long)THREADS * CHUNK);
  long long chunk = blockIdx.x;
  for (;;) {
    if (DYN) {
      __syncthreads();
      if (threadIdx.x == 0) s_chunk = atomicAdd(counter, 1u);
      __syncthreads();
      chunk = s_chunk;
    }
    if (chunk >= nchunks) break;
    const long long base = chunk * THREADS * CHUNK + threadIdx.x;
#pragma unroll 1
    for (int t = 0; t < CHUNK; ++t) {
      const long long i = base + (long long)t * THREADS;
      if (i >= n) break;
      double v[NR];
#pragma unroll
      for (int l = 0; l < NR; ++l) {
        long long j = i;
        if (SHIFT) {
          j = i + offs.o[l];
          if (j < 0) j += n;
          if (j >= n) j -= n;
        }
        v[l] = __ldcg(in + l * stride + j);
      }
#pragma unroll
      for (int l = 0; l < NR; ++l) out[l * stride + i] = v[l] + v[NR - 1 - l];
    }
    if (!DYN) break;
  }
}

static cudaEvent_t e0, e1;
static unsigned int* g_counter;

template <typename K, typename... A>
void timeit(const char* name, int nr, int nw, long long n, K kern, dim3 grid, int threads, size_t smem, A... args) {
  for (int i = 0; i < 2; ++i) kern<<<grid, threads, smem>>>(args...);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  const int reps = 10;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) kern<<<grid, threads, smem>>>(args...);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double bytes = (double)(nr + nw) * 8.0 * (double)n;
  printf("%-58s grid=%-7u %.3f ms  %.1f GB/s\n", name, grid.x, ms, bytes / ms * 1e-6);
  fflush(stdout);
}

template <int NR, int NW, int VEC, int THREADS, int MINB, bool PERSIST, int LDP, int STP, bool SHIFT>
void run(const double* in, double* out, long long n, long long stride, const Offs& offs) {
  auto k = sk<NR, NW, VEC, THREADS, MINB, PERSIST, LDP, STP, SHIFT>;
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, THREADS, 0));
  constexpr int TILE = THREADS * VEC;
  const long long tiles = (n + TILE - 1) / TILE;
  const unsigned grid = PERSIST ? (unsigned)(148 * per_sm) : (unsigned)tiles;
  char name[128];
  snprintf(name, sizeof name, "%s %d->%d vec%d thr%d minb%d(occ %d) %s ld%d st%d", SHIFT ? "pull" : "copy", NR, NW, VEC, THREADS,
           MINB, per_sm, PERSIST ? "persist" : "flat", LDP, STP);
  timeit(name, NR, NW, n, k, dim3(grid), THREADS, 0, in, out, n, stride, offs);
}

template <int NR, int THREADS, int MINB, int STP>
void run_async(const double* in, double* out, long long n, long long stride, const Offs& offs) {
  auto k = sk_async<NR, THREADS, MINB, STP>;
  const size_t smem = (size_t)2 * NR * THREADS * sizeof(double);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, THREADS, smem));
  char name[128];
  snprintf(name, sizeof name, "pull %d->%d cp.async 2-stage thr%d minb%d(occ %d) st%d", NR, NR, THREADS, MINB, per_sm, STP);
  timeit(name, NR, NR, n, k, dim3(148 * per_sm), THREADS, smem, in, out, n, stride, offs);
}

template <int NR, int THREADS, int MINB, int CHUNK, bool DYN, bool SHIFT>
void run_chunk(const double* in, double* out, long long n, long long stride, const Offs& offs) {
  auto k = sk_chunk<NR, THREADS, MINB, CHUNK, DYN, SHIFT>;
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, THREADS, 0));
  const long long nchunks = (n + (long long)THREADS * CHUNK - 1) / ((long long)THREADS * CHUNK);
  const unsigned grid = DYN ? (unsigned)(148 * per_sm) : (unsigned)nchunks;
  char name[128];
  snprintf(name, sizeof name, "%s %d->%d thr%d minb%d(occ %d) %s chunk %d tiles", SHIFT ? "pull" : "copy", NR, NR, THREADS, MINB,
           per_sm, DYN ? "persist+atomic" : "flat", CHUNK);
  for (int i = 0; i < 2; ++i) {
    CK(cudaMemsetAsync(g_counter, 0, 4));
    k<<<grid, THREADS>>>(in, out, n, stride, offs, g_counter);
  }
  CK(cudaDeviceSynchronize());
  const int reps = 10;
  float tot = 0;
  for (int i = 0; i < reps; ++i) {
    CK(cudaMemsetAsync(g_counter, 0, 4));
    CK(cudaEventRecord(e0));
    k<<<grid, THREADS>>>(in, out, n, stride, offs, g_counter);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    tot += ms;
  }
  const float ms = tot / reps;
  printf("%-58s grid=%-7u %.3f ms  %.1f GB/s\n", name, grid, ms, 2.0 * NR * 8.0 * (double)n / ms * 1e-6);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 80485376LL;  // fluid nodes of the cfg5w slab
  const long long stride = (n + 31) / 32 * 32;
  double *in = nullptr, *out = nullptr;
  CK(cudaMalloc(&in, 25 * stride * sizeof(double)));
  CK(cudaMalloc(&out, 22 * stride * sizeof(double)));
  CK(cudaMemset(in, 0, 25 * stride * sizeof(double)));
  CK(cudaMemset(out, 0, 22 * stride * sizeof(double)));
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int row = 614, plane = 614 * 1024;
  const int cx[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
  const int cy[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
  const int cz[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
  Offs offs{};
  for (int l = 0; l < 19; ++l) offs.o[l] = -(cx[l] + row * cy[l] + plane * cz[l]);
  printf("n = %lld nodes per array (%.2f GB per array)\n", n, n * 8e-9);
  // torch-style reference point: cudaMemcpy device to device over 19 arrays
  {
    CK(cudaMemcpy(out, in, 19 * stride * 8, cudaMemcpyDeviceToDevice));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(out, in, 19 * stride * 8, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-58s %.3f ms  %.1f GB/s\n", "cudaMemcpy D2D 19 arrays", ms / 5, 2.0 * 19 * stride * 8 / (ms / 5) * 1e-6);
  }
  CK(cudaMalloc(&g_counter, 4));
  if (argc > 2 && !strcmp(argv[2], "ncu1")) {
    run<19, 19, 1, 256, 2, true, 1, 0, true>(in, out, n, stride, offs);
    return 0;
  }
  if (argc > 2 && !strcmp(argv[2], "ncu2")) {
    run_chunk<19, 256, 2, 4, true, true>(in, out, n, stride, offs);
    return 0;
  }
  if (argc > 2 && !strcmp(argv[2], "dyn")) {
    run<19, 19, 1, 256, 2, true, 1, 0, false>(in, out, n, stride, offs);
    run<19, 19, 1, 256, 2, false, 1, 0, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 1, true, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 4, true, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 16, true, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 4, 4, true, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 2, false, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 4, false, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 16, false, false>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 64, false, false>(in, out, n, stride, offs);
    run<19, 19, 1, 256, 2, true, 1, 0, true>(in, out, n, stride, offs);
    run<19, 19, 1, 256, 2, false, 1, 0, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 1, true, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 4, true, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 16, true, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 4, 4, true, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 2, false, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 4, false, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 16, false, true>(in, out, n, stride, offs);
    run_chunk<19, 256, 2, 64, false, true>(in, out, n, stride, offs);
    return 0;
  }
  // ---- 1 -> 1 copies: what does the skeleton itself cost?
  run<1, 1, 1, 256, 8, true, 1, 0, false>(in, out, n, stride, offs);
  run<1, 1, 2, 256, 8, true, 1, 0, false>(in, out, n, stride, offs);
  run<1, 1, 2, 256, 8, false, 1, 0, false>(in, out, n, stride, offs);
  run<1, 1, 2, 256, 8, false, 0, 0, false>(in, out, n, stride, offs);
  // ---- 19 -> 19 aligned copies
  run<19, 19, 1, 256, 2, true, 1, 0, false>(in, out, n, stride, offs);   // round-1 skeleton
  run<19, 19, 1, 256, 2, true, 1, 1, false>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 2, true, 2, 1, false>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 2, true, 0, 0, false>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 3, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 4, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 2, false, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 4, false, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 1, 128, 4, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 1, 512, 1, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 1, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 2, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 2, true, 1, 1, false>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 2, true, 2, 1, false>(in, out, n, stride, offs);
  run<19, 19, 2, 128, 2, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 2, 128, 4, true, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 2, false, 1, 0, false>(in, out, n, stride, offs);
  run<19, 19, 2, 128, 4, false, 1, 0, false>(in, out, n, stride, offs);
  // ---- 19 -> 19 shifted pulls (loads 64-bit, stores VEC wide)
  run<19, 19, 1, 256, 2, true, 1, 0, true>(in, out, n, stride, offs);    // round-1 skeleton
  run<19, 19, 1, 256, 2, true, 1, 1, true>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 2, true, 0, 0, true>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 2, true, 2, 1, true>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 3, true, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 4, true, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 2, false, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 1, 256, 4, false, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 1, 128, 4, true, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 1, true, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 2, true, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 2, 128, 2, true, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 2, 128, 4, true, 1, 0, true>(in, out, n, stride, offs);
  run<19, 19, 2, 256, 2, false, 1, 0, true>(in, out, n, stride, offs);
  run_async<19, 256, 2, 0>(in, out, n, stride, offs);
  run_async<19, 256, 2, 1>(in, out, n, stride, offs);
  run_async<19, 128, 4, 0>(in, out, n, stride, offs);
  // ---- LB + check (22 -> 22) and MP-like (25 -> 3, 21 -> 3) shapes with the better skeletons
  run<22, 22, 1, 256, 2, true, 1, 0, false>(in, out, n, stride, offs);
  run<22, 22, 2, 256, 2, true, 1, 0, false>(in, out, n, stride, offs);
  run<25, 3, 1, 256, 2, true, 2, 1, false>(in, out, n, stride, offs);
  run<25, 3, 2, 256, 2, true, 2, 1, false>(in, out, n, stride, offs);
  run<25, 3, 1, 256, 4, true, 2, 1, false>(in, out, n, stride, offs);
  return 0;
}
