// streams.cu -- what HBM throughput can a kernel with the step kernels' access SHAPE reach on this GPU?
// Not product code: a yardstick for DESIGN.md section 6.  Same tiling as the step kernels (persistent
// grid of 296 CTAs x 256 threads, tiles of 256 consecutive nodes, fp64 structure of arrays), no lattice
// logic at all:
//   copy   NR input arrays -> NW output arrays, out[l][i] = in[l][i] (+ in[NR-1-l][i] to keep all inputs alive)
//   shift  the same, but input l is read at i + off[l] (a dense all-fluid D3Q19 pull: misaligned runs)
// Prints achieved GB/s (read + write bytes) for LB-like (19 -> 19), LB+check (22 -> 22) and MP-like (25 -> 3) shapes.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
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

template <int NR, int NW, bool SHIFT>
__global__ void __launch_bounds__(256, 2) streams_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                         long long n, long long stride, Offs offs) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
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
    if (NW == NR) {
#pragma unroll
      for (int l = 0; l < NW; ++l) out[l * stride + i] = v[l] + v[NR - 1 - l];
    } else {
      double s[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) s[w] = 0.0;
#pragma unroll
      for (int l = 0; l < NR; ++l) s[l % NW] += v[l];
#pragma unroll
      for (int w = 0; w < NW; ++w) __stcs(out + w * stride + i, s[w]);
    }
  }
}

template <int NR, int NW, bool SHIFT>
void run(const char* name, const double* in, double* out, long long n, long long stride, const Offs& offs) {
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, streams_kernel<NR, NW, SHIFT>, 256, 0));
  const int grid = 148 * per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) streams_kernel<NR, NW, SHIFT><<<grid, 256>>>(in, out, n, stride, offs);
  CK(cudaDeviceSynchronize());
  const int reps = 20;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) streams_kernel<NR, NW, SHIFT><<<grid, 256>>>(in, out, n, stride, offs);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  const double bytes = (double)(NR + NW) * 8.0 * (double)n;
  printf("%-28s NR=%2d NW=%2d grid=%d  %.3f ms  %.1f GB/s\n", name, NR, NW, grid, ms, bytes / ms * 1e-6);
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 80485376LL;  // fluid nodes of the cfg5w slab
  const long long stride = (n + 31) / 32 * 32;
  double *in = nullptr, *out = nullptr;
  CK(cudaMalloc(&in, 25 * stride * sizeof(double)));
  CK(cudaMalloc(&out, 22 * stride * sizeof(double)));
  CK(cudaMemset(in, 0, 25 * stride * sizeof(double)));
  CK(cudaMemset(out, 0, 22 * stride * sizeof(double)));
  // dense D3Q19 pull offsets on a 1024 x 1024 plane, scaled to the fluid fraction 0.6 (fids per row / plane)
  const int row = 614, plane = 614 * 1024;
  const int cx[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
  const int cy[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
  const int cz[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
  Offs offs{};
  for (int l = 0; l < 19; ++l) offs.o[l] = -(cx[l] + row * cy[l] + plane * cz[l]);
  printf("n = %lld nodes per array (%.2f GB per array)\n", n, n * 8e-9);
  run<1, 1, false>("copy 1 -> 1", in, out, n, stride, offs);
  run<19, 19, false>("copy 19 -> 19 (LB)", in, out, n, stride, offs);
  run<19, 19, true>("pull 19 -> 19 (LB, shifted)", in, out, n, stride, offs);
  run<22, 22, false>("copy 22 -> 22 (LB + check)", in, out, n, stride, offs);
  run<25, 3, false>("read 25 -> 3 (MP)", in, out, n, stride, offs);
  run<19, 3, true>("pull 19 -> 3", in, out, n, stride, offs);
  return 0;
}
