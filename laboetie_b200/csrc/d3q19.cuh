// d3q19.cuh -- compile-time D3Q19 table and exactly-ordered arithmetic helpers.
//
// Velocity order, weights and inverse pairs are those of the reference
// (module_lbmodel.f90:66-86,122-162); index L here is the reference's l-1.
//
// Arithmetic contract: every per-node expression is evaluated in
// the reference's association order with no FMA contraction (this directory is
// compiled with -fmad=false).  The helpers below drop terms whose coefficient
// is a structural zero (x + 0*y == x) and replace multiplications by +-1 with
// sign changes; both are value-exact in IEEE-754, so results match the
// reference's literal loops bit for bit (up to the sign of zero).
#pragma once
#include <cstdint>

namespace d3q19 {

constexpr int NV = 19;

__host__ __device__ constexpr int cx(int l) {
  constexpr int t[NV] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
  return t[l];
}
__host__ __device__ constexpr int cy(int l) {
  constexpr int t[NV] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
  return t[l];
}
__host__ __device__ constexpr int cz(int l) {
  constexpr int t[NV] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
  return t[l];
}
__host__ __device__ constexpr int inv(int l) {
  constexpr int t[NV] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
  return t[l];
}
// 0: rest, 1: axis (w=1/18), 2: diagonal (w=1/36)
__host__ __device__ constexpr int kind(int l) { return l == 0 ? 0 : (l <= 6 ? 1 : 2); }

// Per-kind constants, filled on the host exactly as module_lbmodel.f90:122-136
// evaluates them (each operation rounded to fp64).
struct Consts {
  double a0[3], a1[3], a2[3], two_a2[3];
  double csq;      // 1/3
  double c1;       // 1 - csq  == cx**2 - csq for cx = +-1
  double mcsq;     // 0 - csq
};

// (cx*x + cy*y) + cz*z, zero terms dropped, +-1 applied as a sign.
template <int L>
__device__ __forceinline__ double cdot(double x, double y, double z) {
  constexpr int X = cx(L), Y = cy(L), Z = cz(L);
  if constexpr (X == 0 && Y == 0 && Z == 0) {
    return 0.0;
  } else if constexpr (X != 0 && Y == 0 && Z == 0) {
    return X > 0 ? x : -x;
  } else if constexpr (X == 0 && Y != 0 && Z == 0) {
    return Y > 0 ? y : -y;
  } else if constexpr (X == 0 && Y == 0 && Z != 0) {
    return Z > 0 ? z : -z;
  } else if constexpr (X != 0 && Y != 0) {
    return (X > 0 ? x : -x) + (Y > 0 ? y : -y);
  } else if constexpr (X != 0 && Z != 0) {
    return (X > 0 ? x : -x) + (Z > 0 ? z : -z);
  } else {
    return (Y > 0 ? y : -y) + (Z > 0 ? z : -z);
  }
}

}  // namespace d3q19
