// lb_node.cuh -- per-node arithmetic of Phase A shared by the two-lattice and the in-place (AA) kernels:
// the l-ordered moments of equilibration.f90:254,293-300 and the collision of module_collision.f90:77-108,
// both in the reference's association order (no FMA contraction: this directory builds with -fmad=false).
#pragma once
#include "lattice.cuh"

namespace lbg {
using namespace d3q19;

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

}  // namespace lbg
