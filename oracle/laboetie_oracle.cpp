// laboetie_oracle.cpp -- CPU oracle for the laboetie time-stepping hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under laboetie_b200/ may include, link,
// import or execute this file.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker or
// as the timed CPU baseline -- never as the product.
//
// PARITY UNPINNED: the reference (maxlevesque/laboetie, pure Fortran 2008)
// ships no tests, no golden vectors and no expected outputs, and there is no
// Fortran compiler in this image, so the reference binary cannot be run to
// produce any.  This file is a statement-by-statement C++ restatement of the
// reference routines listed below; its fidelity rests on (1) inspection
// against the cited lines, (2) an independently written numpy restatement
// (oracle/numpy_restatement.py) that must agree with it bit for bit, and
// (3) the analytic invariants in tests/test_oracle.py.
//
// Restated routines (file:line relative to /root/reference):
//   src/module_lbmodel.f90:66-86,122-162     D3Q19 velocities, weights, inverse
//   src/module_system.f90:99-112             pbc
//   src/init_simu.f90:24-39                  initial populations
//   src/module_geometry.f90:158-166,206-277  slit, BCC spheres, cylinder
//   src/module_geometry.f90:380-426,430-511  geom.in and geom.pbm readers
//   src/supercell_definition.f90:115-147     detectInterfacialNodes
//   src/module_collision.f90:15-130          collide (default 2nd-order branch)
//   src/equilibration.f90:204-300,339-387    bounce-back, streaming, moments,
//                                            convergence state machine
//   src/drop_tracers.f90:20-55,63-105        tracer phase driver, population rebuild
//   src/module_moment_propagation.f90:30-160,164-289,331-346  init, propagate
//
// Array layouts follow the reference: Phase-A populations n(i,j,k,l) with i
// fastest and l slowest; Phase-B populations n(l,i,j,k) with l fastest;
// Propagated_Quantity(x:z,i,j,k,now:next).  All indices below are 0-based.
//
// Build flags mirror Makefile:13 (-O3 -funroll-loops -fopenmp) plus
// -ffp-contract=off: the reference Makefile targets baseline x86-64, which has
// no FMA, so no contraction happens there either.
//
// The two OpenMP regions are the reference's own: streaming over the velocity
// index (equilibration.f90:227-243) and moment propagation over z-slices
// (module_moment_propagation.f90:201-207).  Everything else is serial there
// (`where` / `do concurrent` run serially under gfortran) and is serial here.

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

constexpr int NV = 19;
constexpr int8_t FLUID = 0, SOLID = 1;  // module_system.f90:32
const double EPS = std::numeric_limits<double>::epsilon();

// module_lbmodel.f90:66-86 (index 0 here is l=1 there)
const int C[NV][3] = {
    {0, 0, 0},  {1, 0, 0},  {-1, 0, 0}, {0, 1, 0},  {0, -1, 0}, {0, 0, 1},  {0, 0, -1},
    {1, 1, 0},  {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1},
    {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}};

struct Table {
  double a0[NV], a1[NV], a2[NV];
  int inv[NV];
  Table() {
    // module_lbmodel.f90:122-145
    volatile double one = 1.0, three = 3.0, eighteen = 18.0, thirtysix = 36.0;
    const double csq = one / three;
    const double a_00 = one / three, a_01 = one / eighteen, a_02 = one / thirtysix;
    const double a_10 = a_00 / csq, a_11 = a_01 / csq, a_12 = a_02 / csq;
    const double csq2 = csq * csq;
    const double a_20 = a_00 / (2 * csq2), a_21 = a_01 / (2 * csq2), a_22 = a_02 / (2 * csq2);
    for (int l = 0; l < NV; ++l) {
      const int kind = (l == 0) ? 0 : (l <= 6 ? 1 : 2);
      a0[l] = kind == 0 ? a_00 : (kind == 1 ? a_01 : a_02);
      a1[l] = kind == 0 ? a_10 : (kind == 1 ? a_11 : a_12);
      a2[l] = kind == 0 ? a_20 : (kind == 1 ? a_21 : a_22);
    }
    // module_lbmodel.f90:156-162
    for (int l = 0; l < NV; ++l)
      for (int li = 0; li < NV; ++li)
        if (C[l][0] == -C[li][0] && C[l][1] == -C[li][1] && C[l][2] == -C[li][2]) inv[l] = li;
  }
};
const Table T;

// NORM2 as libgfortran evaluates it (gcc 13 libgfortran/generated/norm2_r8.c):
// a running scale with rescaled sum of squares, not sqrt(sum x^2).  The
// reference calls NORM2 for the cylinder and BCC thresholds
// (module_geometry.f90:222-233,269), where exact ties (Pythagorean pairs at
// distance == radius) make the last bit matter.
inline double norm2_gfortran(const double* x, int len) {
  double result = 0.0, scale = 1.0;
  for (int n = 0; n < len; ++n) {
    if (x[n] != 0) {
      const double absX = std::fabs(x[n]);
      if (scale < absX) {
        const double val = scale / absX;
        result = 1 + result * val * val;
        scale = absX;
      } else {
        const double val = absX / scale;
        result += val * val;
      }
    }
  }
  return scale * std::sqrt(result);
}

// module_system.f90:99-112, 1-based like the reference
inline int pbc(int i, int imax) {
  if (i == 0) return imax;
  if (i == imax + 1) return 1;
  return i;
}

struct Dim {
  int lx, ly, lz;
  size_t N() const { return (size_t)lx * ly * lz; }
  // 1-based (i,j,k) -> 0-based linear, i fastest
  size_t at(int i, int j, int k) const { return (size_t)(i - 1) + (size_t)lx * ((size_t)(j - 1) + (size_t)ly * (size_t)(k - 1)); }
};

// equilibration.f90:109-119
struct NeighbourTables {
  std::vector<int> il, jl, kl;  // [l][i], 1-based values
  NeighbourTables(const Dim& d) : il((size_t)NV * d.lx), jl((size_t)NV * d.ly), kl((size_t)NV * d.lz) {
    for (int l = 0; l < NV; ++l) {
      for (int i = 1; i <= d.lx; ++i) il[(size_t)l * d.lx + i - 1] = pbc(i + C[l][0], d.lx);
      for (int j = 1; j <= d.ly; ++j) jl[(size_t)l * d.ly + j - 1] = pbc(j + C[l][1], d.ly);
      for (int k = 1; k <= d.lz; ++k) kl[(size_t)l * d.lz + k - 1] = pbc(k + C[l][2], d.lz);
    }
  }
};

}  // namespace

extern "C" {

int orc_nvel() { return NV; }

// module_lbmodel.f90:66-86,122-162
void orc_lbm_table(int* c, double* a0, double* a1, double* a2, int* inv) {
  for (int l = 0; l < NV; ++l) {
    for (int d = 0; d < 3; ++d) c[l * 3 + d] = C[l][d];
    a0[l] = T.a0[l];
    a1[l] = T.a1[l];
    a2[l] = T.a2[l];
    inv[l] = T.inv[l];
  }
}

// supercell_definition.f90:50-59 with module_geometry.f90:158-166 (slit),
// :253-277 (cylinder), :206-245 (BCC "cc").  Returns 0, or <0 for the
// reference's `stop` conditions.
int orc_geometry(int label, int lx, int ly, int lz, int8_t* nature) {
  const Dim d{lx, ly, lz};
  std::memset(nature, FLUID, d.N());
  switch (label) {
    case -1:
      return 0;
    case 1:  // construct_slit
      for (int j = 1; j <= ly; ++j)
        for (int i = 1; i <= lx; ++i) {
          nature[d.at(i, j, 1)] = SOLID;
          nature[d.at(i, j, lz)] = SOLID;
        }
      return 0;
    case 2: {  // construct_cylinder
      if (lx != ly) return -1;
      if (lx < 3) return -2;
      const double ox = (double)(lx + 1) / 2.0, oy = (double)(ly + 1) / 2.0;
      const double radius = (double)(lx - 1) / 2.0;
      for (int i = 1; i <= lx; ++i)
        for (int j = 1; j <= ly; ++j) {
          const double rn[2] = {(double)i - ox, (double)j - oy};
          const double nrm = norm2_gfortran(rn, 2);
          const int8_t v = (nrm >= radius) ? SOLID : FLUID;
          for (int k = 1; k <= lz; ++k) nature[d.at(i, j, k)] = v;
        }
      return 0;
    }
    case 3: {  // construct_cc
      if (lx != ly || lx != lz) return -1;
      const double thr = (double)(lx - 1) * std::sqrt(3.0) / 4.0;
      // corners are `real([..])`, i.e. default (single) real, exact for these integers;
      // the centre is real([lx+1,..])/2._dp
      const double corners[9][3] = {
          {1, 1, 1},
          {(double)lx, 1, 1},
          {1, (double)ly, 1},
          {1, 1, (double)lz},
          {(double)lx, (double)ly, 1},
          {(double)lx, 1, (double)lz},
          {1, (double)ly, (double)lz},
          {(double)lx, (double)ly, (double)lz},
          {(double)(lx + 1) / 2.0, (double)(ly + 1) / 2.0, (double)(lz + 1) / 2.0}};
      for (int k = 1; k <= lz; ++k)
        for (int j = 1; j <= ly; ++j)
          for (int i = 1; i <= lx; ++i) {
            bool in = false;
            for (int s = 0; s < 9 && !in; ++s) {
              const double rn[3] = {i - corners[s][0], j - corners[s][1], k - corners[s][2]};
              if (norm2_gfortran(rn, 3) <= thr) in = true;
            }
            nature[d.at(i, j, k)] = in ? SOLID : FLUID;
          }
      return 0;
    }
    default:
      return -100;  // other labels are outside the scope table
  }
}

// module_geometry.f90:380-426.  One "i j k" line per solid node.  The
// reference's list-directed read leaves (i,j,k) unchanged at end of file, so
// the last line is applied twice there, which is idempotent.
int orc_read_geom_in(const char* path, int lx, int ly, int lz, int8_t* nature) {
  const Dim d{lx, ly, lz};
  std::memset(nature, FLUID, d.N());
  FILE* f = std::fopen(path, "r");
  if (!f) return -1;
  int i, j, k, rc = 0;
  while (std::fscanf(f, "%d %d %d", &i, &j, &k) == 3) {
    if (i <= 0 || j <= 0 || k <= 0 || i > lx || j > ly || k > lz) {
      rc = -2;
      break;
    }
    nature[d.at(i, j, k)] = SOLID;
  }
  std::fclose(f);
  return rc;
}

// module_geometry.f90:430-511.  P1 bitmap, ncolumn == ly, nline == lz, lx == 1.
int orc_read_pbm(const char* path, int lx, int ly, int lz, int8_t* nature) {
  const Dim d{lx, ly, lz};
  if (lx != 1) return -3;
  std::memset(nature, FLUID, d.N());
  FILE* f = std::fopen(path, "r");
  if (!f) return -1;
  char magic[8] = {0};
  int ncol = 0, nline = 0, rc = 0;
  if (std::fscanf(f, "%7s", magic) != 1 || std::strcmp(magic, "P1") != 0) rc = -4;
  if (!rc && std::fscanf(f, "%d %d", &ncol, &nline) != 2) rc = -4;
  if (!rc && (ncol != ly || nline != lz)) rc = -5;
  for (int j = 1; !rc && j <= nline; ++j)
    for (int i = 1; !rc && i <= ncol; ++i) {
      int ch;
      do ch = std::fgetc(f);
      while (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t');
      if (ch == '1') nature[d.at(1, i, j)] = SOLID;
      else if (ch != '0') rc = -6;
    }
  std::fclose(f);
  return rc;
}

// supercell_definition.f90:115-147
void orc_detect_interfacial(int lx, int ly, int lz, const int8_t* nature, int8_t* interfacial) {
  const Dim d{lx, ly, lz};
  std::memset(interfacial, 0, d.N());
  for (int i = 1; i <= lx; ++i)
    for (int j = 1; j <= ly; ++j)
      for (int k = 1; k <= lz; ++k)
        for (int l = 1; l < NV; ++l) {
          const int in = pbc(i + C[l][0], lx), jn = pbc(j + C[l][1], ly), kn = pbc(k + C[l][2], lz);
          if (nature[d.at(i, j, k)] != nature[d.at(in, jn, kn)]) {
            interfacial[d.at(i, j, k)] = 1;
            break;
          }
        }
}

// init_simu.f90:24-39
void orc_init_populations(int lx, int ly, int lz, const int8_t* nature, double rho0, double* n, double* density) {
  const size_t N = Dim{lx, ly, lz}.N();
  for (size_t r = 0; r < N; ++r) density[r] = (nature[r] != SOLID) ? rho0 : 0.0;
  for (int l = 0; l < NV; ++l)
    for (size_t r = 0; r < N; ++r) n[(size_t)l * N + r] = density[r] * T.a0[l];
}

// module_collision.f90:15-130, default branch (:69-108).  Returns -1 for the
// relaxation_time guard (:39) .
int orc_collide(int lx, int ly, int lz, const int8_t* nature, double tau, double* n, const double* density,
                const double* jx, const double* jy, const double* jz, const double* fx, const double* fy,
                const double* fz) {
  if (tau < 0.5) return -1;
  const size_t N = Dim{lx, ly, lz}.N();
  const double csq = 1.0 / 3.0;
  // the scratch arrays are allocated and zero-filled on every call (:67-75)
  std::vector<double> neq(N, 0.0), ux(N, 0.0), uy(N, 0.0), uz(N, 0.0);
  for (size_t r = 0; r < N; ++r) {  // :77-85
    if (nature[r] == FLUID) {
      ux[r] = jx[r] / density[r];
      uy[r] = jy[r] / density[r];
      uz[r] = jz[r] / density[r];
    } else {
      ux[r] = 0;
      uy[r] = 0;
      uz[r] = 0;
    }
  }
  for (int l = 0; l < NV; ++l) {  // :87-108
    const double a0 = T.a0[l], a1 = T.a1[l], a2 = T.a2[l];
    const double cx = C[l][0], cy = C[l][1], cz = C[l][2];
    double* nl = n + (size_t)l * N;
    for (size_t r = 0; r < N; ++r) {
      if (nature[r] != FLUID) continue;
      neq[r] = a0 * density[r] + a1 * (cx * jx[r] + cy * jy[r] + cz * jz[r]) +
               a2 * (jx[r] * ux[r] * (cx * cx - csq) + jx[r] * uy[r] * cx * cy + jx[r] * uz[r] * cx * cz +
                     jy[r] * ux[r] * cy * cx + jy[r] * uy[r] * (cy * cy - csq) + jy[r] * uz[r] * cy * cz +
                     jz[r] * ux[r] * cz * cx + jz[r] * uy[r] * cz * cy + jz[r] * uz[r] * (cz * cz - csq));
    }
    for (size_t r = 0; r < N; ++r) {
      if (nature[r] != FLUID) continue;
      nl[r] = (1.0 - 1.0 / tau) * nl[r] + (1.0 / tau) * neq[r] +
              (1.0 - 1.0 / (2.0 * tau)) *
                  (a1 * ((cx - ux[r]) * fx[r] + (cy - uy[r]) * fy[r] + (cz - uz[r]) * fz[r]) +
                   2.0 * a2 * (cx * ux[r] + cy * uy[r] + cz * uz[r]) * (cx * fx[r] + cy * fy[r] + cz * fz[r]));
    }
  }
  return 0;
}

// equilibration.f90:204-222
void orc_bounce_back(int lx, int ly, int lz, const int8_t* nature, double* n) {
  const Dim d{lx, ly, lz};
  const size_t N = d.N();
  const NeighbourTables nb(d);
  for (int l = 0; l < NV; l += 2)  // l = lmin, lmin+2, ... (1,3,..,19 there)
    for (int k = 1; k <= lz; ++k) {
      const int kp = nb.kl[(size_t)l * lz + k - 1];
      for (int j = 1; j <= ly; ++j) {
        const int jp = nb.jl[(size_t)l * ly + j - 1];
        for (int i = 1; i <= lx; ++i) {
          const int ip = nb.il[(size_t)l * lx + i - 1];
          const size_t r = d.at(i, j, k), rp = d.at(ip, jp, kp);
          if (nature[r] != nature[rp]) {
            const double n_loc = n[(size_t)l * N + r];
            n[(size_t)l * N + r] = n[(size_t)T.inv[l] * N + rp];
            n[(size_t)T.inv[l] * N + rp] = n_loc;
          }
        }
      }
    }
}

// equilibration.f90:227-243 (OpenMP region P1, over the velocity index)
void orc_stream(int lx, int ly, int lz, double* n) {
  const Dim d{lx, ly, lz};
  const size_t N = d.N();
  const NeighbourTables nb(d);
#pragma omp parallel for schedule(static)
  for (int l = 0; l < NV; ++l) {
    double* nl = n + (size_t)l * N;
    std::vector<double> n_old(nl, nl + N);  // private copy of n(:,:,:,l)
    for (int k = 1; k <= lz; ++k) {
      const int kp = nb.kl[(size_t)l * lz + k - 1];
      for (int j = 1; j <= ly; ++j) {
        const int jp = nb.jl[(size_t)l * ly + j - 1];
        for (int i = 1; i <= lx; ++i) {
          const int ip = nb.il[(size_t)l * lx + i - 1];
          nl[d.at(ip, jp, kp)] = n_old[d.at(i, j, k)];
        }
      }
    }
  }
}

// equilibration.f90:248-300,339-343.  Returns 1 if ANY(n<0) (the ERROR STOP at :248).
// On return jx/jy/jz hold the new momentum density, j*_old the previous one.
int orc_moments(int lx, int ly, int lz, const double* n, double* density, double* jx, double* jy, double* jz,
                double* jx_old, double* jy_old, double* jz_old, const double* fx, const double* fy,
                const double* fz, double* l2err) {
  const size_t N = Dim{lx, ly, lz}.N();
  int negative = 0;
  for (size_t q = 0; q < (size_t)NV * N; ++q)
    if (n[q] < 0) {
      negative = 1;
      break;
    }
  for (size_t r = 0; r < N; ++r) density[r] = 0.0;  // density = SUM(n,4)
  for (int l = 0; l < NV; ++l)
    for (size_t r = 0; r < N; ++r) density[r] += n[(size_t)l * N + r];
  std::memcpy(jx_old, jx, N * sizeof(double));
  std::memcpy(jy_old, jy, N * sizeof(double));
  std::memcpy(jz_old, jz, N * sizeof(double));
  for (size_t r = 0; r < N; ++r) {
    jx[r] = fx[r] / 2.0;
    jy[r] = fy[r] / 2.0;
    jz[r] = fz[r] / 2.0;
  }
  for (int l = 0; l < NV; ++l) {
    const double cx = C[l][0], cy = C[l][1], cz = C[l][2];
    const double* nl = n + (size_t)l * N;
    for (size_t r = 0; r < N; ++r) jx[r] = jx[r] + nl[r] * cx;
    for (size_t r = 0; r < N; ++r) jy[r] = jy[r] + nl[r] * cy;
    for (size_t r = 0; r < N; ++r) jz[r] = jz[r] + nl[r] * cz;
  }
  double ex = 0, ey = 0, ez = 0;  // maxval(abs(..)) >= 0
  for (size_t r = 0; r < N; ++r) {
    ex = std::fmax(ex, std::fabs(jx[r] - jx_old[r]));
    ey = std::fmax(ey, std::fabs(jy[r] - jy_old[r]));
    ez = std::fmax(ez, std::fabs(jz[r] - jz_old[r]));
  }
  *l2err = std::fmax(ex, std::fmax(ey, ez));
  return negative;
}

// One body of the `do t` loop, equilibration.f90:194-343 (without the I/O).
int orc_lb_step(int lx, int ly, int lz, const int8_t* nature, double tau, double* n, double* density, double* jx,
                double* jy, double* jz, double* jx_old, double* jy_old, double* jz_old, const double* fx,
                const double* fy, const double* fz, double* l2err) {
  if (orc_collide(lx, ly, lz, nature, tau, n, density, jx, jy, jz, fx, fy, fz)) return -1;
  orc_bounce_back(lx, ly, lz, nature, n);
  orc_stream(lx, ly, lz, n);
  return orc_moments(lx, ly, lz, n, density, jx, jy, jz, jx_old, jy_old, jz_old, fx, fy, fz, l2err);
}

// equilibration.f90:59-119 (setup), :143-491 (time loop with the convergence
// state machine, uniform-force branch :381-386), :551-555 (write-back).
// n, density must come from orc_init_populations.  l2err_hist[t-1] receives
// the l2err of step t for t <= hist_cap.  Returns 0 on the two-stage exit,
// 1 for negative populations, 2 if max_steps was reached first.
int orc_equilibration(int lx, int ly, int lz, const int8_t* nature, double tau, double target_error,
                      const double* f_ext, int max_steps, double* n, double* density, double* jx, double* jy,
                      double* jz, double* l2err_hist, int hist_cap, int* t_exit, int* t_fext) {
  const size_t N = Dim{lx, ly, lz}.N();
  std::vector<double> jxo(N, 0.0), jyo(N, 0.0), jzo(N, 0.0), fx(N, 0.0), fy(N, 0.0), fz(N, 0.0);
  for (size_t r = 0; r < N; ++r) jx[r] = jy[r] = jz[r] = 0.0;
  bool without = false, with = false;
  *t_fext = 0;
  *t_exit = 0;
  for (int t = 1; t <= max_steps; ++t) {
    double l2err;
    const int rc = orc_lb_step(lx, ly, lz, nature, tau, n, density, jx, jy, jz, jxo.data(), jyo.data(), jzo.data(),
                               fx.data(), fy.data(), fz.data(), &l2err);
    *t_exit = t;
    if (rc) return rc < 0 ? -1 : 1;
    if (t <= hist_cap) l2err_hist[t - 1] = l2err;
    const bool converged = (l2err <= target_error && t > 2);  // :346-350
    if (converged) {
      if (!without) without = true;
      else with = true;
      if (without && with && t > 2) return 0;  // :373-374
      if (without && !with) {  // :377-386
        *t_fext = t + 1;
        for (size_t r = 0; r < N; ++r)
          if (nature[r] == FLUID) {
            fx[r] = f_ext[0];
            fy[r] = f_ext[1];
            fz[r] = f_ext[2];
          }
      }
    }
  }
  return 2;
}

// equilibration.f90:388-487: the `compensate_f_ext` force field ("Dominika particle").  A particle of
// odd diameter pd centred on (px,py,pz) (1-based) carries the force f_ext, spread evenly over its l nodes;
// in the bulk cell (geometryLabel == -1) a uniform background -f_ext/fluid_nodes compensates it.
// Returns 0, or <0 for the reference's stops (-1 even diameter, -2 even lattice extent, -3 particle on a
// solid node, -4 compensation check failed).  *nodes_in_particle receives l.
int orc_compensate_force(int lx, int ly, int lz, const int8_t* nature, const double* f_ext, int pd, int px, int py,
                         int pz, int geometry_label, double* fx, double* fy, double* fz, int* nodes_in_particle) {
  const Dim d{lx, ly, lz};
  const size_t N = d.N();
  if (pd % 2 == 0) return -1;                                 // :391-395
  if (lx % 2 == 0 || ly % 2 == 0 || lz % 2 == 0) return -2;   // :397-402
  const int pdr = pd / 2;                                     // :403
  long fluid_nodes = 0;
  for (size_t r = 0; r < N; ++r) fluid_nodes += (nature[r] == FLUID);
  for (size_t r = 0; r < N; ++r) fx[r] = fy[r] = fz[r] = 0.0;  // :405-407
  int l = 0;
  bool err = false;
  for (int i = px - pdr; i <= px + pdr; ++i)                   // :419-432
    for (int j = py - pdr; j <= py + pdr; ++j)
      for (int k = pz - pdr; k <= pz + pdr; ++k) {
        const double v[3] = {(double)(i - px), (double)(j - py), (double)(k - pz)};
        if (norm2_gfortran(v, 3) > (double)pd / 2.0) continue;
        const size_t r = d.at(i, j, k);
        if (nature[r] != FLUID) err = true;
        fx[r] = f_ext[0];
        fy[r] = f_ext[1];
        fz[r] = f_ext[2];
        ++l;
      }
  *nodes_in_particle = l;
  if (err) return -3;                                          // :435-438
  if (geometry_label == -1) {                                  // :449-466
    for (size_t r = 0; r < N; ++r) {
      if (fx[r] == f_ext[0] && fy[r] == f_ext[1] && fz[r] == f_ext[2]) {
        fx[r] = -f_ext[0] / (double)fluid_nodes + fx[r] / (double)l;
        fy[r] = -f_ext[1] / (double)fluid_nodes + fy[r] / (double)l;
        fz[r] = -f_ext[2] / (double)fluid_nodes + fz[r] / (double)l;
      } else {
        fx[r] = -f_ext[0] / (double)fluid_nodes;
        fy[r] = -f_ext[1] / (double)fluid_nodes;
        fz[r] = -f_ext[2] / (double)fluid_nodes;
      }
    }
    double sx = 0, sy = 0, sz = 0;
    for (size_t r = 0; r < N; ++r) {
      sx += fx[r];
      sy += fy[r];
      sz += fz[r];
    }
    if (std::fabs(sx / fluid_nodes) > EPS || std::fabs(sy / fluid_nodes) > EPS || std::fabs(sz / fluid_nodes) > EPS) return -4;
  } else {                                                     // :467-477
    for (size_t r = 0; r < N; ++r) {
      if (fx[r] == f_ext[0] && fy[r] == f_ext[1] && fz[r] == f_ext[2]) {
        fx[r] = fx[r] / (double)l;
        fy[r] = fy[r] / (double)l;
        fz[r] = fz[r] / (double)l;
      } else {
        fx[r] = fy[r] = fz[r] = 0.0;
      }
    }
  }
  for (size_t r = 0; r < N; ++r)                               // :480-484
    if (nature[r] != FLUID) fx[r] = fy[r] = fz[r] = 0.0;
  return 0;
}

// equilibration.f90:161-172,505-516: one row of 4 per index along `axis`
// (0=x,1=y,2=z): SUM(jx), SUM(jy), SUM(jz), SUM(density)/MAX(COUNT(density>eps),1).
void orc_profiles(int lx, int ly, int lz, const double* density, const double* jx, const double* jy, const double* jz,
                  int axis, double* out) {
  const Dim d{lx, ly, lz};
  const int len = axis == 0 ? lx : (axis == 1 ? ly : lz);
  for (int p = 1; p <= len; ++p) {
    double sx = 0, sy = 0, sz = 0, sd = 0;
    long cnt = 0;
    // Fortran array-section order: first remaining index fastest
    const int n1 = axis == 0 ? ly : lx, n2 = axis == 2 ? ly : lz;
    for (int b = 1; b <= n2; ++b)
      for (int a = 1; a <= n1; ++a) {
        const int i = axis == 0 ? p : a;
        const int j = axis == 0 ? a : (axis == 1 ? p : b);
        const int k = axis == 2 ? p : b;
        const size_t r = d.at(i, j, k);
        sx += jx[r];
        sy += jy[r];
        sz += jz[r];
        sd += density[r];
        if (density[r] > EPS) ++cnt;
      }
    out[(p - 1) * 4 + 0] = sx;
    out[(p - 1) * 4 + 1] = sy;
    out[(p - 1) * 4 + 2] = sz;
    out[(p - 1) * 4 + 3] = sd / (double)(cnt > 1 ? cnt : 1);
  }
}

// equilibration.f90:260
void orc_total_flux(int lx, int ly, int lz, const double* jx, const double* jy, const double* jz, double* out) {
  const size_t N = Dim{lx, ly, lz}.N();
  double sx = 0, sy = 0, sz = 0;
  for (size_t r = 0; r < N; ++r) sx += jx[r];
  for (size_t r = 0; r < N; ++r) sy += jy[r];
  for (size_t r = 0; r < N; ++r) sz += jz[r];
  out[0] = sx;
  out[1] = sy;
  out[2] = sz;
}

// drop_tracers.f90:63-105.  ntr is n(l,i,j,k): l fastest.  tracer charge q=0,
// elec_slope=0 (module storage, never set for neutral runs), so the last term
// evaluates to -(rho*0*D*0) = -0.
int orc_update_tracer_population(int lx, int ly, int lz, const int8_t* nature, const double* density,
                                 const double* jx, const double* jy, const double* jz, const double* f_ext,
                                 double tracer_Db, double* ntr) {
  if (tracer_Db <= EPS) return -1;  // :89
  const size_t N = Dim{lx, ly, lz}.N();
  const double q = 0.0, D = tracer_Db, slope[3] = {0.0, 0.0, 0.0};
  for (size_t r = 0; r < N; ++r) {
    const double jr[3] = {jx[r], jy[r], jz[r]};
    for (int l = 0; l < NV; ++l) {
      double s = 0.0;  // sum(...) of a 3-element array expression
      if (nature[r] == FLUID) {
        for (int dd = 0; dd < 3; ++dd) s += (double)C[l][dd] * (jr[dd] + f_ext[dd] - density[r] * q * D * slope[dd]);
      } else {
        for (int dd = 0; dd < 3; ++dd) s += (double)C[l][dd] * jr[dd];
      }
      ntr[r * NV + l] = T.a0[l] * density[r] + T.a1[l] * s;
    }
  }
  return 0;
}

// module_moment_propagation.f90:341-346
static inline double calc_scattprop(double n, double rho, double w, double lambda, double fermi) {
  return n / rho - w + lambda * w * fermi;
}

// module_moment_propagation.f90:30-160.  P and Pads are
// Propagated_Quantity(x:z,i,j,k,now:next): [t][k][j][i][d], zeroed here.
// vacf0 receives vacf(:,tini).  Returns <0 for the reference's stops.
int orc_mp_init(int lx, int ly, int lz, const int8_t* nature, const int8_t* interfacial, const double* ntr,
                const double* density, double tracer_Db, double tracer_ka, double tracer_kd, double* P,
                double* Pads, double* vacf0, int* consider_adsorption) {
  const Dim d{lx, ly, lz};
  const size_t N = d.N();
  if (tracer_ka < -EPS) return -1;
  if (tracer_kd < -EPS) return -2;
  const double K = (std::fabs(tracer_kd) <= EPS) ? 0.0 : tracer_ka / tracer_kd;  // :46-50
  *consider_adsorption = std::fabs(K) > EPS;                                      // :52-56
  const double kBT = 1.0 / 3.0;
  const double lambda = 4.0 * tracer_Db / kBT;  // :68,336
  std::memset(P, 0, 2 * N * 3 * sizeof(double));
  if (Pads) std::memset(Pads, 0, 2 * N * 3 * sizeof(double));
  long nf = 0, nif = 0;
  for (size_t r = 0; r < N; ++r) {
    if (nature[r] == FLUID) {
      ++nf;
      if (interfacial[r]) ++nif;
    }
  }
  const double Pstat = (double)nf + K * (double)nif;  // :97
  vacf0[0] = vacf0[1] = vacf0[2] = 0.0;
  double* Pnow = P;  // index tini+1 == now
  // gfortran lowers DO CONCURRENT(i,j,k) to a nest with the first index innermost
  for (int k = 1; k <= lz; ++k)
    for (int j = 1; j <= ly; ++j)
      for (int i = 1; i <= lx; ++i) {
        const size_t r = d.at(i, j, k);
        if (nature[r] != FLUID) continue;
        const double bw = 1.0 / Pstat;  // :108
        const double* n_loc = ntr + r * NV;
        const double rho = density[r];
        for (int l = 1; l < NV; ++l) {  // :119-136
          const int ip = pbc(i + C[l][0], lx), jp = pbc(j + C[l][1], ly), kp = pbc(k + C[l][2], lz);
          const size_t rp = d.at(ip, jp, kp);
          if (nature[rp] == SOLID) continue;
          const double exp_dphi = 1.0;
          const double exp_min_dphi = 1.0 / exp_dphi;
          const double fermi = 1.0 / (1.0 + exp_dphi);
          const double sp = calc_scattprop(n_loc[l], rho, T.a0[l], lambda, fermi);
          for (int dd = 0; dd < 3; ++dd) vacf0[dd] = vacf0[dd] + bw * sp * (double)(C[l][dd] * C[l][dd]);
          const int li = T.inv[l];
          const double spp = calc_scattprop(ntr[rp * NV + li], density[rp], T.a0[li], lambda, 1.0 - fermi);
          for (int dd = 0; dd < 3; ++dd)
            Pnow[r * 3 + dd] = Pnow[r * 3 + dd] + exp_min_dphi * spp * (double)C[li][dd] * bw;
        }
      }
  return 0;
}

// module_moment_propagation.f90:164-289, one call.  P/Pads as in orc_mp_init
// ([0]=now, [1]=next).  vacf_out receives the vacf(:,now) written to vacf.dat.
// Returns 1 for 'somewhere restpart is negative' (:257), else 0.
int orc_mp_propagate(int lx, int ly, int lz, const int8_t* nature, const int8_t* interfacial, const double* ntr,
                     const double* density, double tracer_Db, double tracer_ka, double tracer_kd,
                     int consider_adsorption, int it, double* P, double* Pads, double* vacf_out,
                     int* is_converged) {
  const Dim d{lx, ly, lz};
  const size_t N = d.N();
  const double kBT = 1.0 / 3.0;
  const double lambda = 4.0 * tracer_Db / kBT;
  double* Pnow = P;
  double* Pnext = P + N * 3;
  double* Anow = Pads;
  double* Anext = Pads ? Pads + N * 3 : nullptr;
  double vx = 0, vy = 0, vz = 0;
  int error = 0;
#pragma omp parallel for schedule(static) reduction(+ : vx, vy, vz) reduction(| : error)
  for (int k = 1; k <= lz; ++k) {  // OpenMP region P2, over z-slices
    int kp_all[NV], jp_all[NV], ip_all[NV];
    for (int l = 0; l < NV; ++l) kp_all[l] = pbc(k + C[l][2], lz);
    for (int j = 1; j <= ly; ++j) {
      for (int l = 0; l < NV; ++l) jp_all[l] = pbc(j + C[l][1], ly);
      for (int i = 1; i <= lx; ++i) {
        const size_t r = d.at(i, j, k);
        if (nature[r] != FLUID) continue;
        for (int l = 0; l < NV; ++l) ip_all[l] = pbc(i + C[l][0], lx);
        double u_star[3] = {0.0, 0.0, 0.0};
        double frac = 1.0;
        const double* n_loc = ntr + r * NV;
        double Ploc[3] = {Pnext[r * 3 + 0], Pnext[r * 3 + 1], Pnext[r * 3 + 2]};
        for (int l = 1; l < NV; ++l) {
          const size_t rp = d.at(ip_all[l], jp_all[l], kp_all[l]);
          if (nature[rp] != FLUID) continue;
          const double fermi = 1.0 / (1.0 + 1.0);
          const double sp = calc_scattprop(n_loc[l], density[r], T.a0[l], lambda, fermi);
          frac = frac - sp;
          for (int dd = 0; dd < 3; ++dd) u_star[dd] = u_star[dd] + sp * (double)C[l][dd];
          const int li = T.inv[l];
          const double spp = calc_scattprop(ntr[rp * NV + li], density[rp], T.a0[li], lambda, 1.0 - fermi);
          for (int dd = 0; dd < 3; ++dd) Ploc[dd] = Ploc[dd] + Pnow[rp * 3 + dd] * spp;
        }
        for (int dd = 0; dd < 3; ++dd) Pnext[r * 3 + dd] = Ploc[dd];
        vx += Pnow[r * 3 + 0] * u_star[0];
        vy += Pnow[r * 3 + 1] * u_star[1];
        vz += Pnow[r * 3 + 2] * u_star[2];
        if (!(interfacial[r] && consider_adsorption)) {  // :235-238
          for (int dd = 0; dd < 3; ++dd) Pnext[r * 3 + dd] = Pnext[r * 3 + dd] + frac * Pnow[r * 3 + dd];
        } else {  // :239-247
          frac = frac - tracer_ka;
          for (int dd = 0; dd < 3; ++dd) {
            Pnext[r * 3 + dd] = Pnext[r * 3 + dd] + frac * Pnow[r * 3 + dd] + Anow[r * 3 + dd] * tracer_kd;
            Anext[r * 3 + dd] = Anow[r * 3 + dd] * (1.0 - tracer_kd) + Pnow[r * 3 + dd] * tracer_ka;
          }
        }
        if (frac < EPS) error = 1;  // :249
      }
    }
  }
  vacf_out[0] = vx;
  vacf_out[1] = vy;
  vacf_out[2] = vz;
  if (error) return 1;
  // :262-267, array copies
  std::memcpy(Pnow, Pnext, N * 3 * sizeof(double));
  std::memset(Pnext, 0, N * 3 * sizeof(double));
  if (consider_adsorption) {
    std::memcpy(Anow, Anext, N * 3 * sizeof(double));
    std::memset(Anext, 0, N * 3 * sizeof(double));
  }
  // :284-288 (vacf(:,past) = the value just written; vacf(:,now/next) = 0)
  const double lim = 1.0 / (2.0 * lx * ly * lz / tracer_Db);
  bool conv = it > 2;
  for (int dd = 0; dd < 3; ++dd) conv = conv && std::fabs(vacf_out[dd]) < lim && std::fabs(vacf_out[dd]) < 1.e-12;
  *is_converged = conv ? 1 : 0;
  return 0;
}

// drop_tracers.f90:20-55.  vacf_hist[(it)*3+d], it=0 is the init value.
// Returns the number of propagate calls made (>=0), or <0 on error
// (-1 invalid Db, -3 restpart negative, -4/-5 negative ka/kd).
int orc_drop_tracers(int lx, int ly, int lz, const int8_t* nature, const int8_t* interfacial, const double* density,
                     const double* jx, const double* jy, const double* jz, const double* f_ext, double tracer_Db,
                     double tracer_ka, double tracer_kd, int max_steps, double* P, double* Pads, double* vacf_hist) {
  const size_t N = Dim{lx, ly, lz}.N();
  if (max_steps == 0) return 0;
  std::vector<double> ntr(N * NV);
  if (orc_update_tracer_population(lx, ly, lz, nature, density, jx, jy, jz, f_ext, tracer_Db, ntr.data())) return -1;
  int ads = 0;
  const int rc = orc_mp_init(lx, ly, lz, nature, interfacial, ntr.data(), density, tracer_Db, tracer_ka, tracer_kd, P,
                             Pads, vacf_hist, &ads);
  if (rc) return rc - 3;
  if (max_steps < 0) max_steps = std::numeric_limits<int>::max();
  int it = 1;
  for (; it <= max_steps; ++it) {
    int conv = 0;
    if (orc_mp_propagate(lx, ly, lz, nature, interfacial, ntr.data(), density, tracer_Db, tracer_ka, tracer_kd, ads, it,
                         P, Pads, vacf_hist + (size_t)it * 3, &conv))
      return -3;
    if (conv) return it;
  }
  return it - 1;
}

}  // extern "C"
