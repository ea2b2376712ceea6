// lbg_internal.h -- shared between the kernel translation units and api.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "d3q19.cuh"

namespace lbg {

constexpr int BLOCK = 256;

// mask word per node: bit 0 = node is fluid; bit l (1..18) = node r+c_l is fluid;
// bit 19 = node is interfacial (supercell_definition.f90:115-147).
constexpr uint32_t MASK_FLUID = 1u;
constexpr uint32_t MASK_INTERFACIAL = 1u << 19;

// Slab geometry.  Arrays carry one halo plane below (plane 0) and one above
// (plane nzl+1); own planes are 1..nzl.  With zwrap (single slab) the halo
// planes of the fields are unused and z neighbours wrap inside the kernel.
struct Geo {
  int lx, ly, plane, nzl, zwrap;
  long long nalloc;  // plane * (nzl + 2)
};

// Device-side control block: lets a batch of step kernels stop itself at the
// reference's exit step without a host round trip per step.
struct Ctrl {
  int stop;                 // set by the first kernel that sees the criterion met
  int neg_step_idx;         // 1 + batch index of the first step with a negative population (0 = none)
  int stop_idx;             // 1 + batch index of the converged step
  unsigned int ticket;      // last-block election for the vacf reduction
};

enum ForceMode { FORCE_NONE = 0, FORCE_UNIFORM = 1, FORCE_FIELD = 2 };

struct LBArgs {
  Geo geo;
  d3q19::Consts k;
  const double* fin;   // 19 arrays, stride geo.nalloc: post-collision populations n*(t)
  double* fout;        // n*(t+1)
  const uint32_t* mask;
  long long g_begin, g_end;  // linear alloc index range to process
  double w1, w2, w3;         // 1-1/tau, 1/tau, 1-1/(2 tau)
  double fj[3];              // uniform force in effect for step t (enters j as f/2)
  double fc[3];              // uniform force of the collision that follows
  const double* fj_field;    // 3 arrays, stride nalloc (FORCE_FIELD)
  const double* fc_field;
  const double* jold;        // 3 arrays: j(t-1)
  double* jnew;              // 3 arrays: j(t)
  unsigned long long* l2_slots;  // one per batch step, bit pattern of a non-negative double
  int batch_idx;                 // index of this step in the batch
  int prev_checked;              // step batch_idx-1 was a checked step of this batch
  int prev_may_stop;             // ... and its global t-1 > 2
  double target;
  Ctrl* ctrl;
};

struct CollideArgs {
  Geo geo;
  d3q19::Consts k;
  const double* fin;  // pre-collision n(t)
  double* fout;
  const uint32_t* mask;
  const double* mom;  // rho, jx, jy, jz: 4 arrays, stride nalloc
  long long g_begin, g_end;
  double w1, w2, w3;
  double fc[3];
  const double* fc_field;
};

struct MomArgs {
  Geo geo;
  const double* fin;  // n*(t)
  const uint32_t* mask;
  double* mom;        // out: rho, jx, jy, jz
  double* pops;       // out (optional): n(t), 19 arrays stride nalloc
  long long g_begin, g_end;
  double fj[3];
  const double* fj_field;
};

struct MPInitArgs {
  Geo geo;
  d3q19::Consts k;
  const uint32_t* mask;
  const double* mom;   // rho, jx, jy, jz with valid halos (or zwrap)
  double* q;           // 18 arrays: incoming link probabilities q_l(r) = p_{inv l}(r + c_l)
  double* s;           // 4 arrays: remaining fraction (after -ka where adsorbing), u*_x, u*_y, u*_z
  double* P0;          // 3 arrays: Propagated_Quantity(:, now) at t=0
  long long g_begin, g_end;
  double f[3];
  double lambda_w[3];  // lambda * w per kind
  double bw;           // 1 / Pstat
  double ka;
  int ads;
  double* partial;     // per block: vacf0 x,y,z
  int* err;            // set if the remaining fraction < eps somewhere
};

struct MPArgs {
  Geo geo;
  const uint32_t* mask;
  const double* q;
  const double* s;
  const double* Pnow;   // 3 arrays
  double* Pnext;
  const double* Anow;   // adsorbed, 3 arrays
  double* Anext;
  int p_begin, p_end;   // plane range [p_begin, p_end) to process (own planes are 1..nzl)
  double ka, kd, one_minus_kd;
  int ads;
  double* partial;      // per block partial vacf (3 each)
  double* vacf_slots;   // 3 per batch step
  int batch_idx;
  int accumulate;       // add into the slot instead of overwriting (second launch of a split step)
  int nblocks_total;
  int check_prev;       // evaluate the convergence criterion on slot batch_idx-1
  double lim;           // 1/(2 lx ly lz / Db)
  Ctrl* ctrl;
};

struct ProfileArgs {
  Geo geo;
  const double* mom;
  int axis;
  double eps;
  double* out;  // 5 per row: sum jx, jy, jz, sum rho, count(rho > eps)
};

// launchers (lb_kernels.cu / mp_kernels.cu).  Each returns the number of kernels launched.
int launch_build_mask(const Geo& g, const int8_t* nature_halo, uint32_t* mask, cudaStream_t st);
int launch_lb_init(const Geo& g, const uint32_t* mask, double rho0, const double a0[3], double* f, double* mom,
                   cudaStream_t st);
int launch_collide(const CollideArgs& a, bool tau1, int fmode, int grid, cudaStream_t st);
int launch_lb_step(const LBArgs& a, bool tau1, int fmode, bool check, bool writej, int minb, int grid,
                   cudaStream_t st);
int launch_moments(const MomArgs& a, int fmode, int grid, cudaStream_t st);
int launch_fill_force(const Geo& g, const uint32_t* mask, const double f[3], double* field, cudaStream_t st);
int launch_profile(const ProfileArgs& a, int rows, cudaStream_t st);
int launch_count_flags(const Geo& g, const uint32_t* mask, unsigned long long* counts2, cudaStream_t st);
int launch_extract_flag(const Geo& g, const uint32_t* mask, uint32_t bit, int8_t* out_own, cudaStream_t st);
int launch_mp_init(const MPInitArgs& a, int grid, cudaStream_t st);
int launch_mp_step(const MPArgs& a, int variant, int grid, cudaStream_t st);
int launch_soa_to_aos3(const Geo& g, const double* soa, double* aos_own, cudaStream_t st);
int occupancy_grid_lb(int sm_count, int minb);
int occupancy_grid_mp(int sm_count, int variant);

}  // namespace lbg
