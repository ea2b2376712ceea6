// laboetie_driver.cpp -- compiled mirror of the reference's Fortran driver, on top of the C ABI.
//
// The north-star keeps the Fortran driver (lb.in parsing, geom.in / geometryLabel set-up, the
// equilibration phases, output files) and moves only the loop bodies behind ISO_C_BINDING
// (fortran/laboetie_gpu_iface.f90).  No Fortran compiler exists in the build image, so this C++
// program restates that driver's control flow and file formats and is what the tests run:
//   main.f90:12-34               init_simu -> equilibration -> drop_tracers
//   module_input.f90:62-110      `tag = value` lookup in ./lb.in, `#` comments
//   supercell_definition.f90     geometryLabel -1, 0 (geom.in), 1, 2, 3, 11 (geom.pbm)
//   equilibration.f90:143-557    time loop, convergence state machine, profile / l2err files
//   drop_tracers.f90:20-55       moment propagation loop, vacf.dat
// Numbers written are the reference's quantities; the text layout is C's %24.16E rather than
// Fortran list-directed output.
//
// usage: laboetie_driver [--input lb.in] [--outdir output] [--check-input] [--quiet]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/laboetie_gpu.h"

namespace {

struct Input {
  std::map<std::string, std::string> kv;
  bool load(const std::string& path) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    while (std::getline(f, line)) {
      const size_t h = line.find('#');
      if (h != std::string::npos) line.erase(h);
      const size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      std::string tag = line.substr(0, eq), val = line.substr(eq + 1);
      auto trim = [](std::string& s) {
        const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
        s = (a == std::string::npos) ? "" : s.substr(a, b - a + 1);
      };
      trim(tag);
      trim(val);
      if (!tag.empty() && !kv.count(tag)) kv[tag] = val;  // the reference takes the first match
    }
    return true;
  }
  bool has(const std::string& t) const { return kv.count(t) != 0; }
  double dp(const std::string& t, double def) const { return has(t) ? std::atof(kv.at(t).c_str()) : def; }
  long integer(const std::string& t, long def) const { return has(t) ? std::atol(kv.at(t).c_str()) : def; }
  bool logical(const std::string& t, bool def) const {
    if (!has(t)) return def;
    const std::string& v = kv.at(t);
    return !v.empty() && (v[0] == 'T' || v[0] == 't' || v.find(".true.") != std::string::npos || v.find(".TRUE.") != std::string::npos);
  }
  void dp3(const std::string& t, double out[3]) const {
    out[0] = out[1] = out[2] = 0.0;
    if (!has(t)) return;
    std::istringstream ss(kv.at(t));
    ss >> out[0] >> out[1] >> out[2];
  }
};

[[noreturn]] void stop(const std::string& msg) {
  std::fprintf(stderr, "%s\n", msg.c_str());
  std::exit(1);
}

void ck(int rc, lbg_handle h, const char* what) {
  if (rc == LBG_OK) return;
  // the reference's own stop messages (equilibration.f90:248, module_moment_propagation.f90:257, ...)
  stop(std::string("ERROR STOP in ") + what + ": " + lbg_status_string(rc) + " [" + lbg_last_error(h) + "]");
}

// module_geometry.f90:158-166, :253-277, :206-245 with tie-free integer thresholds; :380-426; :430-511
std::vector<int8_t> build_geometry(int label, int lx, int ly, int lz, const std::string& dir) {
  std::vector<int8_t> nat((size_t)lx * ly * lz, 0);
  auto at = [&](int i, int j, int k) -> int8_t& { return nat[(size_t)i + (size_t)lx * ((size_t)j + (size_t)ly * k)]; };
  switch (label) {
    case -1:
      break;
    case 0: {
      std::ifstream f(dir + "geom.in");
      if (!f) stop("Cant find file containing custom geometry: geom.in. Check lb.in if you really wanted custom geometry");
      long i, j, k;
      while (f >> i >> j >> k) {
        if (i <= 0 || j <= 0 || k <= 0 || i > lx || j > ly || k > lz) stop("Index in geom.in out of range. It must be between 1 and lx/ly/lz");
        at((int)i - 1, (int)j - 1, (int)k - 1) = 1;
      }
      break;
    }
    case 1:
      for (int j = 0; j < ly; ++j)
        for (int i = 0; i < lx; ++i) at(i, j, 0) = at(i, j, lz - 1) = 1;
      break;
    case 2: {
      if (lx != ly) stop("wall=2 is for cylinders, which should have same lx and ly");
      if (lx < 3) stop("the diameter of the cylinder (lx) should be greater than 3");
      for (int j = 0; j < ly; ++j)
        for (int i = 0; i < lx; ++i) {
          const long dx = 2L * (i + 1) - (lx + 1), dy = 2L * (j + 1) - (ly + 1);
          const int8_t v = (dx * dx + dy * dy >= (long)(lx - 1) * (lx - 1)) ? 1 : 0;  // |r-o| >= (lx-1)/2
          for (int k = 0; k < lz; ++k) at(i, j, k) = v;
        }
      break;
    }
    case 3: {
      if (lx != ly || lx != lz) stop("with wall = 3, i.e. cfc cell, the supercell should be cubic with lx=ly=lz");
      const long thr = 3L * (lx - 1) * (lx - 1);  // 16 d^2 <= 3 (lx-1)^2
      const long c2[2] = {2, 2L * lx};
      for (int k = 0; k < lz; ++k)
        for (int j = 0; j < ly; ++j)
          for (int i = 0; i < lx; ++i) {
            const long X = 2L * (i + 1), Y = 2L * (j + 1), Z = 2L * (k + 1);
            bool in = false;
            for (int a = 0; a < 2 && !in; ++a)
              for (int b = 0; b < 2 && !in; ++b)
                for (int c = 0; c < 2 && !in; ++c) {
                  const long d2 = (X - c2[a]) * (X - c2[a]) + (Y - c2[b]) * (Y - c2[b]) + (Z - c2[c]) * (Z - c2[c]);
                  in = 4 * d2 <= thr;
                }
            if (!in) {
              const long m = lx + 1;
              in = 4 * ((X - m) * (X - m) + (Y - m) * (Y - m) + (Z - m) * (Z - m)) <= thr;
            }
            at(i, j, k) = in ? 1 : 0;
          }
      break;
    }
    case 11: {
      if (lx != 1) stop("lx must be 1 if geom.pbm is used");
      std::ifstream f(dir + "geom.pbm");
      std::string magic;
      int ncol = 0, nline = 0;
      if (!(f >> magic) || magic != "P1") stop("geom.pbm doesnt seem to be valid. It's magic number (first line) is not P1");
      f >> ncol >> nline;
      if (ncol != ly) stop("ncolumn /= ny in geom.pbm");
      if (nline != lz) stop("nline /= nz in geom.pbm");
      for (int k = 0; k < nline; ++k)
        for (int j = 0; j < ncol; ++j) {
          char c;
          if (!(f >> c) || (c != '0' && c != '1')) stop("module_geometry. Pbm format allow 0 or 1 only");
          if (c == '1') at(0, j, k) = 1;
        }
      break;
    }
    default:
      stop("supercell%geometry%label tag in input file is invalid (this driver mirrors labels -1,0,1,2,3,11)");
  }
  return nat;
}

void write_profiles(lbg_handle h, int lx, int ly, int lz, FILE* fz, FILE* fy, FILE* fx, FILE* dz, FILE* dy, FILE* dx) {
  const int len[3] = {lx, ly, lz};
  FILE* ff[3] = {fx, fy, fz};
  FILE* dd[3] = {dx, dy, dz};
  for (int axis = 0; axis < 3; ++axis) {
    std::vector<double> p((size_t)len[axis] * 4);
    ck(lbg_lb_profiles(h, axis, 0, p.data()), h, "equilibration (profiles)");
    for (int k = 0; k < len[axis]; ++k) {
      std::fprintf(ff[axis], "%12d %24.16E %24.16E %24.16E\n", k + 1, p[4 * k], p[4 * k + 1], p[4 * k + 2]);
      std::fprintf(dd[axis], "%12d %24.16E\n", k + 1, p[4 * k + 3]);
    }
  }
}

}  // namespace

int main(int argc, char** argv) {
  std::string input = "lb.in", outdir = "output";
  bool check_only = false, quiet = false;
  for (int a = 1; a < argc; ++a) {
    const std::string s = argv[a];
    if (s == "--input" && a + 1 < argc) input = argv[++a];
    else if (s == "--outdir" && a + 1 < argc) outdir = argv[++a];
    else if (s == "--check-input") check_only = true;
    else if (s == "--quiet") quiet = true;
    else stop("usage: laboetie_driver [--input lb.in] [--outdir output] [--check-input] [--quiet]");
  }
  Input in;
  if (!in.load(input)) stop("cannot open input file " + input);
  std::string dir = input;
  const size_t slash = dir.find_last_of('/');
  dir = (slash == std::string::npos) ? "" : dir.substr(0, slash + 1);

  // ---- init_simu.f90 / supercell_definition.f90 / module_lbmodel.f90 ----
  const std::string lbmodel = in.has("lbmodel") ? in.kv["lbmodel"] : "D3Q19";
  if (lbmodel.rfind("D3Q19", 0) != 0) stop("You ask for a DnQm lattice that is not implemented");  // module_lbmodel.f90:138-149
  const int lx = (int)in.integer("lx", 0), ly = (int)in.integer("ly", 0), lz = (int)in.integer("lz", 0);
  if (lx <= 0 || ly <= 0 || lz <= 0) stop("lx, ly, lz must be > 0");
  const int label = (int)in.integer("geometryLabel", 0);
  const double eps = 2.220446049250313e-16;
  if (std::fabs(in.dp("sigma", 0.0)) > eps) stop("ERROR: laboetie can only consider uncharged systems.");  // equilibration.f90:28-33
  if (in.logical("first_order_only", false)) stop("first_order_only = T is not supported (undefined behaviour in the reference, module_collision.f90:110-125)");
  const bool compensate = in.logical("compensate_f_ext", false);
  const double tau = in.dp("relaxation_time", 1.0);
  const double target_error = in.dp("target_error", 1.e-10);
  const double rho0 = in.dp("initialSolventDensity", 1.0);
  const long print_frequency = in.integer("print_frequency", std::max(50000L / ((long)lx * ly * lz), 1L));
  const long print_files_frequency = in.integer("print_files_frequency", 2147483647L);
  const bool write_total_mass_flux = in.logical("write_total_mass_flux", false);
  double f_ext[3];
  in.dp3("f_ext", f_ext);
  const long max_mp = in.integer("maximum_moment_propagation_steps", 0);
  const double Db = in.dp("tracer_Db", 0.0), ka = in.dp("tracer_ka", 0.0), kd = in.dp("tracer_kd", 0.0);
  if (std::fabs(in.dp("tracer_Ds", 0.0)) > eps) stop("Tracers you defined have non-zero surface diffusion coefficient. This is not implemented yet");
  if (std::fabs(in.dp("tracer_z", 0.0)) > eps) stop("charged tracers are not implemented");
  const bool print_vacf = in.logical("print_vacf", true);

  std::vector<int8_t> nature = build_geometry(label, lx, ly, lz, dir);
  long nsolid = 0;
  for (int8_t v : nature) nsolid += v;
  if (!quiet)
    std::printf(" laboetie (B200 driver mirror): %d x %d x %d, geometryLabel %d, %ld solid nodes, f_ext = %g %g %g\n", lx, ly,
                lz, label, nsolid, f_ext[0], f_ext[1], f_ext[2]);
  if (check_only) return 0;

  const std::string mk = "mkdir -p '" + outdir + "'";
  if (std::system(mk.c_str()) != 0) stop("cannot create " + outdir);
  auto open = [&](const char* name) {
    FILE* f = std::fopen((outdir + "/" + name).c_str(), "w");
    if (!f) stop(std::string("cannot open ") + name);
    return f;
  };

  lbg_handle h = nullptr;
  ck(lbg_create(&h, lx, ly, lz, nature.data(), 0), nullptr, "supercell_definition");
  ck(lbg_lb_init(h, rho0), h, "init_simu");

  // ---- equilibration.f90 ----
  FILE* fz = open("mass-flux_profile_along_z.dat");
  FILE* fy = open("mass-flux_profile_along_y.dat");
  FILE* fx = open("mass-flux_profile_along_x.dat");
  std::fprintf(fz, "# z, <rho.v_x>_{x,y}, <rho.v_y>_{x,y}, <rho.v_z>_{x,y}\n");
  std::fprintf(fy, "# y, <rho.v_x>_{x,z}, <rho.v_y>_{x,z}, <rho.v_z>_{x,z}\n");
  std::fprintf(fx, "# x, <rho.v_x>_{y,z}, <rho.v_y>_{y,z}, <rho.v_z>_{y,z}\n");
  FILE* dz = open("mean-density_profile_along_z.dat");
  FILE* dy = open("mean-density_profile_along_y.dat");
  FILE* dx = open("mean-density_profile_along_x.dat");
  FILE* l2f = open("l2err.dat");
  FILE* tmf = write_total_mass_flux ? open("total_mass_flux.dat") : nullptr;
  if (!quiet) std::printf("\n Lattice Boltzmann\n =================\n        step\n        ----\n");

  // compensate_f_ext (equilibration.f90:123-124,185-188,388-487): particle centre and probe file
  int pcx = lx / 2 + 1, pcy = ly / 2 + 1, pcz = lz / 2 + 1;
  if (in.has("particle_coordinates")) {
    std::istringstream ss(in.kv["particle_coordinates"]);
    ss >> pcx >> pcy >> pcz;
  }
  FILE* vcn = compensate ? open("v_centralnode.dat") : nullptr;
  bool without_fext = false;
  long t = 0, tfext = 0;
  std::vector<double> hist(4096);
  // profiles are written at t == 1 and every print_files_frequency steps, with the density / momentum
  // the step starts from (equilibration.f90:154-176); total flux every step if asked (:259-261)
  const bool every_step_io = write_total_mass_flux;
  for (;;) {
    long chunk = every_step_io ? 1 : 4096;
    if (compensate && without_fext) {  // :185-188: momentum at the particle centre before every step
      double pr[4];
      ck(lbg_lb_probe(h, pcx - 1, pcy - 1, pcz - 1, pr), h, "equilibration (v_centralnode)");
      std::fprintf(vcn, "%12ld %24.16E %24.16E %24.16E\n", t + 1 - tfext, pr[0], pr[1], pr[2]);
      chunk = 1;
    }
    const long next_write = (t / print_files_frequency + 1) * print_files_frequency;  // smallest multiple > t
    const bool write_now = (t + 1 == next_write) || (t + 1 == 1);
    if (write_now) {
      for (FILE* f : {fz, fy, fx, dz, dy, dx}) std::fprintf(f, "# timestep %ld\n", t + 1);
      write_profiles(h, lx, ly, lz, fz, fy, fx, dz, dy, dx);
      for (FILE* f : {fz, fy, fx}) std::fprintf(f, "\n");
    }
    // never run past the step before the next profile dump
    const long limit = (t + 1 == next_write) ? print_files_frequency : next_write - 1 - t;
    chunk = std::min(chunk, std::max(limit, 1L));
    if (tmf) {  // equilibration.f90:259-261 writes `t, SUM(jx), ...` before jx is updated in step t: the flux of step t-1
      double tf[3];
      ck(lbg_lb_total_flux(h, tf), h, "equilibration (total flux)");
      std::fprintf(tmf, "%12ld %15.7E %15.7E %15.7E\n", t + 1, tf[0], tf[1], tf[2]);
    }
    int done = 0, conv = 0;
    ck(lbg_lb_step(h, tau, (int)chunk, 1, target_error, hist.data(), &done, &conv), h, "equilibration");
    for (int i = 0; i < done; ++i) {
      ++t;
      std::fprintf(l2f, "%12ld %24.16E\n", t, hist[i]);
      if (!quiet && t % print_frequency == 0) std::printf(" %11ld %14.7E (target %10.3E )\n", t, hist[i], target_error);
    }
    if (!conv) continue;
    if (!without_fext) {  // :377-386
      without_fext = true;
      tfext = t + 1;
      if (!compensate) {
        ck(lbg_lb_set_force_uniform(h, f_ext), h, "equilibration (f_ext)");
      } else {  // :388-487
        const int pd = (int)in.integer("dominika_particle_diameter", 1);
        if (pd % 2 == 0) stop("ERROR: l. 285 particle diameter must be odd");
        if (lx % 2 == 0 || ly % 2 == 0 || lz % 2 == 0)
          stop("ERROR: when compensate_f_ext, there should be odd number of nodes in all directions");
        const int pdr = pd / 2;
        const size_t N = nature.size();
        std::vector<double> fx(N, 0.0), fy(N, 0.0), fz(N, 0.0);
        std::vector<char> part(N, 0);
        long l = 0, fluid_nodes = 0;
        for (int8_t v : nature) fluid_nodes += (v == 0);
        for (int i = pcx - pdr; i <= pcx + pdr; ++i)
          for (int j = pcy - pdr; j <= pcy + pdr; ++j)
            for (int k = pcz - pdr; k <= pcz + pdr; ++k) {
              const long d2 = (long)(i - pcx) * (i - pcx) + (long)(j - pcy) * (j - pcy) + (long)(k - pcz) * (k - pcz);
              if (4 * d2 > (long)pd * pd) continue;  // norm2 > pd/2
              if (i < 1 || j < 1 || k < 1 || i > lx || j > ly || k > lz) stop("particle outside the lattice");
              const size_t r = (size_t)(i - 1) + (size_t)lx * ((size_t)(j - 1) + (size_t)ly * (k - 1));
              if (nature[r] != 0) stop("ERROR: l306 of equilibration.f90. Dominika's particle at a solid node");
              fx[r] = f_ext[0];
              fy[r] = f_ext[1];
              fz[r] = f_ext[2];
              ++l;
            }
        for (size_t r = 0; r < N; ++r) {
          // the reference recognises particle nodes by comparing the three components with f_ext (:450,468)
          const bool match = fx[r] == f_ext[0] && fy[r] == f_ext[1] && fz[r] == f_ext[2];
          if (label == -1) {
            fx[r] = -f_ext[0] / (double)fluid_nodes + (match ? fx[r] / (double)l : 0.0);
            fy[r] = -f_ext[1] / (double)fluid_nodes + (match ? fy[r] / (double)l : 0.0);
            fz[r] = -f_ext[2] / (double)fluid_nodes + (match ? fz[r] / (double)l : 0.0);
          } else {
            fx[r] = match ? fx[r] / (double)l : 0.0;
            fy[r] = match ? fy[r] / (double)l : 0.0;
            fz[r] = match ? fz[r] / (double)l : 0.0;
          }
          if (nature[r] != 0) fx[r] = fy[r] = fz[r] = 0.0;
        }
        if (!quiet) std::printf("        Dominika's particle has diameter (lb units) %d, %ld nodes\n", pd, l);
        ck(lbg_lb_set_force_field(h, fx.data(), fy.data(), fz.data()), h, "equilibration (compensate_f_ext)");
      }
    } else {
      break;  // :373-374
    }
  }
  for (FILE* f : {fz, fy, fx, dz, dy, dx}) std::fprintf(f, "# Steady state with convergence criteria %15.7E\n", target_error);
  write_profiles(h, lx, ly, lz, fz, fy, fx, dz, dy, dx);
  for (FILE* f : {fz, fy, fx, dz, dy, dx, l2f}) std::fclose(f);
  if (tmf) std::fclose(tmf);
  if (vcn) std::fclose(vcn);
  {  // equilibration.f90:527-533
    std::vector<double> rho((size_t)lx * ly * lz), jx(rho.size()), jy(rho.size()), jz(rho.size());
    ck(lbg_lb_download_moments(h, rho.data(), jx.data(), jy.data(), jz.data()), h, "equilibration (write-back)");
    FILE* f2 = open("mass-flux_field_2d_at_x.eq.1.dat");
    for (int j = 0; j < ly; ++j)
      for (int k = 0; k < lz; ++k) {
        const size_t r = (size_t)0 + (size_t)lx * ((size_t)j + (size_t)ly * k);
        std::fprintf(f2, "%12d %12d %24.16E %24.16E\n", j + 1, k + 1, jy[r], jz[r]);
      }
    std::fclose(f2);
  }
  if (!quiet) std::printf(" equilibration: left the time loop at t = %ld (f_ext switched on at t = %ld)\n", t, tfext);

  // ---- drop_tracers.f90 ----
  if (max_mp != 0) {
    if (!quiet) std::printf("\n Moment propagation\n ==================\n");
    double v0[3];
    ck(lbg_mp_init(h, Db, ka, kd, f_ext, v0), h, "drop_tracers");
    FILE* vf = print_vacf ? open("vacf.dat") : nullptr;
    if (vf) {
      std::fprintf(vf, "# time t, VACF_x(t), VACF_y(t), VACF_z(t)\n");
      std::fprintf(vf, "%12d %24.16E %24.16E %24.16E\n", 0, v0[0], v0[1], v0[2]);
    }
    if (!quiet) std::printf(" %11d %14.7E %14.7E %14.7E\n", 0, v0[0], v0[1], v0[2]);
    long left = max_mp < 0 ? 2147483647L : max_mp, it = 0;
    std::vector<double> v(3 * 4096);
    int conv = 0;
    while (left > 0 && !conv) {
      int done = 0;
      ck(lbg_mp_step(h, (int)std::min(left, 4096L), v.data(), &done, &conv), h, "drop_tracers (propagate)");
      for (int i = 0; i < done; ++i) {
        ++it;
        if (vf) std::fprintf(vf, "%12ld %24.16E %24.16E %24.16E\n", it, v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        if (!quiet && it % 10000 == 0) std::printf(" %11ld %14.7E %14.7E %14.7E\n", it, v[3 * i], v[3 * i + 1], v[3 * i + 2]);
      }
      left -= done;
    }
    if (vf) std::fclose(vf);
    if (!quiet) std::printf(" moment propagation: %ld steps%s\n", it, conv ? " (converged)" : "");
  }
  lbg_destroy(h);
  return 0;
}
