// xee_diagnose — C++ re-host of the reference driver `bin/diagnose < diag.txt` (src/diagnose/main.f90 and its
// includes) on top of the C-ABI library (include/xee_b200.h).  Same stdin format, same raw float32 field files,
// same output file names and shapes, same result.txt; all field compute runs on the GPU through the library.
// It exists because the Fortran drop-in (fortran/elliptic_tools.f90) cannot be compiled in an image without a
// Fortran compiler; this makes "driver, input format and output layout unchanged" testable end to end.
//
//   main.f90:13-21        debug level from ./debug_mode_1 / ./debug_mode_2
//   read-input.f90:1-118  stdin parser (read_input_tools.f90:7-62: '//' comments, blank lines, 256-char lines)
//   initialize-variables.f90:33-129   field reads, geometry, a/b/c, solver_[abc]-*.bin
//   diagnose.f90:1-55     BAROTROPIC / BAROCLINIC passes: cal_coe -> solve_elliptic -> cal_eta | cal_uw -> writes
//   write-output.f90:1-3  result.txt
//
// Usage: xee_diagnose [--r8] < diag.txt        (--r8: promoted build, the -freal-4-real-8 equivalent)
#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/xee_b200.h"

namespace {

// read_input_tools.f90:7-38
bool read_input(std::istream& in, std::string& line) {
  std::string buf;
  while (std::getline(in, buf)) {
    if (buf.size() > 256) buf.resize(256);
    const size_t k = buf.find("//");
    if (k != std::string::npos) buf.resize(k);
    while (!buf.empty() && (buf.back() == ' ' || buf.back() == '\t' || buf.back() == '\r')) buf.pop_back();
    size_t b = 0;
    while (b < buf.size() && buf[b] == ' ') ++b;     // list-directed reads and == comparisons ignore leading blanks
    if (b == buf.size()) continue;
    line = buf.substr(b);
    return true;
  }
  std::fprintf(stderr, "At line 21 of file read_input_tools.f90: End of file\n");   // what the Fortran runtime would say
  std::exit(2);
}
// read_input_tools.f90:41-62
int split_line(std::string& line, std::string& out, const char* delim) {
  const size_t i = line.find(delim);
  if (i == std::string::npos) { out = line; line.clear(); return 1; }
  out = line.substr(0, i); line = line.substr(i + 1);
  return 0;
}
void error_msg(const char* where, int code, const std::string& msg) {   // message_tools.f90:6-12
  std::printf(" [%s] Error(%d): %s\n", where, code, msg.c_str());
}
std::vector<double> numbers(const std::string& s) {   // read(buffer,*): blank- or comma-separated
  std::string t = s;
  for (char& c : t) if (c == ',') c = ' ';
  std::istringstream is(t);
  std::vector<double> v; double x;
  while (is >> x) v.push_back(x);
  return v;
}

template <class R> struct Api;
template <> struct Api<float> {
  static constexpr auto cal_coe = xee_cal_coe_f32; static constexpr auto solve = xee_solve_elliptic_f32;
  static constexpr auto abc = xee_build_abc_f32; static constexpr auto eta = xee_cal_eta_f32; static constexpr auto uw = xee_cal_uw_f32;
};
template <> struct Api<double> {
  static constexpr auto cal_coe = xee_cal_coe_f64; static constexpr auto solve = xee_solve_elliptic_f64;
  static constexpr auto abc = xee_build_abc_f64; static constexpr auto eta = xee_cal_eta_f64; static constexpr auto uw = xee_cal_uw_f64;
};

// field_tools.f90:30-76: one direct-access record of 4*nx*ny bytes = headerless little-endian float32, i fastest.
template <class R>
void read_2Dfield(const std::string& fn, std::vector<R>& f, int nx, int ny) {
  std::vector<float> raw((size_t)nx * ny, 0.f);
  FILE* fp = std::fopen(fn.c_str(), "rb");
  if (!fp || std::fread(raw.data(), 4, raw.size(), fp) != raw.size()) {
    std::printf(" Reading field error. File name: %s\n", fn.c_str());
    if (!fp) { std::fprintf(stderr, "xee_diagnose: cannot open %s\n", fn.c_str()); std::exit(2); }
  }
  if (fp) std::fclose(fp);
  f.assign(raw.begin(), raw.end());
}
template <class R>
void write_2Dfield(const std::string& fn, const std::vector<R>& f, int nx, int ny) {
  std::vector<float> raw(f.begin(), f.begin() + (size_t)nx * ny);
  FILE* fp = std::fopen(fn.c_str(), "wb");
  if (!fp || std::fwrite(raw.data(), 4, raw.size(), fp) != raw.size()) std::printf(" Writing field error. File name: %s\n", fn.c_str());
  if (fp) std::fclose(fp);
}
bool exists(const char* p) { struct stat st; return ::stat(p, &st) == 0; }

template <class R>
int run() {
  enum { DYNAMIC_EFFICIENCY = 0, SECONDARY_CIRCULATION = 1, NONE = 2 };
  std::printf(" Dynamic Efficiency Diagnose Program\n");
  int debug_mode = 0;                                   // main.f90:13-21
  if (exists("./debug_mode_1")) debug_mode = 1;
  if (exists("./debug_mode_2")) debug_mode = 2;
  // ---------------------------------------------------------------- read-input.f90
  std::string mode_str, word[4], buffer;
  read_input(std::cin, mode_str);
  for (int i = 0; i < 4; ++i)
    if (split_line(mode_str, word[i], "-") != 0 && i != 3) { error_msg("INIT", 1, "There should be 4 inputs separated by dashes."); return 0; }
  int diag_param, geometry, density_mode, operator_complexity;
  if (word[0] == "DYNAMIC_EFFICIENCY") diag_param = DYNAMIC_EFFICIENCY;
  else if (word[0] == "SECONDARY_CIRCULATION") diag_param = SECONDARY_CIRCULATION;
  else if (word[0] == "NONE") diag_param = NONE;
  else { error_msg("INIT", 1, "Unknown Mode [" + word[0] + "]"); return 0; }
  if (word[1] == "CYLINDRICAL") geometry = 0;
  else if (word[1] == "SPHERICAL") geometry = 1;
  else { error_msg("INIT", 1, "Unknown Mode [" + word[1] + "]"); return 0; }
  if (word[2] == "DENSITY_NORMAL") density_mode = 0;
  else if (word[2] == "DENSITY_BOUSSINESQ") density_mode = 1;
  else { error_msg("INIT", 1, "Unknown Mode [" + word[2] + "]"); return 0; }
  if (word[3] == "BARO_ALL") operator_complexity = 2;
  else if (word[3] == "BAROCLINIC") operator_complexity = 1;
  else if (word[3] == "BAROTROPIC") operator_complexity = 0;
  else { error_msg("INIT", 1, "Unknown Mode [" + word[3] + "]"); return 0; }
  R Lr[2] = {0, 0}, Lz[2] = {0, 0}, Lat[2] = {R(-90.0), R(90.0)}, planet_radius = 0;
  const R MATH_PI = std::acos(R(-1.0)), DEG2RAD = MATH_PI / R(180.0);
  read_input(std::cin, buffer);
  {
    const std::vector<double> v = numbers(buffer);
    if (geometry == 0) {
      if (v.size() < 4) { std::fprintf(stderr, "xee_diagnose: domain line needs 4 numbers\n"); return 2; }
      Lr[0] = (R)v[0]; Lr[1] = (R)v[1]; Lz[0] = (R)v[2]; Lz[1] = (R)v[3];
      if (Lr[1] <= Lr[0]) error_msg("INIT", 1, "Domain size in radial direction must be positive.");
      if (Lz[1] <= Lz[0]) error_msg("INIT", 1, "Domain size in z direction must be positive.");
    } else {
      if (v.size() < 3) { std::fprintf(stderr, "xee_diagnose: domain line needs 3 numbers\n"); return 2; }
      planet_radius = (R)v[0]; Lz[0] = (R)v[1]; Lz[1] = (R)v[2];
      Lr[0] = Lat[0] * DEG2RAD * planet_radius; Lr[1] = Lat[1] * DEG2RAD * planet_radius;
      if (Lz[1] <= Lz[0]) error_msg("INIT", 1, "Domain size in z direction must be positive.");
    }
  }
  read_input(std::cin, buffer);
  const std::vector<double> np = numbers(buffer);
  const int nr = (int)np.at(0), nz = (int)np.at(1);
  std::string input_folder, output_folder, A_file, B_file, C_file, forcing_file, bc_init_file;
  read_input(std::cin, input_folder); read_input(std::cin, output_folder);
  read_input(std::cin, A_file); read_input(std::cin, B_file); read_input(std::cin, C_file);
  if (diag_param == SECONDARY_CIRCULATION) read_input(std::cin, forcing_file);
  read_input(std::cin, bc_init_file);
  read_input(std::cin, buffer);
  const std::vector<double> rc = numbers(buffer);
  const R saved_r1 = (R)rc.at(0), saved_r2 = (R)rc.at(1); const int saved_max_iter = (int)rc.at(2); const R alpha_strf = (R)rc.at(3);
  std::printf(" ----- Diagnose Input -----\n");
  std::printf(" Diagnose parameter:  %11d\n Geometry:  %11d\n Density distribution:  %11d\n Operator complexity:  %11d\n", diag_param, geometry, density_mode, operator_complexity);
  if (geometry == 0) std::printf(" Lr:  %.7E  %.7E\n Lz:  %.7E  %.7E\n", (double)Lr[0], (double)Lr[1], (double)Lz[0], (double)Lz[1]);
  else std::printf(" Using spherical mode, domain is forced to be global.\n Planet Radius:   %.7E\n Lat:  %.7E  %.7E\n Lz:  %.7E  %.7E\n", (double)planet_radius, (double)Lat[0], (double)Lat[1], (double)Lz[0], (double)Lz[1]);
  std::printf(" nr: %11d , nz: %11d\n Input folder:  %s\n Output folder: %s\n A file:        %s\n B file:        %s\n C file:        %s\n", nr, nz, input_folder.c_str(), output_folder.c_str(), A_file.c_str(), B_file.c_str(), C_file.c_str());
  if (diag_param == SECONDARY_CIRCULATION) std::printf(" forcing file:  %s\n", forcing_file.c_str());
  std::printf(" bc_init file:  %s\n absolute, relative residue, iter:   %.7E  %.7E %11d  %.7E\n --------------------------\n", bc_init_file.c_str(), (double)saved_r1, (double)saved_r1, saved_max_iter, (double)alpha_strf);
  std::printf(" Read input complete.\n");
  // ---------------------------------------------------------------- initialize-variables.f90
  const size_t nn = (size_t)nr * nz;
  std::vector<R> A, B, C, bc_init, forcing(nn, R(0));
  std::printf(" Allocation complete.\n");
  read_2Dfield(input_folder + "/" + A_file, A, nr, nz);
  read_2Dfield(input_folder + "/" + B_file, B, nr, nz);
  read_2Dfield(input_folder + "/" + C_file, C, nr, nz);
  read_2Dfield(input_folder + "/" + bc_init_file, bc_init, nr, nz);
  if (diag_param == SECONDARY_CIRCULATION) read_2Dfield(input_folder + "/" + forcing_file, forcing, nr, nz);
  else if (diag_param == DYNAMIC_EFFICIENCY) for (size_t q = 0; q < nn; ++q) forcing[q] = -B[q];      // :41
  // constants.f90:4-5
  const R g0 = R(9.8), theta0 = R(298.0), Rd = R(287.0), Cv = R(5.0) / R(2.0) * Rd, Cp = Cv + Rd, kappa = Rd / Cp,
          h0 = Cp * theta0 / g0, p0 = R(101300.0);
  const R dr = (Lr[1] - Lr[0]) / R(nr - 1), dz = (Lz[1] - Lz[0]) / R(nz - 1);                       // :45
  std::vector<R> ra(nr), rcuva(nr), za(nz), exner(nz), rho(nz);
  for (int i = 1; i <= nr; ++i) ra[i - 1] = Lr[0] + R(i - 1) * dr;
  for (int j = 1; j <= nz; ++j) {
    za[j - 1] = Lz[0] + R(j - 1) * dz;
    exner[j - 1] = density_mode == 0 ? (R(1.0) - za[j - 1] / h0) : R(1.0);
    rho[j - 1] = density_mode == 0 ? p0 / (theta0 * Rd) * std::pow(exner[j - 1], R(1.0) / kappa - R(1.0)) : R(1.0);
  }
  if (geometry == 0) rcuva = ra;
  else {   // :61-66, restated as written (cos() of degrees is the reference's behaviour)
    const R dlat = (Lat[1] - Lat[0]) / R(nr - 1);
    for (int i = 1; i <= nr; ++i) rcuva[i - 1] = planet_radius * std::cos(Lat[0] + R(i - 1) * dlat);
  }
  std::printf(" Geometry complete.\n");
  std::vector<R> sa((size_t)(nr - 1) * (nz - 2)), sb((size_t)(nr - 1) * (nz - 1)), sc((size_t)(nr - 2) * (nz - 1));
  Api<R>::abc(A.data(), B.data(), C.data(), rcuva.data(), rho.data(), sa.data(), sb.data(), sc.data(), &nr, &nz);   // :72-95 on the GPU
  const std::vector<R> saved_sb = sb;
  std::printf(" Solver coe part I complete.\n Solver coe part II complete.\n");
  write_2Dfield(output_folder + "/solver_a-sA.bin", sa, nr - 1, nz - 2);
  write_2Dfield(output_folder + "/solver_b-B.bin", sb, nr - 1, nz - 1);
  write_2Dfield(output_folder + "/solver_c-sC.bin", sc, nr - 2, nz - 1);
  std::printf(" Solver complete.\n Initialization complete.\n");
  // ---------------------------------------------------------------- diagnose.f90
  const auto t_beg = std::chrono::steady_clock::now();
  std::vector<R> f = forcing, coe(9 * nn, R(0)), strf(nn), wksp(nn), eta((size_t)(nr - 1) * nz), w_A((size_t)(nr - 1) * nz), u_C((size_t)nr * (nz - 1));
  int err = 0;
  auto pass = [&](bool barotropic) {
    std::printf(barotropic ? " Solving CHI with L(A,B=0,C) = -B\n" : " Solving CHI with L(A,B,C) = -B\n");
    if (barotropic) std::fill(sb.begin(), sb.end(), R(0)); else sb = saved_sb;
    Api<R>::cal_coe(sa.data(), sb.data(), sc.data(), coe.data(), &dr, &dz, &nr, &nz, &err);
    int max_iter = saved_max_iter; R r1 = saved_r1, r2 = saved_r2; const R alpha = alpha_strf;
    strf = bc_init;
    const int cs = 100, ct = 10, lr = 5;
    std::fflush(stdout);
    Api<R>::solve(&max_iter, &cs, &ct, &lr, &r1, &r2, &alpha, strf.data(), coe.data(), f.data(), wksp.data(), &nr, &nz, &err, &debug_mode);
    std::printf(" Relaxation uses  %11d  steps. Final residue is   %.7E ,  %.7E\n", max_iter, (double)r1, (double)r2);
    const std::string tag = barotropic ? "[BAROTROPIC]" : "[BAROCLINIC]";
    if (diag_param == DYNAMIC_EFFICIENCY) {
      Api<R>::eta(strf.data(), eta.data(), ra.data(), rcuva.data(), rho.data(), exner.data(), &nr, &nz);
      write_2Dfield(output_folder + "/eta-" + tag + "-A.bin", eta, nr - 1, nz);
      write_2Dfield(output_folder + "/rchi-" + tag + "-O.bin", strf, nr, nz);
    } else if (diag_param == SECONDARY_CIRCULATION) {
      Api<R>::uw(strf.data(), u_C.data(), w_A.data(), ra.data(), rcuva.data(), za.data(), rho.data(), &nr, &nz);
      write_2Dfield(output_folder + "/w-" + tag + "-A.bin", w_A, nr - 1, nz);
      write_2Dfield(output_folder + "/u-" + tag + "-C.bin", u_C, nr, nz - 1);
      write_2Dfield(output_folder + "/rpsi-" + tag + "-O.bin", strf, nr, nz);
    }
  };
  if (operator_complexity == 0 || operator_complexity == 2) pass(true);
  if (operator_complexity == 1 || operator_complexity == 2) pass(false);
  const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_beg).count();
  std::printf(" Diagnose complete.\n");
  // ---------------------------------------------------------------- write-output.f90
  FILE* fp = std::fopen((output_folder + "/result.txt").c_str(), "w");
  if (fp) { std::fprintf(fp, " Time elapsed (sec) :   %.8E\n", elapsed); std::fclose(fp); }
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  bool r8 = false;
  for (int i = 1; i < argc; ++i) if (!std::strcmp(argv[i], "--r8")) r8 = true;
  return r8 ? run<double>() : run<float>();
}
