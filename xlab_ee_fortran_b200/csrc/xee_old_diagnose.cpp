// xee_old_diagnose — C++ re-host of the reference's LEGACY driver (src/old-diagnose/diagnose.f90, 1177 lines):
// the only place in the reference where the full chain heating -> secondary circulation -> kinetic-energy
// generation -> efficiency scalars exists (TENDENCY decomposition, up to nine elliptic solves per case).
// The driver's own single-pass loops stay host code, exactly as they are host Fortran in the reference; every
// cal_coe / solve_elliptic call (the hot path, diagnose.f90:449-714) goes to the GPU through the C-ABI
// (include/xee_b200.h, legacy 12-argument solve_elliptic).  Same stdin format, same .bin files, same output file
// names, efficiency.txt in the line format xtt-lib-python/XEffReader.py:15-28 parses.
//
// The legacy code carries latent bugs.  This re-host states the intended maths and says where:
//   [D1] Q and F are B-grid (nr-1, nz-1) files (the legacy code reads nr*nz values into (nr-1,nz-1) arrays, :241-242);
//   [D2] m2: seed column 1, accumulate i = 2..nr-1 (legacy :361-367 uses stale i,j and touches m2(0,j), ra(0));
//   [D3] rows of wksp_A that d_dz_B2A never writes (1, nz-1, nz) are 0 (uninitialised in the legacy code, :423-433, :499-500);
//   [D4] INSTANT mode: b_anomaly = 0 and theta = background state (uninitialised in the legacy code, :447-520);
//   [D5] cal_exchange_conversion uses real r, dr, dz (declared INTEGER in the legacy code, :1146);
//   [D6] efficiency.txt is opened also when only the BAROCLINIC block runs (legacy opens it in the first block only);
//   [D7] spherical geometry is rejected (the legacy code applies cos() to degrees, :275-279).
// Usage: xee_old_diagnose [--r8] < config.txt
#include <sys/stat.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/xee_b200.h"

namespace {

bool read_input(std::istream& in, std::string& line) {   // read_input_tools.f90:7-38
  std::string buf;
  while (std::getline(in, buf)) {
    if (buf.size() > 256) buf.resize(256);
    const size_t k = buf.find("//");
    if (k != std::string::npos) buf.resize(k);
    while (!buf.empty() && (buf.back() == ' ' || buf.back() == '\t' || buf.back() == '\r')) buf.pop_back();
    size_t b = 0;
    while (b < buf.size() && buf[b] == ' ') ++b;
    if (b == buf.size()) continue;
    line = buf.substr(b);
    return true;
  }
  std::fprintf(stderr, "At line 21 of file read_input_tools.f90: End of file\n");
  std::exit(2);
}
int split_line(std::string& line, std::string& out, const char* delim) {   // read_input_tools.f90:41-62
  const size_t i = line.find(delim);
  if (i == std::string::npos) { out = line; line.clear(); return 1; }
  out = line.substr(0, i); line = line.substr(i + 1);
  return 0;
}
std::vector<double> numbers(const std::string& s) {
  std::string t = s;
  for (char& c : t) if (c == ',') c = ' ';
  std::istringstream is(t);
  std::vector<double> v; double x;
  while (is >> x) v.push_back(x);
  return v;
}
template <class R> struct Api;
template <> struct Api<float> { static constexpr auto cal_coe = xee_cal_coe_f32; static constexpr auto solve = xee_solve_elliptic_old_f32; };
template <> struct Api<double> { static constexpr auto cal_coe = xee_cal_coe_f64; static constexpr auto solve = xee_solve_elliptic_old_f64; };

template <class R>
void read_field(const std::string& fn, std::vector<R>& f, size_t n) {   // field_tools.f90:30-52
  std::vector<float> raw(n, 0.f);
  FILE* fp = std::fopen(fn.c_str(), "rb");
  if (!fp) { std::printf(" Reading field error. File name: %s\n", fn.c_str()); std::fprintf(stderr, "xee_old_diagnose: cannot open %s\n", fn.c_str()); std::exit(2); }
  if (std::fread(raw.data(), 4, n, fp) != n) std::printf(" Reading field error. File name: %s\n", fn.c_str());
  std::fclose(fp);
  f.assign(raw.begin(), raw.end());
}
template <class R>
void write_field(const std::string& fn, const std::vector<R>& f, size_t n) {   // field_tools.f90:55-76
  std::vector<float> raw(f.begin(), f.begin() + n);
  FILE* fp = std::fopen(fn.c_str(), "wb");
  if (!fp || std::fwrite(raw.data(), 4, n, fp) != n) std::printf(" Writing field error. File name: %s\n", fn.c_str());
  if (fp) std::fclose(fp);
}
bool exists(const char* p) { struct stat st; return ::stat(p, &st) == 0; }

// Column-major 1-based accessor, as in the Fortran source.
#define AT(v, n1, i, j) (v)[((size_t)(i) - 1) + ((size_t)(j) - 1) * (size_t)(n1)]

template <class R>
int run() {
  const auto t_beg = std::chrono::steady_clock::now();
  const int debug_mode = exists("./debug_mode") ? 1 : 0;                    // :72
  std::string mode_str, word[4], buffer;
  read_input(std::cin, mode_str);
  for (int i = 0; i < 4; ++i) {
    if (split_line(mode_str, word[i], "-") != 0 && i != 3) { std::printf(" [INIT] Error(1): Mode number is not correct\n"); return 0; }
    std::printf(" READ:::%s\n", word[i].c_str());
  }
  int mode[4] = {0, 0, 0, 0};
  if (word[0] == "CYLINDRICAL") mode[0] = 0;
  else if (word[0] == "SPHERICAL") { std::printf(" [INIT] Error(1): SPHERICAL geometry is not provided by this re-host [D7]\n"); return 0; }
  else { std::printf(" [INIT] Error(1): Unknown Mode : [%s]\n", word[0].c_str()); return 0; }
  if (word[1] == "TENDENCY") mode[1] = 0; else if (word[1] == "INSTANT") mode[1] = 1;
  else { std::printf(" [INIT] Error(1): Unknown Mode : [%s]\n", word[1].c_str()); return 0; }
  if (word[2] == "DENSITY_NORMAL") mode[2] = 0; else if (word[2] == "DENSITY_BOUSSINESQ") mode[2] = 1;
  else { std::printf(" [INIT] Error(1): Unknown Mode : [%s]\n", word[2].c_str()); return 0; }
  if (word[3] == "BARO_ALL") mode[3] = 2; else if (word[3] == "BAROCLINIC") mode[3] = 1; else if (word[3] == "BAROTROPIC") mode[3] = 0;
  else { std::printf(" [INIT] Error(1): Unknown Mode : [%s]\n", word[3].c_str()); return 0; }
  R testing_dt = 0;
  if (mode[1] == 0) { read_input(std::cin, buffer); testing_dt = (R)numbers(buffer).at(0); }          // :126-128
  read_input(std::cin, buffer);
  const std::vector<double> dom = numbers(buffer);
  const R Lr[2] = {(R)dom.at(0), (R)dom.at(1)}, Lz[2] = {(R)dom.at(2), (R)dom.at(3)};
  read_input(std::cin, buffer);
  const int nr = (int)numbers(buffer).at(0), nz = (int)numbers(buffer).at(1);
  std::string input_folder, output_folder, A_file, B_file, C_file, Q_file, F_file, yes_or_no, rpsi_bc_file, rchi_bc_file;
  read_input(std::cin, input_folder); read_input(std::cin, output_folder);
  read_input(std::cin, A_file); read_input(std::cin, B_file); read_input(std::cin, C_file);
  read_input(std::cin, Q_file); read_input(std::cin, F_file);
  read_input(std::cin, buffer);
  std::vector<double> v = numbers(buffer);
  const int saved_strategy_rpsi = (int)v.at(0); const R saved_strategy_rpsi_r = (R)v.at(1); const int max_iter_rpsi = (int)v.at(2); const R alpha_rpsi = (R)v.at(3);
  read_input(std::cin, buffer);
  v = numbers(buffer);
  const int saved_strategy_rchi = (int)v.at(0); const R saved_strategy_rchi_r = (R)v.at(1); const int max_iter_rchi = (int)v.at(2); const R alpha_rchi = (R)v.at(3);
  bool use_rpsi_bc = false, use_rchi_bc = false;
  read_input(std::cin, yes_or_no);
  if (yes_or_no == "yes") { read_input(std::cin, rpsi_bc_file); use_rpsi_bc = true; }
  read_input(std::cin, yes_or_no);
  if (yes_or_no == "yes") { read_input(std::cin, rchi_bc_file); use_rchi_bc = true; }
  std::printf(" mode:  %d , %d , %d , %d\n", mode[0], mode[1], mode[2], mode[3]);
  if (mode[1] == 0) std::printf(" Testing time:   %.7E\n", (double)testing_dt);
  std::printf(" nr: %d , nz: %d\n", nr, nz);

  const size_t nO = (size_t)nr * nz, nA = (size_t)(nr - 1) * nz, nB = (size_t)(nr - 1) * (nz - 1), nC = (size_t)nr * (nz - 1);
  std::vector<R> rhoA_in, rhoB_in, rhoC_in, Q_in, F_in, rpsi_bc, rchi_bc;
  read_field(input_folder + "/" + A_file, rhoA_in, nO);
  read_field(input_folder + "/" + B_file, rhoB_in, nO);
  read_field(input_folder + "/" + C_file, rhoC_in, nO);
  read_field(input_folder + "/" + Q_file, Q_in, nB);                           // [D1]
  read_field(input_folder + "/" + F_file, F_in, nB);                           // [D1]
  if (use_rpsi_bc) read_field(input_folder + "/" + rpsi_bc_file, rpsi_bc, nO);
  if (use_rchi_bc) read_field(input_folder + "/" + rchi_bc_file, rchi_bc, nO);

  // constants.f90:4-5, geometry :256-273
  const R g0 = R(9.8), theta0 = R(298.0), Rd = R(287.0), Cv = R(5.0) / R(2.0) * Rd, Cp = Cv + Rd, kappa = Rd / Cp, h0 = Cp * theta0 / g0, p0 = R(101300.0);
  const R dr = (Lr[1] - Lr[0]) / R(nr - 1), dz = (Lz[1] - Lz[0]) / R(nz - 1);
  std::vector<R> ra(nr + 2), za(nz + 2), exner(nz + 2), rho(nz + 2), rcuva(nr + 2);   // 1-based below
  auto RA = [&](int i) -> R& { return ra[i]; }; auto ZA = [&](int j) -> R& { return za[j]; };
  auto RC = [&](int i) -> R& { return rcuva[i]; }; auto RHO = [&](int j) -> R& { return rho[j]; }; auto EX = [&](int j) -> R& { return exner[j]; };
  for (int i = 1; i <= nr; ++i) { RA(i) = Lr[0] + R(i - 1) * dr; RC(i) = RA(i); }
  for (int j = 1; j <= nz; ++j) {
    ZA(j) = Lz[0] + R(j - 1) * dz;
    EX(j) = mode[2] == 0 ? (R(1.0) - ZA(j) / h0) : R(1.0);
    RHO(j) = mode[2] == 0 ? p0 / (theta0 * Rd) * std::pow(EX(j), R(1.0) / kappa - R(1.0)) : R(1.0);
  }
  // integrals :1029-1113
  auto integrate_weight_B = [&](const std::vector<R>& w) {
    R s = R(0.0);
    for (int i = 1; i <= nr - 1; ++i)
      for (int j = 1; j <= nz - 1; ++j) {
        const R rcuv = (RC(i) + RC(i + 1)) / R(2.0), ddr = RA(i + 1) - RA(i), ddz = ZA(j + 1) - ZA(j), rho_ = (RHO(j + 1) + RHO(j)) / R(2.0);
        s = s + AT(w, nr - 1, i, j) * rho_ * rcuv * ddr * ddz;
      }
    return s;
  };
  auto cal_sum_Qeta = [&](const std::vector<R>& Q, const std::vector<R>& eta) {
    R s = R(0.0);
    for (int i = 1; i <= nr - 1; ++i)
      for (int j = 1; j <= nz - 1; ++j) {
        const R rcuv = (RC(i) + RC(i + 1)) / R(2.0), ddr = RA(i + 1) - RA(i), ddz = ZA(j + 1) - ZA(j), rho_ = (RHO(j + 1) + RHO(j)) / R(2.0);
        s = s + ((AT(eta, nr - 1, i, j) + AT(eta, nr - 1, i, j + 1)) / R(2.0)) * AT(Q, nr - 1, i, j) * rho_ * rcuv * ddr * ddz;
      }
    return s;
  };
  // FD operators :943-1027
  auto d_dz_B2A = [&](const std::vector<R>& from, std::vector<R>& to) {
    std::fill(to.begin(), to.end(), R(0));                                     // [D3]
    for (int i = 1; i <= nr - 1; ++i) for (int j = 2; j <= nz - 2; ++j)
      AT(to, nr - 1, i, j) = (AT(from, nr - 1, i, j) - AT(from, nr - 1, i, j - 1)) / ((ZA(j + 1) - ZA(j - 1)) / R(2.0));
  };
  auto d_dz_O2C = [&](const std::vector<R>& from, std::vector<R>& to) {
    for (int i = 1; i <= nr; ++i) for (int j = 1; j <= nz - 1; ++j) AT(to, nr, i, j) = (AT(from, nr, i, j + 1) - AT(from, nr, i, j)) / (ZA(j + 1) - ZA(j));
  };
  auto d_dr_B2B = [&](const std::vector<R>& from, std::vector<R>& to) {
    for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j) {
      int m, n;
      if (i == 1) { m = 0; n = 1; } else if (i == nr - 1) { m = -1; n = 0; } else { m = -1; n = 1; }
      AT(to, nr - 1, i, j) = (AT(from, nr - 1, i + m, j) - AT(from, nr - 1, i + n, j)) / (RA(i + m) - RA(i + n));
    }
  };
  auto d_dr_B2C = [&](const std::vector<R>& from, std::vector<R>& to) {
    std::fill(to.begin(), to.end(), R(0));
    for (int i = 2; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j)
      AT(to, nr, i, j) = (AT(from, nr - 1, i, j) - AT(from, nr - 1, i - 1, j)) / ((RA(i + 1) - RA(i - 1)) / R(2.0));
  };
  auto d_rcuvdr_O2A = [&](const std::vector<R>& from, std::vector<R>& to) {
    for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz; ++j) AT(to, nr - 1, i, j) = (AT(from, nr, i + 1, j) - AT(from, nr, i, j)) / (RA(i + 1) - RA(i));
    for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz; ++j) AT(to, nr - 1, i, j) = AT(to, nr - 1, i, j) / ((RC(i) + RC(i + 1)) / R(2.0));
  };
  auto rpsiToUW = [&](const std::vector<R>& rpsi, std::vector<R>& u, std::vector<R>& w) {   // :915-941
    d_rcuvdr_O2A(rpsi, w); d_dz_O2C(rpsi, u);
    for (auto& x : u) x = -x;
    for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz; ++j) AT(w, nr - 1, i, j) = AT(w, nr - 1, i, j) / RHO(j);
    for (int i = 1; i <= nr; ++i) for (int j = 1; j <= nz - 1; ++j) {
      if (RA(i) != 0) AT(u, nr, i, j) = AT(u, nr, i, j) / (RC(i) * (RHO(j) + RHO(j + 1)) / R(2.0)); else AT(u, nr, i, j) = R(0.0);
    }
  };
  auto cal_eta = [&](const std::vector<R>& rchi, std::vector<R>& eta) {                    // :1129-1141
    d_rcuvdr_O2A(rchi, eta);
    for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz; ++j) AT(eta, nr - 1, i, j) = AT(eta, nr - 1, i, j) * g0 / (RHO(j) * Cp * EX(j) * theta0);
  };
  auto relativeTheta = [&](std::vector<R>& th, const std::vector<R>& dth_dz_A, const std::vector<R>& dth_dr_C) {   // :893-912
    std::fill(th.begin(), th.end(), theta0);
    for (int i = 2; i <= nr - 1; ++i) AT(th, nr - 1, i, 1) = AT(th, nr - 1, i - 1, 1) + ((RA(i + 1) - RA(i - 1)) / R(2.0)) * AT(dth_dr_C, nr, i, 1);
    for (int i = 1; i <= nr - 1; ++i) for (int j = 2; j <= nz - 1; ++j)
      AT(th, nr - 1, i, j) = AT(th, nr - 1, i, j - 1) + ((ZA(j + 1) - ZA(j - 1)) / R(2.0)) * AT(dth_dz_A, nr - 1, i, j);
  };

  const R sum_Q = integrate_weight_B(Q_in);                                                // :284
  // a/b/c normalisation :289-312
  std::vector<R> solverA_A((size_t)(nr - 1) * (nz - 2)), solverB_B(nB), solverC_C((size_t)(nr - 2) * (nz - 1)), solver_b_basic_B, solver_b_anomaly_B(nB, R(0));
  for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 2; ++j)
    AT(solverA_A, nr - 1, i, j) = (AT(rhoA_in, nr, i, j + 1) + AT(rhoA_in, nr, i + 1, j + 1)) / (RC(i) + RC(i + 1)) / RHO(j + 1);
  for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j)
    AT(solverB_B, nr - 1, i, j) = (AT(rhoB_in, nr, i, j) + AT(rhoB_in, nr, i + 1, j) + AT(rhoB_in, nr, i, j + 1) + AT(rhoB_in, nr, i + 1, j + 1)) / (RC(i) + RC(i + 1)) / (RHO(j) + RHO(j + 1));
  solver_b_basic_B = solverB_B;
  for (int i = 1; i <= nr - 2; ++i) for (int j = 1; j <= nz - 1; ++j)
    AT(solverC_C, nr - 2, i, j) = (AT(rhoC_in, nr, i + 1, j) + AT(rhoC_in, nr, i + 1, j + 1)) / RC(i + 1) / (RHO(j) + RHO(j + 1));
  // staggered averages :329-354
  std::vector<R> rhoA_A(nA), rhoB_C(nC), rhoB_B(nB), rhoC_C(nC), b_basic_B, b_anomaly_B(nB, R(0));
  for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz; ++j) AT(rhoA_A, nr - 1, i, j) = (AT(rhoA_in, nr, i, j) + AT(rhoA_in, nr, i + 1, j)) / R(2.0);
  for (int i = 1; i <= nr; ++i) for (int j = 1; j <= nz - 1; ++j) AT(rhoB_C, nr, i, j) = (AT(rhoB_in, nr, i, j) + AT(rhoB_in, nr, i, j + 1)) / R(2.0);
  for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j)
    AT(rhoB_B, nr - 1, i, j) = (AT(rhoB_in, nr, i, j) + AT(rhoB_in, nr, i + 1, j) + AT(rhoB_in, nr, i, j + 1) + AT(rhoB_in, nr, i + 1, j + 1)) / R(4.0);
  b_basic_B = rhoB_B;
  for (int i = 1; i <= nr; ++i) for (int j = 1; j <= nz - 1; ++j) AT(rhoC_C, nr, i, j) = (AT(rhoC_in, nr, i, j) + AT(rhoC_in, nr, i, j + 1)) / R(2.0);
  // m2 :359-367  [D2]
  std::vector<R> m2(nB);
  for (int j = 1; j <= nz - 1; ++j) AT(m2, nr - 1, 1, j) = std::pow((RC(2) - RC(1)) / R(4.0), R(3.0)) * AT(rhoC_C, nr, 1, j) * (RA(2) - RA(1)) / R(2.0);
  for (int i = 2; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j)
    AT(m2, nr - 1, i, j) = AT(m2, nr - 1, i - 1, j) + std::pow(RC(i), R(3.0)) * AT(rhoC_C, nr, i, j) * (RA(i + 1) - RA(i - 1)) / R(2.0);
  // J :383-387
  std::vector<R> JJ_B(nB);
  for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j) AT(JJ_B, nr - 1, i, j) = AT(Q_in, nr - 1, i, j) / (Cp * EX(j));
  write_field(output_folder + "/J-B.bin", JJ_B, nB);
  write_field(output_folder + "/solver_a-sA.bin", solverA_A, solverA_A.size());
  write_field(output_folder + "/solver_b-B.bin", solverB_B, nB);
  write_field(output_folder + "/solver_c-sC.bin", solverC_C, solverC_C.size());
  // RHS thermal :396-406, momentum :412-436
  std::vector<R> RHS_thm(nO, R(0)), RHS_mom(nO, R(0)), wksp_O(nO), wksp_A(nA), wksp_B(nB), wksp_C(nC);
  d_dr_B2C(JJ_B, wksp_C);
  for (int i = 2; i <= nr - 1; ++i) for (int j = 2; j <= nz - 1; ++j) AT(RHS_thm, nr, i, j) = (AT(wksp_C, nr, i, j) + AT(wksp_C, nr, i, j - 1)) / R(2.0);
  for (auto& x : RHS_thm) x = x * g0 / theta0;
  write_field(output_folder + "/RHS_rpsi_thm-O.bin", RHS_thm, nO);
  for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j) AT(wksp_B, nr - 1, i, j) = std::sqrt(AT(m2, nr - 1, i, j)) * AT(F_in, nr - 1, i, j);
  d_dz_B2A(wksp_B, wksp_A);
  for (int i = 2; i <= nr - 1; ++i) for (int j = 2; j <= nz - 1; ++j) AT(RHS_mom, nr, i, j) = -(AT(wksp_A, nr - 1, i, j) + AT(wksp_A, nr - 1, i - 1, j)) / (RC(i) * RC(i));
  write_field(output_folder + "/RHS_rpsi_mom-O.bin", RHS_mom, nO);
  std::printf(" Initialization complete.\n");

  std::vector<R> coe(9 * nO, R(0)), rpsi(nO, R(0)), rchi(nO, R(0)), f(nO), u_C(nC), w_A(nA), theta(nB, theta0), eta(nA), wtheta_B(nB);
  int err = 0;
  const int dbg = debug_mode;
  auto solve = [&](std::vector<R>& field, int max_iter, int strategy_in, R strategy_r_in, R alpha) {
    int strategy = strategy_in; R strategy_r = strategy_r_in;
    std::fflush(stdout);
    Api<R>::solve(&max_iter, &strategy, &strategy_r, &alpha, field.data(), coe.data(), f.data(), wksp_O.data(), &nr, &nz, &err, &dbg);
    std::printf(" Relaxation uses  %11d  steps. Final residue is   %.7E .\n", strategy, (double)strategy_r);
  };
  auto set_operator = [&](bool with_B) {
    if (with_B) for (size_t q = 0; q < nB; ++q) solverB_B[q] = solver_b_basic_B[q] + solver_b_anomaly_B[q];
    else std::fill(solverB_B.begin(), solverB_B.end(), R(0));
    Api<R>::cal_coe(solverA_A.data(), solverB_B.data(), solverC_C.data(), coe.data(), &dr, &dz, &nr, &nz, &err);
  };
  R sum_dtheta_dt = 0;
  if (mode[1] == 0) {
    // ---- STAGE I :449-463
    Api<R>::cal_coe(solverA_A.data(), solverB_B.data(), solverC_C.data(), coe.data(), &dr, &dz, &nr, &nz, &err);
    std::fill(rpsi.begin(), rpsi.end(), R(0));
    if (use_rpsi_bc) rpsi = rpsi_bc;
    std::printf(" Solving rpsi...\n");
    for (size_t q = 0; q < nO; ++q) f[q] = RHS_thm[q] + RHS_mom[q];
    solve(rpsi, max_iter_rpsi, saved_strategy_rpsi, saved_strategy_rpsi_r, alpha_rpsi);
    write_field(output_folder + "/rpsi_before-O.bin", rpsi, nO);
    // ---- STAGE II :468-518
    rpsiToUW(rpsi, u_C, w_A);
    for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j)
      AT(theta, nr - 1, i, j) = AT(JJ_B, nr - 1, i, j)
          - theta0 / g0 * (AT(rhoA_A, nr - 1, i, j) * AT(w_A, nr - 1, i, j) + AT(rhoA_A, nr - 1, i, j + 1) * AT(w_A, nr - 1, i, j + 1)) / R(2.0)
          + theta0 / g0 * (AT(rhoB_C, nr, i, j) * AT(u_C, nr, i, j) + AT(rhoB_C, nr, i + 1, j) * AT(u_C, nr, i + 1, j)) / R(2.0);
    write_field(output_folder + "/w_before-A.bin", w_A, nA);
    write_field(output_folder + "/u_before-C.bin", u_C, nC);
    write_field(output_folder + "/dtheta_dt-B.bin", theta, nB);
    sum_dtheta_dt = integrate_weight_B(theta);
    for (auto& x : theta) x = x * testing_dt;
    d_dr_B2B(theta, wksp_B);
    for (size_t q = 0; q < nB; ++q) { b_anomaly_B[q] = -g0 / theta0 * wksp_B[q]; rhoB_B[q] = rhoB_B[q] + b_anomaly_B[q]; }
    d_dz_B2A(theta, wksp_A);
    for (int i = 1; i <= nr - 1; ++i) for (int j = 2; j <= nz - 1; ++j) AT(rhoA_A, nr - 1, i, j) = AT(rhoA_A, nr - 1, i, j) + g0 / theta0 * AT(wksp_A, nr - 1, i, j);
    for (int i = 2; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j) AT(rhoB_C, nr, i, j) = (AT(rhoB_B, nr - 1, i - 1, j) + AT(rhoB_B, nr - 1, i, j)) / R(2.0);
  } else {   // [D4]
    for (int i = 2; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j) AT(rhoB_C, nr, i, j) = (AT(rhoB_B, nr - 1, i - 1, j) + AT(rhoB_B, nr - 1, i, j)) / R(2.0);
  }
  {
    std::vector<R> tz(nA), tr(nC);
    for (size_t q = 0; q < nA; ++q) tz[q] = rhoA_A[q] * (theta0 / g0);
    for (size_t q = 0; q < nC; ++q) tr[q] = rhoB_C[q] * (-theta0 / g0);
    relativeTheta(theta, tz, tr);
  }
  write_field(output_folder + "/theta_after-B.bin", theta, nB);
  for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j)
    AT(solver_b_anomaly_B, nr - 1, i, j) = AT(b_anomaly_B, nr - 1, i, j) / ((RC(i) + RC(i + 1)) / R(2.0)) / ((RHO(j) + RHO(j + 1)) / R(2.0));
  // ---- STAGE III :524-673
  std::vector<R> f_basic(nO, R(0)), f_anomaly(nO, R(0));
  for (int i = 2; i <= nr - 1; ++i) for (int j = 2; j <= nz - 1; ++j) {
    AT(f_basic, nr, i, j) = -(AT(b_basic_B, nr - 1, i - 1, j - 1) + AT(b_basic_B, nr - 1, i - 1, j) + AT(b_basic_B, nr - 1, i, j) + AT(b_basic_B, nr - 1, i, j - 1)) / R(4.0);
    AT(f_anomaly, nr, i, j) = -(AT(b_anomaly_B, nr - 1, i - 1, j - 1) + AT(b_anomaly_B, nr - 1, i - 1, j) + AT(b_anomaly_B, nr - 1, i, j) + AT(b_anomaly_B, nr - 1, i, j - 1)) / R(4.0);
  }
  for (size_t q = 0; q < nO; ++q) f[q] = f_basic[q] + f_anomaly[q];
  write_field(output_folder + "/RHS_rchi-O.bin", f, nO);
  R sQ_0_0 = 0, sQ_B_0 = 0, sQ_0_dB = 0, sQ_B_dB = 0, sQ_0_B0 = 0, sQ_B_B0 = 0, sW_0 = 0, sW_B = 0, sBnd_0 = 0, sBnd_B = 0, sBnd2_0 = 0, sBnd2_B = 0;
  auto chi_solve = [&](bool with_B, const std::vector<R>* rhs, const char* tag, const char* msg, R& sum_out) {
    std::printf(" %s\n", msg);
    if (rhs) f = *rhs; else std::fill(f.begin(), f.end(), R(0));
    set_operator(with_B);
    solve(rchi, max_iter_rchi, saved_strategy_rchi, saved_strategy_rchi_r, alpha_rchi);
    cal_eta(rchi, eta);
    sum_out = cal_sum_Qeta(Q_in, eta);
    write_field(output_folder + "/eta-[" + tag + "]-A.bin", eta, nA);
    write_field(output_folder + "/rchi-[" + tag + "]-O.bin", rchi, nO);
  };
  const bool baro0 = mode[3] == 0 || mode[3] == 2, baro1 = mode[3] == 1 || mode[3] == 2;
  if (use_rchi_bc) {
    rchi = rchi_bc;
    if (baro0) chi_solve(false, nullptr, "0_0", "Solving CHI with L(A,B=0,C) = 0 with boundary condition", sQ_0_0);
    if (baro1) chi_solve(true, nullptr, "B0dB_0", "Solving CHI with L(A,B=B0+dB,C) = 0 with boundary condition", sQ_B_0);
  }
  std::fill(rchi.begin(), rchi.end(), R(0));
  if (baro0) chi_solve(false, &f_anomaly, "0_dB", "Solving CHI with L(A,B=0,C) = -dB", sQ_0_dB);
  if (baro1) chi_solve(true, &f_anomaly, "B0dB_dB", "Solving CHI with L(A,B=B0+dB,C) = -dB", sQ_B_dB);
  if (baro0) chi_solve(false, &f_basic, "0_B0", "Solving CHI with L(A,B=0,C) = -B0", sQ_0_B0);
  if (baro1) chi_solve(true, &f_basic, "B0dB_B0", "Solving CHI with L(A,B=B0+dB,C) = -B0", sQ_B_B0);
  // ---- Integral check :677-725
  std::printf(" Integral check...\n");
  std::fill(rpsi.begin(), rpsi.end(), R(0));
  if (use_rpsi_bc) rpsi = rpsi_bc;
  auto psi_check = [&](bool with_B, const char* tag, R& sum_out) {
    for (size_t q = 0; q < nO; ++q) f[q] = RHS_thm[q] + RHS_mom[q];
    set_operator(with_B);
    solve(rpsi, max_iter_rpsi, saved_strategy_rpsi, saved_strategy_rpsi_r, alpha_rpsi);
    rpsiToUW(rpsi, u_C, w_A);
    write_field(output_folder + "/rpsi_after-[" + tag + "]-O.bin", rpsi, nO);
    write_field(output_folder + "/w_after-[" + tag + "]-A.bin", w_A, nA);
    write_field(output_folder + "/u_after-[" + tag + "]-C.bin", u_C, nC);
    for (int i = 1; i <= nr - 1; ++i) for (int j = 1; j <= nz - 1; ++j)
      AT(wtheta_B, nr - 1, i, j) = ((AT(w_A, nr - 1, i, j) + AT(w_A, nr - 1, i, j + 1)) / R(2.0)) * AT(theta, nr - 1, i, j);   // :1117-1127
    sum_out = integrate_weight_B(wtheta_B) * (g0 / theta0);
    write_field(output_folder + "/wtheta_JF_after-[" + tag + "]-B.bin", wtheta_B, nB);
  };
  if (baro0) { std::printf(" Solving rpsi... L(A, B=0, C) = dJ/dr + dF/dz\n"); psi_check(false, "0", sW_0); }
  if (baro1) { std::printf(" Solving rpsi... L(A, B=B0dB, C) = dJ/dr + dF/dz\n"); psi_check(true, "B0dB", sW_B); }
  // ---- Exchange conversion :730-772, :1143-1174  [D5]
  auto exchange = [&](const std::vector<R>& psi_, const std::vector<R>& chi_, std::vector<R>& bnd, R& total) {
    total = R(0.0);
    const R ddz = ZA(2) - ZA(1), ddr = RA(2) - RA(1);
    for (int i = 1; i <= nr - 1; ++i) {
      const R r = (RA(i) + RA(i + 1)) / R(2.0);
      AT(bnd, nr - 1, i, 1) = ((AT(rhoC_in, nr, i, 1) + AT(rhoC_in, nr, i + 1, 1)) / (R(2.0) * RHO(1))) *
          (((AT(psi_, nr, i, 1) + AT(psi_, nr, i + 1, 1)) / R(2.0)) * ((AT(chi_, nr, i, 2) + AT(chi_, nr, i + 1, 2) - AT(chi_, nr, i, 1) - AT(chi_, nr, i + 1, 1)) / (R(2.0) * ddz)) -
           ((AT(chi_, nr, i, 1) + AT(chi_, nr, i + 1, 1)) / R(2.0)) * ((AT(psi_, nr, i, 2) + AT(psi_, nr, i + 1, 2) - AT(psi_, nr, i, 1) - AT(psi_, nr, i + 1, 1)) / (R(2.0) * ddz))) / (r * r);
      AT(bnd, nr - 1, i, 2) = ((AT(rhoC_in, nr, i, nz) + AT(rhoC_in, nr, i + 1, nz)) / (R(2.0) * RHO(nz))) *
          (((AT(psi_, nr, i, nz) + AT(psi_, nr, i + 1, nz)) / R(2.0)) * ((AT(chi_, nr, i, nz) + AT(chi_, nr, i + 1, nz) - AT(chi_, nr, i, nz - 1) - AT(chi_, nr, i + 1, nz - 1)) / (R(2.0) * ddz)) -
           ((AT(chi_, nr, i, nz) + AT(chi_, nr, i + 1, nz)) / R(2.0)) * ((AT(psi_, nr, i, nz) + AT(psi_, nr, i + 1, nz) - AT(psi_, nr, i, nz - 1) - AT(psi_, nr, i + 1, nz - 1)) / (R(2.0) * ddz))) / (r * r);
      total = total - (AT(bnd, nr - 1, i, 2) - AT(bnd, nr - 1, i, 1)) * r * ddr;
    }
  };
  if (use_rchi_bc) {
    std::printf(" Exchange conversion term check...\n");
    std::vector<R> psi_(nO), chi_(nO), tmp(nO), bnd((size_t)(nr - 1) * 2);
    auto add_from = [&](const std::string& fn) { read_field(fn, tmp, nO); for (size_t q = 0; q < nO; ++q) chi_[q] += tmp[q]; };
    if (baro0) {
      read_field(output_folder + "/rpsi_after-[0]-O.bin", psi_, nO); read_field(output_folder + "/rchi-[0_0]-O.bin", chi_, nO);
      add_from(output_folder + "/rchi-[0_dB]-O.bin"); add_from(output_folder + "/rchi-[0_B0]-O.bin");
      exchange(psi_, chi_, bnd, sBnd_0); write_field(output_folder + "/bndconv-[0].bin", bnd, bnd.size());
    }
    if (baro1) {
      read_field(output_folder + "/rpsi_after-[B0dB]-O.bin", psi_, nO); read_field(output_folder + "/rchi-[B0dB_0]-O.bin", chi_, nO);
      add_from(output_folder + "/rchi-[B0dB_dB]-O.bin"); add_from(output_folder + "/rchi-[B0dB_B0]-O.bin");
      exchange(psi_, chi_, bnd, sBnd_B); write_field(output_folder + "/bndconv-[B0dB].bin", bnd, bnd.size());
    }
    // Boundary conversion method 2 (:753-772): the same term without the boundary-condition chi; the reference prints the
    // same message again
    std::printf(" Exchange conversion term check...\n");
    if (baro0) {
      read_field(output_folder + "/rpsi_after-[0]-O.bin", psi_, nO);
      read_field(output_folder + "/rchi-[0_dB]-O.bin", chi_, nO); add_from(output_folder + "/rchi-[0_B0]-O.bin");
      exchange(psi_, chi_, bnd, sBnd2_0); write_field(output_folder + "/bndconv2-[0].bin", bnd, bnd.size());
    }
    if (baro1) {
      read_field(output_folder + "/rpsi_after-[B0dB]-O.bin", psi_, nO);
      read_field(output_folder + "/rchi-[B0dB_dB]-O.bin", chi_, nO); add_from(output_folder + "/rchi-[B0dB_B0]-O.bin");
      exchange(psi_, chi_, bnd, sBnd2_B); write_field(output_folder + "/bndconv2-[B0dB].bin", bnd, bnd.size());
    }
  }
  // ---- efficiency.txt :779-841 (list-directed: " label : value , ratio")
  const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_beg).count();
  FILE* fp = std::fopen((output_folder + "/efficiency.txt").c_str(), "w");   // [D6]
  if (!fp) { std::fprintf(stderr, "xee_old_diagnose: cannot write efficiency.txt\n"); return 2; }
  auto line1 = [&](const char* label, double v) { std::fprintf(fp, " %s%.8E\n", label, v); };
  auto line2 = [&](const char* label, double v) { std::fprintf(fp, " %s%.8E , %.8E\n", label, v, v / (double)sum_Q); };
  if (baro0) {
    line1("Time elapsed (sec)                          : ", elapsed);
    line1("sum Q                                       : ", (double)sum_Q);
    line1("sum dtheta_dt                               : ", (double)sum_dtheta_dt);
    line1("Local heat response (sum Q / sum dtheta_dt) : ", (double)(sum_dtheta_dt / sum_Q));
    std::fprintf(fp, " # Boundary efficiency\n");
    if (use_rchi_bc) line2("eta [L(B=0)    = 0]      w/  boundary : ", (double)sQ_0_0);
    std::fprintf(fp, " # Internal efficiency\n");
    line2("eta [L(B=0)    = dB]     wo/ boundary : ", (double)sQ_0_dB);
    line2("eta [L(B=0)    = B0]     wo/ boundary : ", (double)sQ_0_B0);
    if (use_rchi_bc) {
      std::fprintf(fp, " # Boundary conversion (Method 1)\n"); line2("bndconv [L(B=0) = B0dB]   w/ boundary : ", (double)sBnd_0);
      std::fprintf(fp, " # Boundary conversion (Method 2)\n"); line2("bndconv2 [L(B=0) = B0dB]   w/ boundary : ", (double)sBnd2_0);
    }
    std::fprintf(fp, " # Decomposition sum\n");
    R t = sQ_0_0 + sQ_0_dB + sQ_0_B0; if (use_rchi_bc) t = t + sBnd_0;
    line2("etaQ [L(B=0)    = J F] w/  boundary : ", (double)t);
    std::fprintf(fp, " # wtheta integral\n");
    line2("wtheta [L(B=0)    = J F] w/  boundary : ", (double)sW_0);
  }
  if (baro1) {
    std::fprintf(fp, " # Boundary efficiency\n");
    if (use_rchi_bc) line2("eta [L(B=B0dB) = 0]      w/  boundary : ", (double)sQ_B_0);
    std::fprintf(fp, " # Internal efficiency\n");
    line2("eta [L(B=B0dB) = dB]     wo/ boundary : ", (double)sQ_B_dB);
    line2("eta [L(B=B0dB) = B0]     wo/ boundary : ", (double)sQ_B_B0);
    if (use_rchi_bc) {
      std::fprintf(fp, " # Boundary conversion (Method 1)\n"); line2("bndconv [L(B=B0dB) = B0dB]w/ boundary : ", (double)sBnd_B);
      std::fprintf(fp, " # Boundary conversion (Method 2)\n"); line2("bndconv2 [L(B=B0dB) = B0dB]w/ boundary : ", (double)sBnd2_B);
    }
    std::fprintf(fp, " # Decomposition sum\n");
    R t = sQ_B_0 + sQ_B_dB + sQ_B_B0; if (use_rchi_bc) t = t + sBnd_B;
    line2("etaQ [L(B=B0dB) = J F] w/  boundary : ", (double)t);
    std::fprintf(fp, " # wtheta integral\n");
    line2("wtheta [L(B=B0dB) = J F] w/  boundary : ", (double)sW_B);
  }
  std::fclose(fp);
  std::printf(" Time elapsed (sec):   %.7E\n", elapsed);
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  bool r8 = false;
  for (int i = 1; i < argc; ++i) if (!std::strcmp(argv[i], "--r8")) r8 = true;
  return r8 ? run<double>() : run<float>();
}
