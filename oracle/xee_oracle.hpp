// ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
//
// CPU restatement of the XLab-EE-fortran elliptic hot path, used only as the
// parity checker (tests/, __graft_entry__.smoke()) and as the CPU baseline
// (bench.py cpu_baseline / --impl reference).  The product library
// (xlab_ee_fortran_b200/csrc) never includes or links this file.
//
// PARITY STATUS: "parity unpinned" by the reference's own tests.  The reference
// commits only the INPUTS of test/test1 (A,B,C,bc_init .bin + diag.txt) and no
// expected outputs, and no Fortran compiler exists in this image, so the real
// binary cannot be run.  This restatement is pinned instead against
//   (1) an independent vectorised numpy restatement (oracle/numpy_ref.py,
//       golden vectors in tests/golden/ made by tests/golden/make_golden.py),
//   (2) the surveyor-derived known-answer values in SURVEY.md section 6,
//   (3) structural invariants (coefficients sum to 0, flux-form equivalence).
//
// Every routine keeps the reference's operation order, loop order (do i / do j
// with j innermost), 1-based staggered indexing and pass structure.  Compile
// with -ffp-contract=off so no FMA contraction changes rounding (the reference
// build, make-diagnosis.sh:10-11, is plain gfortran -O0 on x86-64: no FMA).
//
// Templated on the real type: R=float is the reference's real(4); R=double is
// the same code with reals promoted (gfortran -freal-4-real-8 equivalent).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace xee_oracle {

// Fortran column-major 1-based accessors.
#define F2(p, n1, i, j) (p)[((size_t)(i) - 1) + ((size_t)(j) - 1) * (size_t)(n1)]
#define F3(p, n1, n2, k, i, j) \
  (p)[((size_t)(k) - 1) + ((size_t)(i) - 1) * (size_t)(n1) + ((size_t)(j) - 1) * (size_t)(n1) * (size_t)(n2)]

// xtt-lib-fortran/elliptic_tools.f90:3-4
enum { err_over_max_iteration = 1, err_explode = 2 };

// xtt-lib-fortran/constants.f90:4-5 (evaluated in R, as the promoted build would)
template <class R>
struct Constants {
  R g0, theta0, Rd, Cv, Cp, kappa, h0, p0;
  Constants() {
    g0 = R(9.8);
    theta0 = R(298.0);
    Rd = R(287.0);
    Cv = R(5.0) / R(2.0) * Rd;
    Cp = Cv + Rd;
    kappa = Rd / Cp;
    h0 = Cp * theta0 / g0;
    p0 = R(101300.0);
  }
};

// ---------------------------------------------------------------------------
// cal_coe — xtt-lib-fortran/elliptic_tools.f90:8-60
// a(nx-1,ny-2) b(nx-1,ny-1) c(nx-2,ny-1) -> coe(9,nx,ny); boundary entries of
// coe are never written.
template <class R>
void cal_coe(const R* a, const R* b, const R* c, R* coe, R dx, R dy, int nx, int ny, int* err) {
  R PP = dx * dx;            // dx**2                          :29
  R QQ = dy * dy;            //                                :30
  R PQ4 = R(4) * dx * dy;    // 4*dx*dy                        :31
  *err = 1;                  //                                :33
  for (int i = 2; i <= nx - 1; ++i) {      // :35
    for (int j = 2; j <= ny - 1; ++j) {    // :36
      R Ap = F2(a, nx - 1, i, j - 1) / PP;
      R Am = F2(a, nx - 1, i - 1, j - 1) / PP;
      R Cp = F2(c, nx - 2, i - 1, j) / QQ;
      R Cm = F2(c, nx - 2, i - 1, j - 1) / QQ;
      R BXp = (F2(b, nx - 1, i, j) + F2(b, nx - 1, i, j - 1)) / (R(2.0) * PQ4);
      R BXm = (F2(b, nx - 1, i - 1, j) + F2(b, nx - 1, i - 1, j - 1)) / (R(2.0) * PQ4);
      R BYp = (F2(b, nx - 1, i - 1, j) + F2(b, nx - 1, i, j)) / (R(2.0) * PQ4);
      R BYm = (F2(b, nx - 1, i - 1, j - 1) + F2(b, nx - 1, i, j - 1)) / (R(2.0) * PQ4);
      F3(coe, 9, nx, 1, i, j) = -(BXm + BYp);
      F3(coe, 9, nx, 2, i, j) = Cp + (BXp - BXm);
      F3(coe, 9, nx, 3, i, j) = BXp + BYp;
      F3(coe, 9, nx, 4, i, j) = Am - (BYp - BYm);
      F3(coe, 9, nx, 5, i, j) = -(Am + Ap + Cm + Cp);
      F3(coe, 9, nx, 6, i, j) = Ap + (BYp - BYm);
      F3(coe, 9, nx, 7, i, j) = BXm + BYm;
      F3(coe, 9, nx, 8, i, j) = Cm - (BXp - BXm);
      F3(coe, 9, nx, 9, i, j) = -(BXp + BYm);
    }
  }
  *err = 0;  // :58
}

// ---------------------------------------------------------------------------
// do_elliptic — elliptic_tools.f90:64-90.  Sum runs left to right in slot order.
// NB the reference sets err=1 and never resets it (:73); callers ignore it.
template <class R>
void do_elliptic(const R* psi, const R* coe, R* out, int nx, int ny, int* err) {
  *err = 1;
  for (int i = 2; i <= nx - 1; ++i) {
    for (int j = 2; j <= ny - 1; ++j) {
      R s = F3(coe, 9, nx, 1, i, j) * F2(psi, nx, i - 1, j + 1);
      s = s + F3(coe, 9, nx, 2, i, j) * F2(psi, nx, i, j + 1);
      s = s + F3(coe, 9, nx, 3, i, j) * F2(psi, nx, i + 1, j + 1);
      s = s + F3(coe, 9, nx, 4, i, j) * F2(psi, nx, i - 1, j);
      s = s + F3(coe, 9, nx, 5, i, j) * F2(psi, nx, i, j);
      s = s + F3(coe, 9, nx, 6, i, j) * F2(psi, nx, i + 1, j);
      s = s + F3(coe, 9, nx, 7, i, j) * F2(psi, nx, i - 1, j - 1);
      s = s + F3(coe, 9, nx, 8, i, j) * F2(psi, nx, i, j - 1);
      s = s + F3(coe, 9, nx, 9, i, j) * F2(psi, nx, i + 1, j - 1);
      F2(out, nx, i, j) = s;
    }
  }
}

// judge_error — elliptic_tools.f90:333-358 (list-directed prints start with a blank)
inline void judge_error(int err) {
  bool known = false;
  if (err == 0) { std::printf(" Elliptic Tools: Iteration success.\n"); known = true; }
  if (err & err_over_max_iteration) { std::printf(" Elliptic Tools: [Error] Max iteration reached.\n"); known = true; }
  if (err & err_explode) { std::printf(" Elliptic Tools: [Error] Iteration explodes.\n"); known = true; }
  if (!known) std::printf(" Elliptic Tools: Unknown error code %12d\n", err);
}

// Optional trace of the check sweeps (not in the reference; for tests only).
struct CheckTrace {
  int cap = 0, n = 0;
  int* iter = nullptr;
  double* err_now = nullptr;
  double* ratio = nullptr;
};

// ---------------------------------------------------------------------------
// solve_elliptic — elliptic_tools.f90:93-265.  Four separate passes per sweep,
// exactly as written.  Returns -1 (after printing the reference's message)
// where the reference would STOP (:126-129); otherwise 0.
//   quiet!=0 suppresses judge_error's print (test convenience only).
template <class R>
int solve_elliptic(int* max_iter, int check_step, int converge_time, int lost_rate, R* strategy_r1,
                   R* strategy_r2, R alpha, R* dat, const R* coe, const R* f, R* workspace, int nx,
                   int ny, int* err, int debug, CheckTrace* trace = nullptr, int quiet = 0) {
  bool check_abs_err, check_rel_err;
  if (*strategy_r1 > 0) check_abs_err = true;
  else { check_abs_err = false; *strategy_r1 = std::numeric_limits<R>::max(); }   // :112-117
  if (*strategy_r2 > 0) check_rel_err = true;
  else { check_rel_err = false; *strategy_r2 = std::numeric_limits<R>::max(); }   // :119-124
  if (!check_abs_err && !check_rel_err) {                                        // :126-129
    std::printf(" ERROR: [check_abs_err] and [check_rel_err] cannot both be non-positive.\n");
    return -1;
  }
  int check_step_use = 100;   if (check_step > 0) check_step_use = check_step;       // :131-134
  int converge_time_use = 10; if (converge_time > 0) converge_time_use = converge_time;
  int lost_rate_use = 5;      if (lost_rate > 0) lost_rate_use = lost_rate;          // :141-144

  if (debug == 1 || debug == 2) {  // :146-158
    std::printf(" ----- Solve Elliptic Inputs -----\n");
    std::printf("   max_iter       :  %11d\n", *max_iter);
    std::printf("   strategy_r1    :  %.8E\n", (double)*strategy_r1);
    std::printf("   strategy_r2    :  %.8E\n", (double)*strategy_r2);
    std::printf("   alpha          :  %.8E\n", (double)alpha);
    std::printf("   (nx, ny)       : ( %11d ,  %11d )\n", nx, ny);
    std::printf("   alpha          :  %.8E\n", (double)alpha);
    std::printf("   debug          :  %11d\n", debug);
    std::printf("   check step     :  %11d\n", check_step_use);
    std::printf("   converge time :  %11d\n", converge_time_use);
    std::printf(" ---------------------------------\n");
  }

  int converge_cnt = 0, lose_chance_cnt = 0;           // :160-161
  R err_before = std::numeric_limits<R>::max();        // :163
  *err = 0;                                            // :164
  // err_now / ratio are uninitialised in the reference until the first check;
  // the oracle defines them as 0 (only observable when max_iter < check_step).
  R err_now = 0, ratio = 0;

  const size_t nn = (size_t)nx * (size_t)ny;
  std::memcpy(workspace, dat, nn * sizeof(R));         // :166-171 (boundary copy then whole copy)

  R* fr_dat = workspace;                               // :173
  R* to_dat = dat;                                     // :174
  bool stop_iteration = false;
  const int max_iter_in = *max_iter;
  for (int cnt = 1; cnt <= max_iter_in; ++cnt) {       // :177
    bool flag = (cnt % check_step_use == 0);           // :179-183
    R* tmp = fr_dat; fr_dat = to_dat; to_dat = tmp;    // :185-187
    int tmp_err;
    do_elliptic(fr_dat, coe, to_dat, nx, ny, &tmp_err);                      // PASS 1 :189
    for (int j = 2; j <= ny - 1; ++j)                                        // PASS 2 :190 (array section)
      for (int i = 2; i <= nx - 1; ++i) F2(to_dat, nx, i, j) = F2(to_dat, nx, i, j) - F2(f, nx, i, j);

    if (flag) {                                                              // PASS 3 :192-234
      err_now = 0;
      for (int i = 2; i <= nx - 1; ++i)
        for (int j = 2; j <= ny - 1; ++j) {
          R v = F2(to_dat, nx, i, j);
          err_now = err_now + v * v;       // x**2.0 == x*x exactly (powf/pow(x,2) is exact-rounded)
        }
      err_now = std::sqrt(err_now / R((nx - 2) * (ny - 2)));                  // :199
      ratio = (err_before - err_now) / err_before;                           // :201
      if (debug == 2)
        std::printf("Iter: %8d, err_now: %12.3E, ratio: %12.3E\n", cnt, (double)err_now, (double)ratio);
      if (trace && trace->n < trace->cap) {
        trace->iter[trace->n] = cnt;
        trace->err_now[trace->n] = (double)err_now;
        trace->ratio[trace->n] = (double)ratio;
        trace->n++;
      }
      ratio = std::fabs(ratio);                                              // :205
      if (err_before == 0) {                                                 // :206
        stop_iteration = true;
        if (debug == 2) std::printf(" Error = 0, hardly to see this!\n");
      } else if ((err_now < *strategy_r1) && (ratio < *strategy_r2)) {       // :211
        converge_cnt = converge_cnt + 1;
        lose_chance_cnt = 0;
        if (debug == 2) std::printf(" converge_cnt:  %11d\n", converge_cnt);
        if (converge_cnt >= converge_time_use) stop_iteration = true;
      } else {
        if (converge_cnt > 0) {                                              // :221
          lose_chance_cnt = lose_chance_cnt + 1;
          if (lose_chance_cnt >= lost_rate_use) {
            converge_cnt = converge_cnt - 1;
            lose_chance_cnt = 0;
            if (debug == 2) std::printf(" Lose one count! converge_cnt now is:  %11d\n", converge_cnt);
          }
        }
      }
      err_before = err_now;                                                  // :233
    }

    for (int i = 2; i <= nx - 1; ++i)                                        // PASS 4 :236-240
      for (int j = 2; j <= ny - 1; ++j)
        F2(to_dat, nx, i, j) =
            F2(fr_dat, nx, i, j) + alpha * F2(to_dat, nx, i, j) / (-F3(coe, 9, nx, 5, i, j));

    if (cnt == max_iter_in) {                                                // :242-248
      stop_iteration = true;
      *err = *err | err_over_max_iteration;
      if (debug == 2) std::printf(" Max iteration reached. Exit iteration.\n");
    }
    if (stop_iteration) {                                                    // :249-256
      if (debug == 2) std::printf(" iter :  %11d , err_avg =   %.8E\n", cnt, (double)err_now);
      *max_iter = cnt; *strategy_r1 = err_now; *strategy_r2 = ratio;
      if (!quiet) judge_error(*err);
      break;
    }
  }
  if (to_dat != dat) {                                                       // :259-264
    if (debug == 2) std::printf(" to_dat is not associated with dat\n");
    std::memcpy(dat, workspace, nn * sizeof(R));
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Driver geometry — src/diagnose/initialize-variables.f90:45-67 (cylindrical;
// spherical restated as written, including cos() applied to degrees).
template <class R>
struct Geometry {
  int nr, nz;
  R dr, dz;
  std::vector<R> ra, za, exner, rho, rcuva, sin_table;
};

template <class R>
Geometry<R> make_geometry(R Lr1, R Lr2, R Lz1, R Lz2, int nr, int nz, int density_mode /*0 normal,1 boussinesq*/,
                          int geometry /*0 cyl, 1 sph*/, R planet_radius = R(0)) {
  Constants<R> k;
  Geometry<R> g;
  g.nr = nr; g.nz = nz;
  g.dr = (Lr2 - Lr1) / R(nr - 1);
  g.dz = (Lz2 - Lz1) / R(nz - 1);
  g.ra.resize(nr); g.za.resize(nz); g.exner.resize(nz); g.rho.resize(nz); g.rcuva.resize(nr);
  g.sin_table.assign(nr, R(0));
  for (int i = 1; i <= nr; ++i) g.ra[i - 1] = Lr1 + R(i - 1) * g.dr;
  for (int j = 1; j <= nz; ++j) {
    g.za[j - 1] = Lz1 + R(j - 1) * g.dz;
    g.exner[j - 1] = (density_mode == 0) ? (R(1.0) - g.za[j - 1] / k.h0) : R(1.0);
    g.rho[j - 1] = (density_mode == 0)
                       ? k.p0 / (k.theta0 * k.Rd) * std::pow(g.exner[j - 1], R(1.0) / k.kappa - R(1.0))
                       : R(1.0);
  }
  if (geometry == 0) {
    g.rcuva = g.ra;
  } else {
    R Lat1 = R(-90.0), Lat2 = R(90.0);
    R dlat = (Lat2 - Lat1) / R(nr - 1);
    for (int i = 1; i <= nr; ++i) {
      g.rcuva[i - 1] = planet_radius * std::cos(Lat1 + R(i - 1) * dlat);  // (sic) degrees into cos()
      g.sin_table[i - 1] = std::sin(Lat1 + R(i - 1) * dlat);
    }
  }
  return g;
}

// a/b/c normalisation — initialize-variables.f90:72-95
// inputs A,B,C on O(nr,nz); outputs a sA(nr-1,nz-2), b B(nr-1,nz-1), c sC(nr-2,nz-1)
template <class R>
void build_abc(const R* A, const R* B, const R* C, const Geometry<R>& g, R* a, R* b, R* c) {
  const int nr = g.nr, nz = g.nz;
  const R* rc = g.rcuva.data();
  const R* rho = g.rho.data();
  for (int i = 1; i <= nr - 1; ++i)
    for (int j = 1; j <= nz - 2; ++j)
      F2(a, nr - 1, i, j) = (F2(A, nr, i, j + 1) + F2(A, nr, i + 1, j + 1)) / (rc[i - 1] + rc[i]) / rho[j];
  for (int i = 1; i <= nr - 1; ++i)
    for (int j = 1; j <= nz - 1; ++j)
      F2(b, nr - 1, i, j) = (F2(B, nr, i, j) + F2(B, nr, i + 1, j) + F2(B, nr, i, j + 1) + F2(B, nr, i + 1, j + 1)) /
                            (rc[i - 1] + rc[i]) / (rho[j - 1] + rho[j]);
  for (int i = 1; i <= nr - 2; ++i)
    for (int j = 1; j <= nz - 1; ++j)
      F2(c, nr - 2, i, j) = (F2(C, nr, i + 1, j) + F2(C, nr, i + 1, j + 1)) / rc[i] / (rho[j - 1] + rho[j]);
}

// Staggered difference operators — src/diagnose/quick-tools2.f90
template <class R>  // :59-68  O(nr,nz) -> A(nr-1,nz)
void d_dr_O2A(const R* from, R* to, const Geometry<R>& g) {
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz; ++j)
      F2(to, g.nr - 1, i, j) = (F2(from, g.nr, i + 1, j) - F2(from, g.nr, i, j)) / (g.ra[i] - g.ra[i - 1]);
}
template <class R>  // :71-85
void d_rcuvdr_O2A(const R* from, R* to, const Geometry<R>& g) {
  d_dr_O2A(from, to, g);
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz; ++j)
      F2(to, g.nr - 1, i, j) = F2(to, g.nr - 1, i, j) / ((g.rcuva[i - 1] + g.rcuva[i]) / R(2.0));
}
template <class R>  // :16-25  O(nr,nz) -> C(nr,nz-1)
void d_dz_O2C(const R* from, R* to, const Geometry<R>& g) {
  for (int i = 1; i <= g.nr; ++i)
    for (int j = 1; j <= g.nz - 1; ++j)
      F2(to, g.nr, i, j) = (F2(from, g.nr, i, j + 1) - F2(from, g.nr, i, j)) / (g.za[j] - g.za[j - 1]);
}
template <class R>  // :1-13  B(nr-1,nz-1) -> A(nr-1,nz), rows 2..nz-2 only
void d_dz_B2A(const R* from, R* to, const Geometry<R>& g) {
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 2; j <= g.nz - 2; ++j)
      F2(to, g.nr - 1, i, j) =
          (F2(from, g.nr - 1, i, j) - F2(from, g.nr - 1, i, j - 1)) / ((g.za[j] - g.za[j - 2]) / R(2.0));
}
template <class R>  // :45-57  B(nr-1,nz-1) -> C(nr,nz-1), columns 2..nr-1 only
void d_dr_B2C(const R* from, R* to, const Geometry<R>& g) {
  for (int i = 2; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz - 1; ++j)
      F2(to, g.nr, i, j) =
          (F2(from, g.nr - 1, i, j) - F2(from, g.nr - 1, i - 1, j)) / ((g.ra[i] - g.ra[i - 2]) / R(2.0));
}
template <class R>  // :27-43  B -> B one-sided at the ends
void d_dr_B2B(const R* from, R* to, const Geometry<R>& g) {
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz - 1; ++j) {
      int m, n;
      if (i == 1) { m = 0; n = 1; }
      else if (i == g.nr - 1) { m = -1; n = 0; }
      else { m = -1; n = 1; }
      F2(to, g.nr - 1, i, j) =
          (F2(from, g.nr - 1, i + m, j) - F2(from, g.nr - 1, i + n, j)) / (g.ra[i + m - 1] - g.ra[i + n - 1]);
    }
}

// cal_eta — src/diagnose/quick-tools1.f90:1-13
template <class R>
void cal_eta(const R* rchi, R* eta, const Geometry<R>& g) {
  Constants<R> k;
  d_rcuvdr_O2A(rchi, eta, g);
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz; ++j)
      F2(eta, g.nr - 1, i, j) = F2(eta, g.nr - 1, i, j) * k.g0 / (g.rho[j - 1] * k.Cp * g.exner[j - 1] * k.theta0);
}

// cal_uw — src/diagnose/quick-tools1.f90:15-41 (== rpsiToUW, old-diagnose/diagnose.f90:915-941)
template <class R>
void cal_uw(const R* rpsi, R* u, R* w, const Geometry<R>& g) {
  d_rcuvdr_O2A(rpsi, w, g);
  d_dz_O2C(rpsi, u, g);
  for (size_t q = 0; q < (size_t)g.nr * (g.nz - 1); ++q) u[q] = -u[q];
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz; ++j) F2(w, g.nr - 1, i, j) = F2(w, g.nr - 1, i, j) / g.rho[j - 1];
  for (int i = 1; i <= g.nr; ++i)
    for (int j = 1; j <= g.nz - 1; ++j) {
      R r = g.ra[i - 1];
      if (r != 0) F2(u, g.nr, i, j) = F2(u, g.nr, i, j) / (g.rcuva[i - 1] * (g.rho[j - 1] + g.rho[j]) / R(2.0));
      else F2(u, g.nr, i, j) = R(0.0);
    }
}

// ---------------------------------------------------------------------------
// Legacy-driver kernels (src/old-diagnose/diagnose.f90).  The legacy code has
// latent bugs; each function restates the INTENDED maths
// and names the deviation.

// integrate_weight_B — :1029-1048 (== cal_sum_Q :1050-1071, cal_sum_wtheta :1094-1113)
template <class R>
R integrate_weight_B(const R* w, const Geometry<R>& g) {
  R s = R(0.0);
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz - 1; ++j) {
      R rcuv = (g.rcuva[i - 1] + g.rcuva[i]) / R(2.0);
      R dr = g.ra[i] - g.ra[i - 1];
      R dz = g.za[j] - g.za[j - 1];
      R rho_ = (g.rho[j] + g.rho[j - 1]) / R(2.0);
      s = s + F2(w, g.nr - 1, i, j) * rho_ * rcuv * dr * dz;
    }
  return s;
}
// cal_sum_Qeta — :1073-1092
template <class R>
R cal_sum_Qeta(const R* Q, const R* eta, const Geometry<R>& g) {
  R s = R(0.0);
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz - 1; ++j) {
      R rcuv = (g.rcuva[i - 1] + g.rcuva[i]) / R(2.0);
      R dr = g.ra[i] - g.ra[i - 1];
      R dz = g.za[j] - g.za[j - 1];
      R rho_ = (g.rho[j] + g.rho[j - 1]) / R(2.0);
      s = s + ((F2(eta, g.nr - 1, i, j) + F2(eta, g.nr - 1, i, j + 1)) / R(2.0)) * F2(Q, g.nr - 1, i, j) * rho_ *
                  rcuv * dr * dz;
    }
  return s;
}
// cal_wtheta — :1117-1127
template <class R>
void cal_wtheta(const R* w_A, const R* theta_B, R* wtheta_B, const Geometry<R>& g) {
  for (int i = 1; i <= g.nr - 1; ++i)
    for (int j = 1; j <= g.nz - 1; ++j)
      F2(wtheta_B, g.nr - 1, i, j) =
          ((F2(w_A, g.nr - 1, i, j) + F2(w_A, g.nr - 1, i, j + 1)) / R(2.0)) * F2(theta_B, g.nr - 1, i, j);
}
// Heating J and thermal RHS — :383-387, :396-406.
// DEVIATION: the legacy code allocates Q_in as (nr-1,nz-1) but reads it as an
// nr x nz record (:211,241); here Q is a genuine B-grid (nr-1,nz-1) field.
template <class R>
void rhs_thermal(const R* Q_B, R* JJ_B, R* rhs_O, const Geometry<R>& g) {
  Constants<R> k;
  const int nr = g.nr, nz = g.nz;
  for (int i = 1; i <= nr - 1; ++i)
    for (int j = 1; j <= nz - 1; ++j) F2(JJ_B, nr - 1, i, j) = F2(Q_B, nr - 1, i, j) / (k.Cp * g.exner[j - 1]);
  std::vector<R> wC((size_t)nr * (nz - 1), R(0));
  d_dr_B2C(JJ_B, wC.data(), g);
  for (size_t q = 0; q < (size_t)nr * nz; ++q) rhs_O[q] = R(0.0);
  for (int i = 2; i <= nr - 1; ++i)
    for (int j = 2; j <= nz - 1; ++j)
      F2(rhs_O, nr, i, j) = (F2(wC.data(), nr, i, j) + F2(wC.data(), nr, i, j - 1)) / R(2.0);
  for (size_t q = 0; q < (size_t)nr * nz; ++q) rhs_O[q] = rhs_O[q] * k.g0 / k.theta0;
}
// Dynamical (momentum) RHS — :412-436 given m2 on B and F on B.
template <class R>
void rhs_momentum(const R* m2_B, const R* F_B, R* rhs_O, const Geometry<R>& g) {
  const int nr = g.nr, nz = g.nz;
  std::vector<R> wB((size_t)(nr - 1) * (nz - 1)), wA((size_t)(nr - 1) * nz, R(0));
  for (int i = 1; i <= nr - 1; ++i)
    for (int j = 1; j <= nz - 1; ++j)
      F2(wB.data(), nr - 1, i, j) = std::sqrt(F2(m2_B, nr - 1, i, j)) * F2(F_B, nr - 1, i, j);
  d_dz_B2A(wB.data(), wA.data(), g);
  for (size_t q = 0; q < (size_t)nr * nz; ++q) rhs_O[q] = R(0.0);
  for (int i = 2; i <= nr - 1; ++i)
    for (int j = 2; j <= nz - 1; ++j)
      F2(rhs_O, nr, i, j) =
          -(F2(wA.data(), nr - 1, i, j) + F2(wA.data(), nr - 1, i - 1, j)) / (g.rcuva[i - 1] * g.rcuva[i - 1]);
}
// m2 from C — :359-367, cylindrical.
// DEVIATION: the legacy loop starts at i=1 and touches m2(0,j), ra(0) and uses
// stale i,j in the seed line; intended maths restated: seed column 1, then
// cumulative trapezoid for i=2..nr-1.
template <class R>
void angular_momentum_sq(const R* rhoC_C /*(nr,nz-1)*/, R* m2_B, const Geometry<R>& g) {
  const int nr = g.nr, nz = g.nz;
  for (int j = 1; j <= nz - 1; ++j) {
    R q = (g.rcuva[1] - g.rcuva[0]) / R(4.0);
    F2(m2_B, nr - 1, 1, j) = std::pow(q, R(3.0)) * F2(rhoC_C, nr, 1, j) * (g.ra[1] - g.ra[0]) / R(2.0);
  }
  for (int i = 2; i <= nr - 1; ++i)
    for (int j = 1; j <= nz - 1; ++j) {
      R rc = g.rcuva[i - 1];
      F2(m2_B, nr - 1, i, j) =
          F2(m2_B, nr - 1, i - 1, j) + std::pow(rc, R(3.0)) * F2(rhoC_C, nr, i, j) * (g.ra[i] - g.ra[i - 2]) / R(2.0);
    }
}
// f_basic — :524-530: negative 4-point average of a B-grid field to O interior.
template <class R>
void rhs_from_B(const R* b_B, R* f_O, const Geometry<R>& g) {
  const int nr = g.nr, nz = g.nz;
  for (size_t q = 0; q < (size_t)nr * nz; ++q) f_O[q] = R(0);
  for (int i = 2; i <= nr - 1; ++i)
    for (int j = 2; j <= nz - 1; ++j)
      F2(f_O, nr, i, j) = -(F2(b_B, nr - 1, i - 1, j - 1) + F2(b_B, nr - 1, i - 1, j) + F2(b_B, nr - 1, i, j) +
                            F2(b_B, nr - 1, i, j - 1)) / R(4.0);
}
// rhoA_A, rhoB_C, rhoB_B, rhoC_C — initialize-variables.f90:100-125
template <class R>
void stagger_averages(const R* A, const R* B, const R* C, R* rhoA_A, R* rhoB_C, R* rhoB_B, R* rhoC_C,
                      const Geometry<R>& g) {
  const int nr = g.nr, nz = g.nz;
  for (int i = 1; i <= nr - 1; ++i)
    for (int j = 1; j <= nz; ++j) F2(rhoA_A, nr - 1, i, j) = (F2(A, nr, i, j) + F2(A, nr, i + 1, j)) / R(2.0);
  for (int i = 1; i <= nr; ++i)
    for (int j = 1; j <= nz - 1; ++j) F2(rhoB_C, nr, i, j) = (F2(B, nr, i, j) + F2(B, nr, i, j + 1)) / R(2.0);
  for (int i = 1; i <= nr - 1; ++i)
    for (int j = 1; j <= nz - 1; ++j)
      F2(rhoB_B, nr - 1, i, j) =
          (F2(B, nr, i, j) + F2(B, nr, i + 1, j) + F2(B, nr, i, j + 1) + F2(B, nr, i + 1, j + 1)) / R(4.0);
  for (int i = 1; i <= nr; ++i)
    for (int j = 1; j <= nz - 1; ++j) F2(rhoC_C, nr, i, j) = (F2(C, nr, i, j) + F2(C, nr, i, j + 1)) / R(2.0);
}
// relativeTheta — old-diagnose/diagnose.f90:893-912
template <class R>
void relative_theta(R* theta_B, const R* dtheta_dz_A /*(nr-1,nz)*/, const R* dtheta_dr_C /*(nr,nz-1)*/,
                    const Geometry<R>& g) {
  Constants<R> k;
  const int nr = g.nr, nz = g.nz;
  for (size_t q = 0; q < (size_t)(nr - 1) * (nz - 1); ++q) theta_B[q] = k.theta0;
  // NB the reference loop runs i=2..nr-1 writing theta_B(i,1) with i up to nr-1 (in range).
  for (int i = 2; i <= nr - 1; ++i) {
    R dist = (g.ra[i] - g.ra[i - 2]) / R(2.0);
    F2(theta_B, nr - 1, i, 1) = F2(theta_B, nr - 1, i - 1, 1) + dist * F2(dtheta_dr_C, nr, i, 1);
  }
  for (int i = 1; i <= nr - 1; ++i)
    for (int j = 2; j <= nz - 1; ++j) {
      R dist = (g.za[j] - g.za[j - 2]) / R(2.0);
      F2(theta_B, nr - 1, i, j) = F2(theta_B, nr - 1, i, j - 1) + dist * F2(dtheta_dz_A, nr - 1, i, j);
    }
}

// ---------------------------------------------------------------------------
// cpu_fair — NOT a restatement: the same arithmetic (same per-point operation
// order, so the iterates are bit-identical to solve_elliptic's) but fused into
// one pass per sweep, contiguous (i innermost) loops and planar coefficients.
// Used only as the "fair" multi-core CPU throughput baseline.
// Runs exactly `sweeps` Jacobi sweeps, optional RMS residual of the LAST sweep.
template <class R>
void jacobi_sweeps_fused(R* x0, R* x1, const R* coe_planar /*[9][ny][nx]*/, const R* f, R alpha, int nx, int ny,
                         int sweeps, double* last_rms) {
  const size_t pl = (size_t)nx * ny;
  R* fr = x0; R* to = x1;
  for (int s = 0; s < sweeps; ++s) {
    double acc = 0;
    for (int j = 1; j < ny - 1; ++j) {
      const R* pm = fr + (size_t)(j - 1) * nx; const R* p0 = fr + (size_t)j * nx; const R* pp = fr + (size_t)(j + 1) * nx;
      const R* c = coe_planar + (size_t)j * nx;
      const R* ff = f + (size_t)j * nx;
      R* o = to + (size_t)j * nx;
      for (int i = 1; i < nx - 1; ++i) {
        R v = c[0 * pl + i] * pp[i - 1];
        v = v + c[1 * pl + i] * pp[i];
        v = v + c[2 * pl + i] * pp[i + 1];
        v = v + c[3 * pl + i] * p0[i - 1];
        v = v + c[4 * pl + i] * p0[i];
        v = v + c[5 * pl + i] * p0[i + 1];
        v = v + c[6 * pl + i] * pm[i - 1];
        v = v + c[7 * pl + i] * pm[i];
        v = v + c[8 * pl + i] * pm[i + 1];
        v = v - ff[i];
        if (last_rms && s == sweeps - 1) acc += (double)v * (double)v;
        o[i] = p0[i] + alpha * v / (-c[4 * pl + i]);
      }
    }
    if (last_rms && s == sweeps - 1) *last_rms = std::sqrt(acc / ((double)(nx - 2) * (ny - 2)));
    R* t = fr; fr = to; to = t;
  }
}

}  // namespace xee_oracle
