// ORACLE — TEST INFRASTRUCTURE ONLY (see xee_oracle.hpp header).
// extern "C" instantiations of the restatement for ctypes (oracle/oracle.py).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
#include "xee_oracle.hpp"

#include <atomic>
#include <chrono>
#include <thread>

using namespace xee_oracle;

namespace {
// Dynamic parallel-for over independent solves with plain std::thread (the image's default
// g++ wrapper has no libgomp spec, so OpenMP is avoided).
template <class Fn>
void parallel_for(int n, int threads, Fn fn) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  if (threads <= 1) { for (int s = 0; s < n; ++s) fn(s); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&]() { for (int s; (s = next.fetch_add(1)) < n;) fn(s); });
  for (auto& th : pool) th.join();
}
template <class R>
Geometry<R> geom(const double* dom, int nr, int nz, int density_mode, int geometry) {
  // dom = {Lr1, Lr2, Lz1, Lz2, planet_radius}; values pass through R as read(buffer,*) would store them.
  return make_geometry<R>(R(dom[0]), R(dom[1]), R(dom[2]), R(dom[3]), nr, nz, density_mode, geometry, R(dom[4]));
}
}  // namespace

#define XEE_INSTANTIATE(SFX, R)                                                                                    \
  extern "C" void xee_oracle_cal_coe_##SFX(const R* a, const R* b, const R* c, R* coe, R dx, R dy, int nx, int ny, \
                                           int* err) {                                                             \
    cal_coe<R>(a, b, c, coe, dx, dy, nx, ny, err);                                                                 \
  }                                                                                                                \
  extern "C" void xee_oracle_do_elliptic_##SFX(const R* psi, const R* coe, R* out, int nx, int ny, int* err) {     \
    do_elliptic<R>(psi, coe, out, nx, ny, err);                                                                    \
  }                                                                                                                \
  extern "C" int xee_oracle_solve_elliptic_##SFX(int* max_iter, int check_step, int converge_time, int lost_rate,  \
                                                 R* r1, R* r2, R alpha, R* dat, const R* coe, const R* f,          \
                                                 R* workspace, int nx, int ny, int* err, int debug, int quiet,     \
                                                 int trace_cap, int* trace_n, int* trace_iter,                     \
                                                 double* trace_err, double* trace_ratio) {                         \
    CheckTrace tr;                                                                                                 \
    tr.cap = trace_cap; tr.iter = trace_iter; tr.err_now = trace_err; tr.ratio = trace_ratio;                      \
    int rc = solve_elliptic<R>(max_iter, check_step, converge_time, lost_rate, r1, r2, alpha, dat, coe, f,        \
                               workspace, nx, ny, err, debug, trace_cap > 0 ? &tr : nullptr, quiet);               \
    if (trace_n) *trace_n = tr.n;                                                                                  \
    return rc;                                                                                                     \
  }                                                                                                                \
  /* Batch of independent literal solves, OpenMP over solves (each solve single-threaded, as the reference).  */  \
  /* coe_stride = 0 shares one coefficient array; arrays are [n][...] contiguous.  Returns wall seconds.      */  \
  extern "C" double xee_oracle_solve_batch_##SFX(int n, int* max_iter, int check_step, int converge_time,          \
                                                 int lost_rate, R* r1, R* r2, R alpha, R* dat, const R* coe,       \
                                                 long long coe_stride, const R* f, int nx, int ny, int* err,       \
                                                 int threads) {                                                    \
    const size_t nn = (size_t)nx * ny;                                                                             \
    auto t0 = std::chrono::steady_clock::now();                                                                    \
    parallel_for(n, threads, [&](int s) {                                                                          \
      std::vector<R> wk(nn);                                                                                       \
      solve_elliptic<R>(&max_iter[s], check_step, converge_time, lost_rate, &r1[s], &r2[s], alpha, dat + s * nn,   \
                        coe + (size_t)s * (size_t)coe_stride, f + s * nn, wk.data(), nx, ny, &err[s], 0, nullptr,  \
                        1);                                                                                        \
    });                                                                                                            \
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();                           \
  }                                                                                                                \
  /* cpu_fair: fused sweeps, planar coefficients, OpenMP over solves.  x is [n][ny][nx]; result left in x.   */   \
  extern "C" double xee_oracle_fair_batch_##SFX(int n, R* x, const R* coe_planar, long long coe_stride,            \
                                                const R* f, R alpha, int nx, int ny, int sweeps, double* rms,      \
                                                int threads) {                                                     \
    const size_t nn = (size_t)nx * ny;                                                                             \
    auto t0 = std::chrono::steady_clock::now();                                                                    \
    parallel_for(n, threads, [&](int s) {                                                                          \
      std::vector<R> x1(x + s * nn, x + (s + 1) * nn);                                                             \
      jacobi_sweeps_fused<R>(x + s * nn, x1.data(), coe_planar + (size_t)s * (size_t)coe_stride, f + s * nn,       \
                             alpha, nx, ny, sweeps, rms ? &rms[s] : nullptr);                                      \
      if (sweeps & 1) std::memcpy(x + s * nn, x1.data(), nn * sizeof(R));                                          \
    });                                                                                                            \
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();                           \
  }                                                                                                                \
  extern "C" void xee_oracle_geometry_##SFX(const double* dom, int nr, int nz, int density_mode, int geometry,     \
                                            R* dr, R* dz, R* ra, R* za, R* exner, R* rho, R* rcuva) {              \
    Geometry<R> g = geom<R>(dom, nr, nz, density_mode, geometry);                                                  \
    *dr = g.dr; *dz = g.dz;                                                                                        \
    std::memcpy(ra, g.ra.data(), nr * sizeof(R)); std::memcpy(za, g.za.data(), nz * sizeof(R));                    \
    std::memcpy(exner, g.exner.data(), nz * sizeof(R)); std::memcpy(rho, g.rho.data(), nz * sizeof(R));            \
    std::memcpy(rcuva, g.rcuva.data(), nr * sizeof(R));                                                            \
  }                                                                                                                \
  extern "C" void xee_oracle_build_abc_##SFX(const R* A, const R* B, const R* C, const double* dom, int nr,        \
                                             int nz, int density_mode, int geometry, R* a, R* b, R* c) {           \
    build_abc<R>(A, B, C, geom<R>(dom, nr, nz, density_mode, geometry), a, b, c);                                  \
  }                                                                                                                \
  extern "C" void xee_oracle_cal_eta_##SFX(const R* rchi, R* eta, const double* dom, int nr, int nz,               \
                                           int density_mode, int geometry) {                                       \
    cal_eta<R>(rchi, eta, geom<R>(dom, nr, nz, density_mode, geometry));                                           \
  }                                                                                                                \
  extern "C" void xee_oracle_cal_uw_##SFX(const R* rpsi, R* u, R* w, const double* dom, int nr, int nz,            \
                                          int density_mode, int geometry) {                                        \
    cal_uw<R>(rpsi, u, w, geom<R>(dom, nr, nz, density_mode, geometry));                                           \
  }                                                                                                                \
  extern "C" R xee_oracle_integrate_weight_B_##SFX(const R* w, const double* dom, int nr, int nz,                  \
                                                   int density_mode, int geometry) {                               \
    return integrate_weight_B<R>(w, geom<R>(dom, nr, nz, density_mode, geometry));                                 \
  }                                                                                                                \
  extern "C" R xee_oracle_cal_sum_Qeta_##SFX(const R* Q, const R* eta, const double* dom, int nr, int nz,          \
                                             int density_mode, int geometry) {                                     \
    return cal_sum_Qeta<R>(Q, eta, geom<R>(dom, nr, nz, density_mode, geometry));                                  \
  }                                                                                                                \
  extern "C" void xee_oracle_cal_wtheta_##SFX(const R* w_A, const R* theta_B, R* wtheta_B, const double* dom,      \
                                              int nr, int nz, int density_mode, int geometry) {                    \
    cal_wtheta<R>(w_A, theta_B, wtheta_B, geom<R>(dom, nr, nz, density_mode, geometry));                           \
  }                                                                                                                \
  extern "C" void xee_oracle_rhs_thermal_##SFX(const R* Q_B, R* JJ_B, R* rhs_O, const double* dom, int nr,         \
                                               int nz, int density_mode, int geometry) {                           \
    rhs_thermal<R>(Q_B, JJ_B, rhs_O, geom<R>(dom, nr, nz, density_mode, geometry));                                \
  }                                                                                                                \
  extern "C" void xee_oracle_rhs_momentum_##SFX(const R* m2_B, const R* F_B, R* rhs_O, const double* dom, int nr,  \
                                                int nz, int density_mode, int geometry) {                          \
    rhs_momentum<R>(m2_B, F_B, rhs_O, geom<R>(dom, nr, nz, density_mode, geometry));                               \
  }                                                                                                                \
  extern "C" void xee_oracle_angular_momentum_sq_##SFX(const R* rhoC_C, R* m2_B, const double* dom, int nr,        \
                                                       int nz, int density_mode, int geometry) {                   \
    angular_momentum_sq<R>(rhoC_C, m2_B, geom<R>(dom, nr, nz, density_mode, geometry));                            \
  }                                                                                                                \
  extern "C" void xee_oracle_rhs_from_B_##SFX(const R* b_B, R* f_O, const double* dom, int nr, int nz,             \
                                              int density_mode, int geometry) {                                    \
    rhs_from_B<R>(b_B, f_O, geom<R>(dom, nr, nz, density_mode, geometry));                                         \
  }                                                                                                                \
  extern "C" void xee_oracle_stagger_averages_##SFX(const R* A, const R* B, const R* C, R* rhoA_A, R* rhoB_C,      \
                                                    R* rhoB_B, R* rhoC_C, const double* dom, int nr, int nz,       \
                                                    int density_mode, int geometry) {                              \
    stagger_averages<R>(A, B, C, rhoA_A, rhoB_C, rhoB_B, rhoC_C, geom<R>(dom, nr, nz, density_mode, geometry));    \
  }                                                                                                                \
  extern "C" void xee_oracle_relative_theta_##SFX(R* theta_B, const R* dtheta_dz_A, const R* dtheta_dr_C,          \
                                                  const double* dom, int nr, int nz, int density_mode,             \
                                                  int geometry) {                                                  \
    relative_theta<R>(theta_B, dtheta_dz_A, dtheta_dr_C, geom<R>(dom, nr, nz, density_mode, geometry));            \
  }

XEE_INSTANTIATE(f32, float)
XEE_INSTANTIATE(f64, double)

extern "C" void xee_oracle_judge_error(int err) { judge_error(err); std::fflush(stdout); }

extern "C" int xee_oracle_max_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}
