// xee_sweep_tma.cuh — v2 sweep kernel for sm_100a: persistent, warp-specialised, TMA-fed.
//
// Same arithmetic as sweep_direct_kernel (K3/K4: xtt-lib-fortran/elliptic_tools.f90:189-190, 193-199,
// 236-240 fused into one pass), different data movement:
//   * one persistent CTA per SM; a work unit = (tile of TW x TH interior points, chunk of solves);
//   * warp 16 (one elected lane) is the PRODUCER: for every solve of the unit it issues
//     cp.async.bulk.tensor (TMA) loads of the psi tile with its one-point halo, the f tile and (Chebyshev)
//     the psi_{k-1} tile into a ring of NSTAGE shared-memory stages, completion on an mbarrier per stage;
//   * warps 0..15 are CONSUMERS: thread = one column x RPT rows of the tile; the 9 coefficients and
//     1/(-coe5) of its points stay in REGISTERS for the whole chunk of solves (the operator is shared
//     by the batch), psi neighbours come from shared memory, results go straight to global (coalesced);
//   * out-of-range halo/tile elements are zero-filled by TMA; stores are predicated.
// The operator planes are read once per unit through L2; psi / f / psi' stream from HBM exactly once per
// sweep (+ (TH+2)/TH halo rows, served by L2 because vertically adjacent tiles run concurrently).
#pragma once
#include <cuda.h>

#include "xee_kernels.cuh"

namespace xee {

namespace tma {
constexpr int TW = 128;       // tile width  (interior points, i)
constexpr int TH = 8;         // tile height (interior points, j)
constexpr int RPT = 2;        // rows per consumer thread
constexpr int NCONS = TW * (TH / RPT);   // 512 consumer threads
constexpr int NTHREADS = NCONS + 32;     // + one producer warp
constexpr int NSTAGE_MAX = 8;

template <class T> struct Cfg {
  static constexpr int HALO_W = TW + (int)(16 / sizeof(T));   // halo row pitch: TW+2 rounded up to 16 bytes
  static constexpr int PSI_BYTES = ((TH + 2) * HALO_W * (int)sizeof(T) + 127) / 128 * 128;
  // f / psi_{k-1} tiles also start at column i0-1 and are HALO_W wide: on B200 a TMA box whose first element is
  // not 16-byte aligned in global memory (inner coordinate * sizeof(T) % 16 != 0) faults with "illegal
  // instruction" (measured, scripts/probe/tma_probe.cu), and i0 = 1 + k*TW is odd.
  static constexpr int FLD_RAW = TH * HALO_W * (int)sizeof(T);
  static constexpr int FLD_BYTES = (FLD_RAW + 127) / 128 * 128;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "XEE_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra XEE_DONE_%=;\n"
      "bra XEE_WAIT_%=;\n"
      "XEE_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }
}  // namespace tma

// Per-solve stop flags -> bit mask in shared memory (warp ballots), so that the persistent kernels test a flag per
// (tile, solve) without a dependent global load.  Returns false (nothing staged) when there are no flags or the batch
// is larger than the mask; the caller must __syncthreads() before reading.
constexpr int kDoneWords = 512;
__device__ __forceinline__ bool stage_done_flags(const int* done, int nbatch, uint32_t* sdone, int tid, int nthreads) {
  if (done == nullptr || nbatch > kDoneWords * 32) return false;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
  for (int w = warp; w * 32 < nbatch; w += nwarps) {
    const int n = w * 32 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, n < nbatch ? done[n] != 0 : true);
    if (lane == 0) sdone[w] = m;
  }
  return true;
}

template <class T>
struct TmaSweepArgs {
  SweepArgs<T> a;
  int tiles_x, tiles_y, nchunks, chunk;   // work units = tiles_x*tiles_y*nchunks; `chunk` solves per unit
  int nstage;
};

// PERSOLVE = one operator per solve (time series): the 10 operator planes of the tile arrive through the same TMA
// stage as psi and f (one 4-D box [1][10][TH][HALO_W]) and are read from shared memory instead of registers.
template <class T, int ARITH, int MODE, bool CHECK, bool PERSOLVE>
__global__ void __launch_bounds__(tma::NTHREADS, 1)
    sweep_tma_kernel(const TmaSweepArgs<T> P, const __grid_constant__ CUtensorMap map_src,
                     const __grid_constant__ CUtensorMap map_prev, const __grid_constant__ CUtensorMap map_f,
                     const __grid_constant__ CUtensorMap map_coe) {
  using namespace tma;
  using R = Rn<T>;
  using C = Cfg<T>;
  constexpr bool CHEB = (MODE == MODE_CHEBYSHEV);
  constexpr int COE_OFF = C::PSI_BYTES + C::FLD_BYTES + (CHEB ? C::FLD_BYTES : 0);
  constexpr int COE_RAW = kPlanes * C::FLD_RAW;
  constexpr int STAGE_BYTES = COE_OFF + (PERSOLVE ? (COE_RAW + 127) / 128 * 128 : 0);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[NSTAGE_MAX], empty_bar[NSTAGE_MAX];
  __shared__ double red[NCONS / 32];
  __shared__ uint32_t sdone[kDoneWords];   // stop flags of the batch as a bit mask: no global load per (tile, solve)

  const SweepArgs<T>& a = P.a;
  const int tid = threadIdx.x;
  const int nstage = P.nstage;
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NCONS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const bool done_in_smem = stage_done_flags(a.done, a.nbatch, sdone, tid, NTHREADS);
  __syncthreads();
  auto is_done = [&](int n) -> bool {
    if (a.done == nullptr) return false;
    return done_in_smem ? ((sdone[n >> 5] >> (n & 31)) & 1u) != 0u : a.done[n] != 0;
  };

  const int ntiles = P.tiles_x * P.tiles_y;
  const int nunits = ntiles * P.nchunks;
  const size_t nn = (size_t)a.field_stride;
  uint32_t it = 0;   // running stage counter (same sequence in producer and consumers)

  if (tid >= NCONS) {
    // ------------------------------------------------------------------ producer warp
    if (tid == NCONS) {
      for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int tile = u % ntiles, ch = u / ntiles;
        const int i0 = 1 + (tile % P.tiles_x) * TW, j0 = 1 + (tile / P.tiles_x) * TH;
        const int n0 = ch * P.chunk, n1 = min(n0 + P.chunk, a.nbatch);
        for (int n = n0; n < n1; ++n) {
          if (is_done(n)) continue;
          const int s = it % nstage;
          const uint32_t ph = (it / nstage) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* st = smem_raw + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], (uint32_t)((TH + 2) * C::HALO_W * sizeof(T) + C::FLD_RAW + (CHEB ? C::FLD_RAW : 0) + (PERSOLVE ? COE_RAW : 0)));
          if (PERSOLVE) tma_load_4d(st + COE_OFF, &map_coe, i0 - 1, j0, 0, n, &full_bar[s]);
          tma_load_3d(st, &map_src, i0 - 1, j0 - 1, n, &full_bar[s]);
          tma_load_3d(st + C::PSI_BYTES, &map_f, i0 - 1, j0, n, &full_bar[s]);
          if (CHEB) tma_load_3d(st + C::PSI_BYTES + C::FLD_BYTES, &map_prev, i0 - 1, j0, n, &full_bar[s]);
          ++it;
        }
      }
    }
    return;
  }
  // -------------------------------------------------------------------- consumers
  const int col = tid % TW;
  const int rg = tid / TW;            // row group: rows rg*RPT .. rg*RPT+RPT-1 of the tile
  const int lane = tid & 31, warp = tid >> 5;
  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int tile = u % ntiles, ch = u / ntiles;
    const int i0 = 1 + (tile % P.tiles_x) * TW, j0 = 1 + (tile / P.tiles_x) * TH;
    const int n0 = ch * P.chunk, n1 = min(n0 + P.chunk, a.nbatch);
    const int gi = i0 + col;
    bool valid[RPT];
    size_t off[RPT];
    T c[RPT][9], rcp[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int gj = j0 + rg * RPT + r;
      valid[r] = (gi < a.nx - 1) && (gj < a.ny - 1);
      off[r] = valid[r] ? (size_t)gj * a.nx + gi : (size_t)a.nx + 1;
      if (!PERSOLVE) {
#pragma unroll
        for (int k = 0; k < 9; ++k) c[r][k] = __ldg(a.coe + k * nn + off[r]);
        rcp[r] = __ldg(a.coe + 9 * nn + off[r]);
      }
    }
    for (int n = n0; n < n1; ++n) {
      if (is_done(n)) continue;
      const int s = it % nstage;
      const uint32_t ph = (it / nstage) & 1;
      mbar_wait(&full_bar[s], ph);
      const unsigned char* st = smem_raw + (size_t)s * STAGE_BYTES;
      const T* sp = reinterpret_cast<const T*>(st);                           // [(TH+2)][HALO_W], origin (i0-1, j0-1)
      const T* sf = reinterpret_cast<const T*>(st + C::PSI_BYTES);            // [TH][HALO_W], origin (i0-1, j0)
      const T* sv = reinterpret_cast<const T*>(st + C::PSI_BYTES + C::FLD_BYTES);
      // rows rg*RPT-1 .. rg*RPT+RPT of the tile = smem rows rg*RPT .. rg*RPT+RPT+1
      T w[RPT + 2][3];
#pragma unroll
      for (int q = 0; q < RPT + 2; ++q) {
        const T* row = sp + (size_t)(rg * RPT + q) * C::HALO_W + col;
        w[q][0] = row[0]; w[q][1] = row[1]; w[q][2] = row[2];
      }
      T fv[RPT], xm[RPT];
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        fv[r] = sf[(rg * RPT + r) * C::HALO_W + col + 1];
        xm[r] = CHEB ? sv[(rg * RPT + r) * C::HALO_W + col + 1] : T(0);
        if (PERSOLVE) {
          const T* sc = reinterpret_cast<const T*>(st + COE_OFF) + (rg * RPT + r) * C::HALO_W + col + 1;   // [10][TH][HALO_W]
#pragma unroll
          for (int k = 0; k < 9; ++k) c[r][k] = sc[k * TH * C::HALO_W];
          rcp[r] = sc[9 * TH * C::HALO_W];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);     // stage consumed: everything is in registers now
      ++it;
      double rr = 0.0;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const T p[9] = {w[r + 2][0], w[r + 2][1], w[r + 2][2], w[r + 1][0], w[r + 1][1], w[r + 1][2], w[r][0], w[r][1], w[r][2]};
        T res = apply9<T, ARITH>(c[r], p);
        res = (ARITH == XEE_ARITH_STRICT) ? R::sub(res, fv[r]) : res - fv[r];
        T out;
        if (!CHEB) {
          out = jacobi_update<T, ARITH>(p[4], res, a.alpha, c[r][4], rcp[r]);
        } else {
          const T om = (PERSOLVE && a.rho_ps) ? (T)cheb_omega(a.cheb_k, (double)a.rho_ps[n]) : a.omega;
          const T xj = jacobi_update<T, ARITH>(p[4], res, T(1), c[r][4], rcp[r]);
          out = R::fma(om, xj - xm[r], xm[r]);
        }
        if (valid[r]) {
          a.dst[(size_t)n * nn + off[r]] = out;
          if (CHECK) rr += (double)res * (double)res;
        }
      }
      if (CHECK) {   // deterministic reduction over the 512 consumers of this (tile, solve)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rr += __shfl_down_sync(0xffffffffu, rr, o);
        if (lane == 0) red[warp] = rr;
        consumer_bar_sync();
        if (tid == 0) {
          double t = 0.0;
          for (int q = 0; q < NCONS / 32; ++q) t += red[q];
          a.partial[(size_t)n * a.ntiles + tile] = t;
        }
        consumer_bar_sync();
      }
    }
  }
}

}  // namespace xee
