// xee_sweep_tb.cuh — v4 sweep kernel for sm_100a: TEMPORAL BLOCKING, up to TB sweeps of
// solve_elliptic (xtt-lib-fortran/elliptic_tools.f90:189-190, 193-199, 236-240) per pass over HBM.
//
// One pass takes the iterate pair (psi_k, psi_{k-1}) of every solve of a shared-operator batch to
// (psi_{k+t}, psi_{k+t-1}), t <= TB, reading psi_k, psi_{k-1}, f once and writing the two results once:
// 5 field transfers per t sweeps instead of 4 per sweep (Chebyshev) — the kernel is no longer bounded by
// the one-sweep-per-pass HBM roofline.  The arithmetic per point and sweep is the shared core of
// xee_common.cuh, so every iterate is bit-identical to the v1/v2/v3 kernels (and, in STRICT mode, to the reference's
// arithmetic); only the data movement differs:
//   * one persistent CTA of 384 threads per SM; a work unit = (tile of 64 x 30 grid points, chunk of solves);
//     tiles overlap by 2*TB points: after s sweeps the outer s rings of a tile are stale, so a tile owns the
//     points at distance >= TB from its edges — except along the domain boundary, whose Dirichlet values
//     never change, so nothing is lost there;
//   * thread = one column x 5 consecutive rows.  The 9 coefficients + 1/(-coe5) of its 5 points stay in
//     REGISTERS for the whole chunk (100 of its 168 registers) together with its own psi_k values; left/right/
//     up/down neighbours, f and psi_{k-1} are read from shared memory by [register + immediate] addresses
//     (5.2 loads + 1 store per point and sweep), the five FMA chains of a thread interleaved term by term;
//   * TMA (cp.async.bulk.tensor.3d) brings the three 64 x 30 boxes (psi_k, f, psi_{k-1}) of the next solves into a
//     4-stage ring (one mbarrier per stage) while the current one is being swept.  There are no separate work
//     tiles: a sweep reads its neighbours from one psi slot of the stage and writes the other one, whose old
//     content (psi_{k-1}) is needed by each thread at its own cells only; one named barrier per sweep; the last
//     sweep stores straight to global memory;
//   * input and output iterates live in DIFFERENT buffer pairs (tiles read their neighbours' cells, so a pass
//     cannot update in place); the host alternates the pairs pass by pass.
// Measured on B200 (512 solves, 512x256, fp64, Chebyshev FAST, TB = 4): 1.04 ms per pass = 261 us per sweep against
// 340-383 us for the one-sweep-per-pass TMA kernel; bounded by the per-warp fp64 dependency chains and the shared-
// memory pipe (58 % busy), not by HBM (2.8 GB per pass = 2.7 TB/s).
#pragma once
#include <cuda.h>

#include "xee_sweep_tma.cuh"

namespace xee {

namespace tb {
#ifndef XEE_TB_P        // tuning knobs of scripts/probe/tb_probe.cu; the library is built with the defaults
#define XEE_TB_P 5
#endif
#ifndef XEE_TB_G
#define XEE_TB_G 5
#endif
#ifndef XEE_TB_RG
#define XEE_TB_RG 6
#endif
#ifndef XEE_TB_NSTAGE
#define XEE_TB_NSTAGE 4
#endif
constexpr int W = 64;            // tile width  (grid points, i)
constexpr int P = XEE_TB_P;      // rows per thread
constexpr int G = XEE_TB_G;      // rows whose FMA chains are interleaved
constexpr int RG = XEE_TB_RG;    // row groups
constexpr int H = P * RG;        // tile height (grid points, j)
constexpr int NT = W * RG;       // 384 threads
constexpr int NSTAGE = XEE_TB_NSTAGE;
static_assert(P % G == 0, "P must be a multiple of G");
constexpr int TBMAX = 8;
template <class T> struct Cfg {
  static constexpr int ES = (int)sizeof(T);
  static constexpr int ROW_BYTES = W * ES;                       // 512 B (fp64)
  static constexpr int FLD_BYTES = W * H * ES;                   // 16 KB (fp64)
  static constexpr int STAGE_BYTES = 3 * FLD_BYTES;              // psi_k | f | psi_{k-1}
  // [pad row][stage 0]...[stage NSTAGE-1][pad row]: a thread reads its left/right/up/down neighbours at fixed offsets
  // from its own cell without clamping; at the tile edges that lands in a pad row or in the adjacent slot.  Whatever is
  // there only ever feeds stale-ring or boundary points, never an owned one.
  // No separate work tiles: a sweep reads its neighbours from one of the stage's psi slots and writes the other one,
  // whose old content (psi_{k-1}) each thread needs at its OWN cells only and has read by then.
  static constexpr int STAGE0 = ROW_BYTES + 128;                 // row -1, column -1 of the first slot stays inside
  static constexpr int SMEM_BYTES = STAGE0 + NSTAGE * STAGE_BYTES + ROW_BYTES + 128;   // row H, column W of the last
};
#ifdef XEE_TB_NOBAR   // probe only: upper bound of what removing the per-sweep barrier stalls could give (wrong results)
__device__ __forceinline__ void cta_bar_sync() {}
#else
__device__ __forceinline__ void cta_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
#endif
// Shared-memory accesses by 32-bit shared address: the byte offsets below are compile-time constants after unrolling,
// so every access is [register + immediate] and the inner loop carries four address registers in total.
template <class T> __device__ __forceinline__ T lds(uint32_t addr);
template <> __device__ __forceinline__ double lds<double>(uint32_t addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }
template <> __device__ __forceinline__ float lds<float>(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v)); }
__device__ __forceinline__ void sts(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }

// One sweep of a thread's P points (one column, rows r0..r0+P-1 of the tile), G rows at a time with their FMA chains
// interleaved term by term to cover the fp64 pipe latency.  sa/xa/fa/wa: shared addresses of the thread's first cell in
// the source tile (psi_k), the psi_{k-1} tile, the f tile and the destination work tile.
template <class T, int ARITH, bool CHEB, bool CHECK, bool LAST>
__device__ __forceinline__ void sweep_rows(const T (&cf)[P][9], const T (&rcp)[P], T (&x)[P], uint32_t sa, uint32_t xa,
                                           uint32_t fa, uint32_t wa, unsigned upd, unsigned own, T om, T alpha,
                                           T* __restrict__ on, T* __restrict__ op, int nxs, double& rr) {
  using R = Rn<T>;
  constexpr int ES = Cfg<T>::ES, RB = Cfg<T>::ROW_BYTES;
  T cprev = T(0);   // old value of the row just below the current group
#pragma unroll
  for (int g = 0; g < P / G; ++g) {
    T Lf[G + 2], Rt[G + 2], Cc[G + 2], fv[G], xm[G];
#pragma unroll
    for (int q = 0; q < G + 2; ++q) {
      Lf[q] = lds<T>(sa + (uint32_t)((g * G - 1 + q) * RB - ES));
      Rt[q] = lds<T>(sa + (uint32_t)((g * G - 1 + q) * RB + ES));
    }
    Cc[0] = g == 0 ? lds<T>(sa + (uint32_t)(-RB)) : cprev;
#pragma unroll
    for (int q = 0; q < G; ++q) Cc[q + 1] = x[g * G + q];
    Cc[G + 1] = g == P / G - 1 ? lds<T>(sa + (uint32_t)(P * RB)) : x[(g + 1) * G < P ? (g + 1) * G : 0];
#pragma unroll
    for (int q = 0; q < G; ++q) {
      fv[q] = lds<T>(fa + (uint32_t)((g * G + q) * RB));
      xm[q] = CHEB ? lds<T>(xa + (uint32_t)((g * G + q) * RB)) : T(0);
    }
    // apply9 (elliptic_tools.f90:77-85): slots 1..3 = row j+1, 4..6 = row j, 7..9 = row j-1, summed left to right
    T acc[G], out[G];
#pragma unroll
    for (int q = 0; q < G; ++q) acc[q] = R::mul(cf[g * G + q][0], Lf[q + 2]);
#pragma unroll
    for (int k = 1; k < 9; ++k) {
#pragma unroll
      for (int q = 0; q < G; ++q) {
        const int dq = 2 - k / 3;
        const T pv = (k % 3 == 0) ? Lf[q + dq] : (k % 3 == 1) ? Cc[q + dq] : Rt[q + dq];
        acc[q] = madd<T, ARITH>(acc[q], cf[g * G + q][k], pv);
      }
    }
#pragma unroll
    for (int q = 0; q < G; ++q) acc[q] = (ARITH == XEE_ARITH_STRICT) ? R::sub(acc[q], fv[q]) : acc[q] - fv[q];
    if (!CHEB) {
#pragma unroll
      for (int q = 0; q < G; ++q) out[q] = jacobi_update<T, ARITH>(Cc[q + 1], acc[q], alpha, cf[g * G + q][4], rcp[g * G + q]);
    } else {
#pragma unroll
      for (int q = 0; q < G; ++q) out[q] = jacobi_update<T, ARITH>(Cc[q + 1], acc[q], T(1), cf[g * G + q][4], rcp[g * G + q]);
#pragma unroll
      for (int q = 0; q < G; ++q) out[q] = out[q] - xm[q];
#pragma unroll
      for (int q = 0; q < G; ++q) out[q] = R::fma(om, out[q], xm[q]);
    }
    cprev = Cc[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const int r = g * G + q;
      const T o = ((upd >> r) & 1u) ? out[q] : Cc[q + 1];     // boundary / outside points keep their value
      x[r] = o;
      if (!LAST) {
        sts(wa + (uint32_t)(r * RB), o);
      } else if ((own >> r) & 1u) {
        on[r * nxs] = o; op[r * nxs] = Cc[q + 1];
        if (CHECK) rr += (double)acc[q] * (double)acc[q];
      }
    }
  }
}
}  // namespace tb

template <class T>
struct TbArgs {
  const T* coe;            // planar shared operator [10][ny][nx]
  T* out_new;              // psi_{k+t}    [n][ny][nx]
  T* out_prev;             // psi_{k+t-1}
  long long field_stride;  // nx*ny
  int nx, ny, nbatch;
  int nsweeps;             // t: sweeps in this pass, 1..TB
  int tbh;                 // halo depth the tiling was laid out for (TB)
  T alpha;                 // Jacobi weight
  T omega[tb::TBMAX];      // Chebyshev weights of the t sweeps
  const int* done;         // per-solve stop flags (NULL = none)
  double* partial;         // [n][ntiles] sum of r^2 of the LAST sweep of the pass (CHECK)
  int tiles_x, tiles_y, nchunks, chunk;
};

template <class T, int ARITH, int MODE, bool CHECK>
__global__ void __launch_bounds__(tb::NT, 1)
    sweep_tb_kernel(const __grid_constant__ TbArgs<T> a, const __grid_constant__ CUtensorMap map_x,
                    const __grid_constant__ CUtensorMap map_xm, const __grid_constant__ CUtensorMap map_f) {
  using namespace tb;
  using tma::mbar_init; using tma::mbar_expect_tx; using tma::mbar_wait; using tma::tma_load_3d;
  using C = Cfg<T>;
  constexpr bool CHEB = (MODE == MODE_CHEBYSHEV);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[NSTAGE];
  __shared__ double red[NT / 32];
  __shared__ uint32_t sdone[kDoneWords];   // stop flags of the batch as a bit mask: no global load per (tile, solve)

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const bool done_in_smem = stage_done_flags(a.done, a.nbatch, sdone, tid, NT);
  __syncthreads();
  auto is_done = [&](int n) -> bool {
    if (a.done == nullptr) return false;
    return done_in_smem ? ((sdone[n >> 5] >> (n & 31)) & 1u) != 0u : a.done[n] != 0;
  };

  const int ntiles = a.tiles_x * a.tiles_y;
  const int nunits = ntiles * a.nchunks;
  const size_t nn = (size_t)a.field_stride;
  const int step_x = W - 2 * a.tbh, step_y = H - 2 * a.tbh;

  // ---- prefetch cursor (thread 0): walks the same (unit, solve) sequence as the consumers, NSTAGE items ahead
  int pu = blockIdx.x, pn = -1;
  uint32_t issued = 0;
  auto issue_next = [&]() {   // issue the TMA loads of the next not-finished (unit, solve); false when exhausted
    for (;;) {
      if (pu >= nunits) return false;
      const int ch = pu / ntiles;
      const int n0 = ch * a.chunk, n1 = min(n0 + a.chunk, a.nbatch);
      if (pn < 0) pn = n0; else ++pn;
      if (pn >= n1) { pu += gridDim.x; pn = -1; continue; }
      if (is_done(pn)) continue;
      const int tile = pu % ntiles;
      const int ox = (tile % a.tiles_x) * step_x, oy = (tile / a.tiles_x) * step_y;
      const int s = issued % NSTAGE;
      unsigned char* st = smem_raw + C::STAGE0 + (size_t)s * C::STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], (uint32_t)((CHEB ? 3 : 2) * C::FLD_BYTES));
      tma_load_3d(st, &map_x, ox, oy, pn, &full_bar[s]);
      tma_load_3d(st + C::FLD_BYTES, &map_f, ox, oy, pn, &full_bar[s]);
      if (CHEB) tma_load_3d(st + 2 * C::FLD_BYTES, &map_xm, ox, oy, pn, &full_bar[s]);
      ++issued;
      return true;
    }
  };
  if (tid == 0)
    for (int q = 0; q < NSTAGE; ++q)
      if (!issue_next()) break;

  const int c = tid & (W - 1);
  const int r0 = (tid / W) * P;
  const uint32_t tofs = (uint32_t)((r0 * W + c) * C::ES);          // the thread's first cell inside a tile
  const uint32_t sm0 = tma::smem_u32(smem_raw) + tofs;
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t it = 0;      // consumed items

  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int tile = u % ntiles, ch = u / ntiles;
    const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
    const int ox = tx * step_x, oy = ty * step_y;
    const int n0 = ch * a.chunk, n1 = min(n0 + a.chunk, a.nbatch);
    const int gi = ox + c;
    // region of the tile that is exact after tbh sweeps and owned by it
    const int xlo = tx == 0 ? 1 : ox + a.tbh, xhi = tx == a.tiles_x - 1 ? a.nx - 1 : ox + W - a.tbh;
    const int ylo = ty == 0 ? 1 : oy + a.tbh, yhi = ty == a.tiles_y - 1 ? a.ny - 1 : oy + H - a.tbh;
    const bool col_int = gi >= 1 && gi < a.nx - 1;
    const bool col_own = gi >= xlo && gi < xhi;
    unsigned upd = 0, own = 0;    // bit r: point (gi, oy+r0+r) is a domain-interior point / is owned by this tile
    T cf[P][9], rcp[P];
#pragma unroll
    for (int r = 0; r < P; ++r) {
      const int gj = oy + r0 + r;
      const bool in = col_int && gj >= 1 && gj < a.ny - 1;
      if (in) upd |= 1u << r;
      if (in && col_own && gj >= ylo && gj < yhi) own |= 1u << r;
      const size_t off = in ? (size_t)gj * a.nx + gi : (size_t)a.nx + 1;
#pragma unroll
      for (int k = 0; k < 9; ++k) cf[r][k] = __ldg(a.coe + k * nn + off);
      rcp[r] = __ldg(a.coe + 9 * nn + off);
    }
    const long long gofs = (long long)(oy + r0) * a.nx + gi;      // the thread's first cell inside a field
    for (int n = n0; n < n1; ++n) {
      if (is_done(n)) continue;
      const uint32_t s = it % NSTAGE;
      mbar_wait(&full_bar[s], (it / NSTAGE) & 1);
      const uint32_t sb = sm0 + C::STAGE0 + s * C::STAGE_BYTES;    // psi_k | f | psi_{k-1} of this item
      uint32_t sa = sb;                       // neighbours of the current sweep: psi_k
      uint32_t xa = sb + 2 * C::FLD_BYTES;    // psi_{k-1} at the own cells; also where this sweep's result goes
      const uint32_t fa = sb + C::FLD_BYTES;  // f (read every sweep: the stage is held for the whole pass)
      T x[P];
#pragma unroll
      for (int r = 0; r < P; ++r) x[r] = lds<T>(sb + (uint32_t)(r * C::ROW_BYTES));
      const int nsw = a.nsweeps;
      double rr = 0.0;
      for (int sw = 0; sw < nsw; ++sw) {
        const T om = CHEB ? a.omega[sw] : T(1);
        if (sw < nsw - 1) {
          sweep_rows<T, ARITH, CHEB, CHECK, false>(cf, rcp, x, sa, xa, fa, xa, upd, own, om, a.alpha, nullptr, nullptr, 0, rr);
          cta_bar_sync();
          const uint32_t t = sa; sa = xa; xa = t;   // the slot just written is the next source, the old source the next psi_{k-1}
        } else {
          T* const on = a.out_new + ((size_t)n * nn + gofs);
          const long long dprev = a.out_prev - a.out_new;
          sweep_rows<T, ARITH, CHEB, CHECK, true>(cf, rcp, x, sa, xa, fa, xa, upd, own, om, a.alpha, on, on + dprev, a.nx, rr);
          if (sw == 0) cta_bar_sync();     // single-sweep pass
        }
        // Every thread has left the previous item (it is past this item's first barrier): that item's stage is free.
        if (sw == 0 && tid == 0 && it >= 1) issue_next();
      }
      if (CHECK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rr += __shfl_down_sync(0xffffffffu, rr, o);
        if (lane == 0) red[warp] = rr;
        cta_bar_sync();
        if (tid == 0) {
          double t = 0.0;
          for (int q = 0; q < NT / 32; ++q) t += red[q];
          a.partial[(size_t)n * ntiles + tile] = t;
        }
        cta_bar_sync();
      }
      ++it;
    }
  }
}

}  // namespace xee
