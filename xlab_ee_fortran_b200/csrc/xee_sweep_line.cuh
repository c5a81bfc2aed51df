// xee_sweep_line.cuh — v5 sweep kernel for sm_100a: SEGMENT-LINE relaxation along the radius.
//
// Same discrete problem, same residual r = L psi - f (do_elliptic's nine-term sum, xtt-lib-fortran/
// elliptic_tools.f90:77-85) and the same stop rule (:193-233) as solve_elliptic, but the correction is not the
// reference's point-wise r / (-coe5) (:238): the radial line is cut into segments of 8 points and every segment's
// tridiagonal system  coe4 z(i-1) + coe5 z(i) + coe6 z(i+1) = r(i)  is solved exactly (Thomas algorithm, factors
// precomputed once per operator), psi' = psi - alpha z.  On the secondary-circulation operator the radial coupling
// carries ~97 % of the diagonal (dr >> dz after the 1/(rho r) scaling), so this block-Jacobi splitting has a 7x
// larger spectral gap than point Jacobi and its Chebyshev acceleration needs ~2.6x fewer sweeps for the same
// residual tolerance.  XEE_METHOD_LINE_JACOBI / XEE_METHOD_LINE_CHEBYSHEV; FAST arithmetic, shared operator.
//
// Data movement (one sweep per pass, HBM-bound like the v2 kernel):
//   * thread = 8 consecutive radial points of one row: the whole Thomas solve runs in its registers, with the
//     9 coefficients of its points (144 registers) kept for the whole chunk of solves, the 2 Thomas factors in
//     thread-private shared-memory cells;
//   * warp = one segment column x 32 rows, lane = row.  TMA boxes are 70 (psi, with halo) and 66 (f, psi_{k-1})
//     elements wide, i.e. an ODD number of 16-byte chunks per shared-memory row, so the 128-bit loads of the 32 lanes
//     of a warp (same columns, consecutive rows) hit distinct banks without any swizzle;
//   * one persistent CTA of 256 threads per SM, tiles of 64 x 32 points, 3-stage TMA ring, one named barrier per
//     (tile, solve) to hand the stage back; results go to global memory as 128-bit stores;
//   * the operator and the factors are repacked once per operator in tile/thread order (line_pack_kernel), so the
//     per-unit reload of a thread's 88 constants is 44 fully coalesced 128-bit loads (2-3 us per unit instead of 10).
// Measured on B200 (512 solves, 512x256, fp64, Chebyshev): 377 us per sweep = 5.66 TB/s algorithmic (87 % of the measured
// copy peak); DRAM traffic 2.20 GB per sweep for 2.13 GB algorithmic.
#pragma once
#include <cuda.h>

#include "xee_sweep_tma.cuh"

namespace xee {

enum { MODE_LINE_JACOBI = 3, MODE_LINE_CHEBYSHEV = 4 };

namespace ln {
constexpr int SEG = 8;             // points per thread = segment length of the line relaxation
constexpr int TW = 64, TH = 32;    // tile (grid points)
constexpr int NSEG = TW / SEG;     // 8 warps
constexpr int NT = NSEG * TH;      // 256 threads
#ifndef XEE_LINE_NSTAGE
#define XEE_LINE_NSTAGE 3
#endif
constexpr int NSTAGE = XEE_LINE_NSTAGE;
template <class T> struct Cfg {
  static constexpr int ES = (int)sizeof(T);
  static constexpr int V = 16 / ES;                 // elements per 16-byte chunk
  static constexpr int NV = SEG / V;                // chunks per segment
  static constexpr int XW = TW + 3 * V;             // psi box: columns i0-V .. i0+TW+2V-1, odd number of chunks
  static constexpr int FW = TW + V;                 // f / psi_{k-1} box: columns i0 .. i0+TW+V-1, odd number of chunks
  static constexpr int XP = XW * ES, FP = FW * ES;  // shared-memory row pitches (bytes)
  static constexpr int X_RAW = XP * (TH + 2), F_RAW = FP * TH;
  static constexpr int X_BYTES = (X_RAW + 127) / 128 * 128, F_BYTES = (F_RAW + 127) / 128 * 128;
  static constexpr int STAGE_BYTES = X_BYTES + 2 * F_BYTES;
  static constexpr int FAC_BYTES = 2 * F_BYTES;     // the two Thomas-factor planes of the current tile (thread-private cells)
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + FAC_BYTES;
  static_assert((XW / V) % 2 == 1 && (FW / V) % 2 == 1, "row pitches must be an odd number of 16-byte chunks");
};
__device__ __forceinline__ void cta_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// 16-byte shared-memory load into V consecutive elements.
__device__ __forceinline__ void lds16(uint32_t addr, double* v) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
}
__device__ __forceinline__ void lds16(uint32_t addr, float* v) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void ldg16(const double* p, double* v) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p)); v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void ldg16(const float* p, float* v) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void sts16(uint32_t addr, const double* v) { asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v[0]), "d"(v[1])); }
__device__ __forceinline__ void sts16(uint32_t addr, const float* v) { asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])); }
__device__ __forceinline__ void stg16(double* p, const double* v) { *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]); }
__device__ __forceinline__ void stg16(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }

// SEG + 2 consecutive elements of one shared-memory row: [0] = left neighbour of the segment, [1..SEG] the segment,
// [SEG+1] = right neighbour.  `seg_addr` = address of the segment's first element (16-byte aligned).
template <class T> __device__ __forceinline__ void load_row(uint32_t seg_addr, T (&w)[SEG + 2]) {
  constexpr int V = Cfg<T>::V, NV = Cfg<T>::NV;
  T t[V];
  lds16(seg_addr - 16u, t); w[0] = t[V - 1];
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    lds16(seg_addr + 16u * q, t);
#pragma unroll
    for (int e = 0; e < V; ++e) w[1 + q * V + e] = t[e];
  }
  lds16(seg_addr + 16u * NV, t); w[SEG + 1] = t[0];
}
template <class T> __device__ __forceinline__ void load_seg(uint32_t seg_addr, T (&w)[SEG]) {
  constexpr int V = Cfg<T>::V, NV = Cfg<T>::NV;
#pragma unroll
  for (int q = 0; q < NV; ++q) lds16(seg_addr + 16u * q, &w[q * V]);
}
}  // namespace ln

// Thomas factors of every radial segment of a shared operator: fac[0] = m(i) = 1 / (coe5(i) - lo(i) u(i-1)),
// fac[1] = u(i) = up(i) m(i), with lo = coe4 except at a segment start, up = coe6 except at a segment end; both 0 on
// boundary points, which therefore get a zero correction.  One thread per (segment, row).
template <class T>
__global__ void line_factor_kernel(const T* __restrict__ coe, T* __restrict__ fac, int nx, int ny) {
  const int sgi = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  const int i0 = sgi * ln::SEG;
  if (i0 >= nx) return;
  const size_t nn = (size_t)nx * ny;
  T u_prev = T(0);
  for (int e = 0; e < ln::SEG && i0 + e < nx; ++e) {
    const int i = i0 + e;
    const size_t o = (size_t)j * nx + i;
    const bool interior = i > 0 && i < nx - 1 && j > 0 && j < ny - 1;
    T m = T(0), u = T(0);
    if (interior) {
      const T lo = e == 0 ? T(0) : coe[3 * nn + o];
      const T up = e == ln::SEG - 1 ? T(0) : coe[5 * nn + o];
      m = T(1) / (coe[4 * nn + o] - lo * u_prev);
      u = up * m;
    }
    fac[o] = m; fac[nn + o] = u;
    u_prev = u;
  }
}

// Operator + factors repacked per tile in the order the sweep kernel's threads read them:
// pack[tile][plane 0..10][chunk q][thread][V] (planes 0..8 = coe1..coe9, 9 = m, 10 = u), zeros outside the field.  A warp's
// 128-bit load of (plane, chunk) is then one contiguous 512-byte run instead of 32 rows 4 KB apart.
constexpr int kLinePlanes = 11;
template <class T>
__global__ void __launch_bounds__(ln::NT) line_pack_kernel(const T* __restrict__ coe, const T* __restrict__ fac,
                                                           T* __restrict__ pack, int nx, int ny, int tiles_x) {
  using C = ln::Cfg<T>;
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int sg = tid >> 5, r = tid & 31;
  const int gi = (tile % tiles_x) * ln::TW + ln::SEG * sg, gj = (tile / tiles_x) * ln::TH + r;
  const size_t nn = (size_t)nx * ny;
  T* out = pack + (size_t)tile * kLinePlanes * ln::SEG * ln::NT;
  for (int k = 0; k < kLinePlanes; ++k) {
    const T* src = k < 9 ? coe + k * nn : fac + (k - 9) * nn;
    for (int e = 0; e < ln::SEG; ++e) {
      const bool in = gj < ny && gi + e < nx;
      out[((size_t)(k * C::NV + e / C::V) * ln::NT + tid) * C::V + e % C::V] = in ? src[(size_t)gj * nx + gi + e] : T(0);
    }
  }
}

template <class T>
struct LineArgs {
  const T* pack;           // operator + Thomas factors in tile/thread order (line_pack_kernel)
  T* dst;                  // psi_{k+1} (holds psi_{k-1} on entry: read through map_xm at the own cells only)
  long long field_stride;  // nx*ny
  int nx, ny, nbatch;
  T alpha, omega;
  const T* rho_ps;         // per-solve spectral radius (probe): omega computed in-kernel from cheb_k
  int cheb_k;
  const int* done;
  double* partial;         // [n][ntiles] sum of r^2 (CHECK)
  int tiles_x, tiles_y, nchunks, chunk;
};

template <class T, bool CHEB, bool CHECK>
__global__ void __launch_bounds__(ln::NT, 1)
    sweep_line_kernel(const __grid_constant__ LineArgs<T> a, const __grid_constant__ CUtensorMap map_x,
                      const __grid_constant__ CUtensorMap map_xm, const __grid_constant__ CUtensorMap map_f) {
  using namespace ln;
  using tma::mbar_init; using tma::mbar_expect_tx; using tma::mbar_wait; using tma::tma_load_3d;
  using C = Cfg<T>;
  constexpr int V = C::V, NV = C::NV;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[NSTAGE];
  __shared__ double red[NT / 32];

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int ntiles = a.tiles_x * a.tiles_y;
  const int nunits = ntiles * a.nchunks;
  const size_t nn = (size_t)a.field_stride;

  // ---- prefetch cursor (thread 0): the same (unit, solve) sequence as the consumers, NSTAGE items ahead
  int pu = blockIdx.x, pn = -1;
  uint32_t issued = 0;
  auto issue_next = [&]() {
    for (;;) {
      if (pu >= nunits) return false;
      const int ch = pu / ntiles;
      const int n0 = ch * a.chunk, n1 = min(n0 + a.chunk, a.nbatch);
      if (pn < 0) pn = n0; else ++pn;
      if (pn >= n1) { pu += gridDim.x; pn = -1; continue; }
      if (a.done != nullptr && a.done[pn]) continue;
      const int tile = pu % ntiles;
      const int i0 = (tile % a.tiles_x) * TW, j0 = (tile / a.tiles_x) * TH;
      const int s = issued % NSTAGE;
      unsigned char* st = smem_raw + (size_t)s * C::STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], (uint32_t)(C::X_RAW + (CHEB ? 2 : 1) * C::F_RAW));
      tma_load_3d(st, &map_x, i0 - V, j0 - 1, pn, &full_bar[s]);
      tma_load_3d(st + C::X_BYTES, &map_f, i0, j0, pn, &full_bar[s]);
      if (CHEB) tma_load_3d(st + C::X_BYTES + C::F_BYTES, &map_xm, i0, j0, pn, &full_bar[s]);
      ++issued;
      return true;
    }
  };
  if (tid == 0)
    for (int q = 0; q < NSTAGE; ++q)
      if (!issue_next()) break;

  const int sg = tid >> 5;          // warp = segment column of the tile
  const int r = tid & 31;           // lane = row of the tile
  const uint32_t sm0 = tma::smem_u32(smem_raw);
  const uint32_t xofs = (uint32_t)((r + 1) * C::XP + (V + SEG * sg) * C::ES);   // own segment in the psi box
  const uint32_t fofs = (uint32_t)(C::X_BYTES + r * C::FP + SEG * sg * C::ES);  // ... in the f box
  const uint32_t faca = sm0 + (uint32_t)(NSTAGE * C::STAGE_BYTES + r * C::FP + SEG * sg * C::ES);   // own cells of the factor planes
  uint32_t it = 0;

  for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
    const int tile = u % ntiles, ch = u / ntiles;
    const int i0 = (tile % a.tiles_x) * TW, j0 = (tile / a.tiles_x) * TH;
    const int n0 = ch * a.chunk, n1 = min(n0 + a.chunk, a.nbatch);
    const int gi = i0 + SEG * sg, gj = j0 + r;
    const bool row_in = gj < a.ny;
    const bool seg_full = row_in && gi + SEG <= a.nx;      // whole segment inside the field: vector accesses
    const long long gofs = (long long)gj * a.nx + gi;
    unsigned inb = 0;   // bit e: point (gi+e, gj) is a domain-interior point (residual norm)
#pragma unroll
    for (int e = 0; e < SEG; ++e)
      if (gi + e >= 1 && gi + e < a.nx - 1 && gj >= 1 && gj < a.ny - 1) inb |= 1u << e;
    // the 9 coefficients of the thread's points stay in registers for the whole chunk; its 2 x 8 Thomas factors go to
    // thread-private cells of shared memory (read back by the same thread only: no barrier needed)
    T cf[9][SEG];
    {
      const T* pk = a.pack + ((size_t)tile * kLinePlanes * SEG * NT + (size_t)tid * V);
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int q = 0; q < NV; ++q) ldg16(pk + (size_t)(k * NV + q) * NT * V, &cf[k][q * V]);
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        T t[V];
        ldg16(pk + (size_t)(9 * NV + q) * NT * V, t); sts16(faca + 16u * q, t);
        ldg16(pk + (size_t)(10 * NV + q) * NT * V, t); sts16(faca + C::F_BYTES + 16u * q, t);
      }
    }
    for (int n = n0; n < n1; ++n) {
      if (a.done != nullptr && a.done[n]) continue;
      const uint32_t s = it % NSTAGE;
      mbar_wait(&full_bar[s], (it / NSTAGE) & 1);
      const uint32_t sb = sm0 + s * C::STAGE_BYTES;
      const uint32_t xa = sb + xofs;
      // ---- residual: the nine-term sum in the reference's order (rows j+1, j, j-1; i-1, i, i+1), minus f
      T acc[SEG];
      {
        T w[SEG + 2];
        load_row<T>(xa + C::XP, w);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = cf[0][e] * w[e];
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[1][e], w[e + 1], acc[e]);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[2][e], w[e + 2], acc[e]);
        load_row<T>(xa, w);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[3][e], w[e], acc[e]);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[4][e], w[e + 1], acc[e]);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[5][e], w[e + 2], acc[e]);
        load_row<T>(xa - C::XP, w);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[6][e], w[e], acc[e]);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[7][e], w[e + 1], acc[e]);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = Rn<T>::fma(cf[8][e], w[e + 2], acc[e]);
      }
      {
        T fv[SEG];
        load_seg<T>(sb + fofs, fv);
#pragma unroll
        for (int e = 0; e < SEG; ++e) acc[e] = acc[e] - fv[e];
      }
      double rr = 0.0;
      if (CHECK) {
#pragma unroll
        for (int e = 0; e < SEG; ++e)
          if ((inb >> e) & 1u) rr += (double)acc[e] * (double)acc[e];
      }
      // ---- Thomas solve of the segment: forward  y(i) = (r(i) - coe4(i) y(i-1)) m(i),  back  z(i) = y(i) - u(i) z(i+1)
      {
        T mf[SEG];
        load_seg<T>(faca, mf);
        acc[0] = acc[0] * mf[0];
#pragma unroll
        for (int e = 1; e < SEG; ++e) acc[e] = Rn<T>::fma(-cf[3][e], acc[e - 1], acc[e]) * mf[e];
      }
      {
        T uf[SEG];
        load_seg<T>(faca + C::F_BYTES, uf);
#pragma unroll
        for (int e = SEG - 2; e >= 0; --e) acc[e] = Rn<T>::fma(-uf[e], acc[e + 1], acc[e]);
      }
      // ---- update
      T out[SEG];
      {
        T x[SEG];
        load_seg<T>(xa, x);
        if (!CHEB) {
#pragma unroll
          for (int e = 0; e < SEG; ++e) out[e] = Rn<T>::fma(-a.alpha, acc[e], x[e]);
        } else {
          T xm[SEG];
          load_seg<T>(sb + fofs + C::F_BYTES, xm);
          const T om = a.rho_ps ? (T)cheb_omega(a.cheb_k, (double)a.rho_ps[n]) : a.omega;
#pragma unroll
          for (int e = 0; e < SEG; ++e) out[e] = Rn<T>::fma(om, (x[e] - acc[e]) - xm[e], xm[e]);
        }
      }
      cta_bar_sync();                           // every thread is done with the stage
      if (tid == 0) issue_next();               // ... refill it, NSTAGE items ahead
      T* const o = a.dst + ((size_t)n * nn + gofs);
      if (seg_full) {
#pragma unroll
        for (int q = 0; q < NV; ++q) stg16(o + q * V, &out[q * V]);
      } else if (row_in) {
#pragma unroll
        for (int e = 0; e < SEG; ++e)
          if (gi + e < a.nx) o[e] = out[e];
      }
      if (CHECK) {
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) rr += __shfl_down_sync(0xffffffffu, rr, q);
        if ((tid & 31) == 0) red[tid >> 5] = rr;
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
