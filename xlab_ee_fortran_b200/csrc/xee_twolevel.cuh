// xee_twolevel.cuh — coarse-grid half of the TWO-LEVEL block-line methods (XEE_METHOD_LINE2_*), sm_100a.
//
// Same discrete problem, same residual r = L psi - f (do_elliptic's nine-term sum, xtt-lib-fortran/elliptic_tools.f90:77-85)
// and the same stop rule (:193-233) as solve_elliptic.  The approximate inverse applied to the residual is ADDITIVE:
//
//     z = M^-1 r  +  P Ac^-1 P^T r,          psi' = psi - gamma z        (then Chebyshev acceleration)
//
// M = the 32-point radial block systems of xee_sweep_line.cuh, P = bilinear interpolation from a coarse grid with nodes
// every HR = 16 radial and HZ = 16 vertical grid points (Dirichlet boundary: no nodes on it), Ac = P^T L P the Galerkin
// coarse operator.  Why: the block-line Chebyshev iteration still needs ~950 sweeps at 512x256 because the modes that are
// smooth in BOTH directions are damped slowly; the coarse space carries exactly those.  Measured with scipy on the bench
// operator (scripts/prototypes/twolevel_spectrum.py): condition number of the preconditioned operator 3,877 -> 309,
// Chebyshev sweeps to 1e-12 rms(f): 873-922 -> 251-255 (31 x 15 = 465 coarse unknowns).
//
// Per sweep: (1) sweep_line_kernel<.., TWO> relaxes the lines and leaves P^T r of its tile as 32 doubles per (tile, solve);
// (2) coarse_gather_kernel sums the tile contributions per coarse node (fixed order: deterministic); (3) coarse_gemm_kernel
// applies the dense inverse Ac^-1 to the whole batch, E[n] = Ac^-1 Rc[n] (465 x 465 x nbatch: one small fp64 GEMM on the FP64
// tensor cores), scaled by -omega_k gamma.  The prolongation P E is NOT a pass over the fields: the iterate is kept as
// (stored field y, coarse vector c), psi = y + P c, and the sweep kernel adds P c to what it reads (xee_sweep_line.cuh);
// prolong_add_kernel applies it once, when a solve ends.  Once per operator: galerkin_kernel assembles Ac (9-point coarse
// stencil, a banded matrix), band_lu_kernel factors the band (no pivoting: Ac is definite like L) and band_inverse_kernel
// forms the dense inverse column by column.
#pragma once
#include "xee_kernels.cuh"

namespace xee {
namespace tl {

constexpr int HR = 16, HZ = 16;   // coarse spacing in grid points (HR = 2 thread segments, HZ = the tile height of the v5 kernel)

struct Dims {
  int ncx, ncz, nc;   // coarse nodes at i = HR p (p = 1..ncx), j = HZ q (q = 1..ncz); node index = (q-1) ncx + (p-1)
  int px, pz;         // the coarse correction is stored with a zero rim, node (p, q) at [q][p + 1]: pz = ncz + 2 rows of pitch
                      // px = ncx + 3 rounded up to even (one spare column in front, so that the TMA boxes of the sweep kernel,
                      // which start at node 4 tx - 1, begin on a 16-byte boundary)
  int ncp;            // nc padded to a multiple of 64 (GEMM tiles)
};
constexpr int COL0 = 1;   // column of node 0
inline Dims dims(int nx, int ny) {
  Dims d{};
  d.ncx = (nx - 2) / HR; d.ncz = (ny - 2) / HZ; d.nc = d.ncx * d.ncz;
  d.px = (d.ncx + 3 + 1) & ~1; d.pz = d.ncz + 2;
  d.ncp = (d.nc + 63) / 64 * 64;
  return d;
}

// 1-D hat function of node c (grid index) at grid index i; 0 outside the interior [1, n-2] (boundary points are not unknowns)
__device__ __forceinline__ double hat(int i, int c, int h, int n) {
  if (i < 1 || i > n - 2) return 0.0;
  const int dd = i > c ? i - c : c - i;
  return dd >= h ? 0.0 : 1.0 - (double)dd / (double)h;
}

// Ac = P^T L P, one block per coarse node q (column): L applied to the hat function of q on its support, tested against
// the hats of the 9 neighbouring nodes.  Ac is [ncp][ncp] row-major, zero elsewhere, identity on the padding.
template <class T>
__global__ void __launch_bounds__(256) galerkin_kernel(const T* __restrict__ coe, double* __restrict__ Ac, int nx, int ny,
                                                       int ncx, int ncz, int ncp) {
  __shared__ double red[32];
  const int q = blockIdx.x, qx = q % ncx + 1, qz = q / ncx + 1;
  const int ic = HR * qx, jc = HZ * qz;
  const size_t nn = (size_t)nx * ny;
  coe += (size_t)blockIdx.y * kPlanes * nn;            // operator set (one per solve for a time series)
  Ac += (size_t)blockIdx.y * ncp * ncp;
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  constexpr int W = 2 * HR + 1, H = 2 * HZ + 1;
  for (int t = threadIdx.x; t < W * H; t += blockDim.x) {
    const int i = ic - HR + t % W, j = jc - HZ + t / W;
    if (i < 1 || i > nx - 2 || j < 1 || j > ny - 2) continue;
    const size_t o = (size_t)j * nx + i;
    // L phi_q at (i, j): slots 1..3 at j+1, 4..6 at j, 7..9 at j-1, each (i-1, i, i+1)
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int di = k % 3 - 1, dj = 1 - k / 3;
      v += (double)coe[k * nn + o] * hat(i + di, ic, HR, nx) * hat(j + dj, jc, HZ, ny);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int a = k % 3 - 1, b = k / 3 - 1;
      acc[k] += hat(i, ic + a * HR, HR, nx) * hat(j, jc + b * HZ, HZ, ny) * v;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const double tot = block_sum(acc[k], red, threadIdx.x, blockDim.x / 32);
    const int px = qx + (k % 3 - 1), pz = qz + (k / 3 - 1);
    if (threadIdx.x == 0 && px >= 1 && px <= ncx && pz >= 1 && pz <= ncz)
      Ac[(size_t)((pz - 1) * ncx + (px - 1)) * ncp + q] = tot;
  }
}
static __global__ void pad_identity_kernel(double* __restrict__ A, int nc, int ncp) {
  const int k = nc + blockIdx.x * blockDim.x + threadIdx.x;
  if (k < ncp) A[(size_t)blockIdx.y * ncp * ncp + (size_t)k * ncp + k] = 1.0;
}

// Inverse of the Galerkin operator.  With the nodes numbered row by row Ac is BANDED (9-point coarse stencil: half-bandwidth
// ncx + 1), so it is factored as a band (LU without pivoting: Ac is definite like L) and the dense inverse is obtained
// column by column with band substitutions: O(nc^2 bw) instead of the O(nc^3) of a dense elimination.
// band_lu_kernel: one block per operator set; the band is held as Lb[i][d] = L(i, i-bw+d) (d < bw, unit diagonal not stored)
// and Ub[i][d] = U(i, i+d) (d <= bw), both [nc][bw+1] doubles in `band` (2 nc (bw+1) doubles per set).
static __global__ void __launch_bounds__(1024) band_lu_kernel(const double* __restrict__ Ac, double* __restrict__ band, int nc, int ncp, int bw) {
  extern __shared__ double wk[];                 // the active window: matrix rows k..k+bw, each as columns i-bw..i+bw (2 bw + 1 values),
  const int W = 2 * bw + 1, R = bw + 1;          // matrix row i in window slot i % (bw + 1)
  Ac += (size_t)blockIdx.x * ncp * ncp;
  double* Lb = band + (size_t)blockIdx.x * 2 * nc * (bw + 1);
  double* Ub = Lb + (size_t)nc * (bw + 1);
  const int tid = threadIdx.x, nt = blockDim.x;
  auto load_row = [&](int i) {                   // (all threads) matrix row i -> its slot
    for (int t = tid; t < W; t += nt) {
      const int j = i - bw + t;
      wk[(i % R) * W + t] = (i < nc && j >= 0 && j < nc) ? Ac[(size_t)i * ncp + j] : 0.0;
    }
  };
  for (int i = 0; i <= bw; ++i) load_row(i);
  __syncthreads();
  for (int k = 0; k < nc; ++k) {
    const double* pr = wk + (k % R) * W;         // pivot row k: U(k, k + t) at position bw + t
    const double piv = pr[bw];
    for (int t = tid; t <= bw; t += nt) Ub[(size_t)k * (bw + 1) + t] = pr[bw + t];
    // eliminate column k from rows k+1..k+bw: row k+r has its column k at position bw - r, column k + c at bw - r + c
    for (int t = tid; t < bw * (bw + 1); t += nt) {
      const int r = 1 + t / (bw + 1), c = t % (bw + 1);
      if (k + r < nc) {
        double* row = wk + ((k + r) % R) * W;
        const double l = row[bw - r] / piv;
        if (c == 0) Lb[(size_t)(k + r) * (bw + 1) + (bw - r)] = l;            // L(k+r, k)
        else row[bw - r + c] -= l * pr[bw + c];
      }
    }
    __syncthreads();
    load_row(k + bw + 1);                        // the slot of row k is free now
    __syncthreads();
  }
}
// band_inverse_kernel: thread = one column j of the inverse of one operator set: forward substitution L y = e_j, back
// substitution U x = y, the running vector kept in the output column itself (consecutive threads -> consecutive addresses).
static __global__ void __launch_bounds__(128) band_inverse_kernel(const double* __restrict__ band, double* __restrict__ Ainv, int nc, int ncp, int bw) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nc) return;
  const double* Lb = band + (size_t)blockIdx.y * 2 * nc * (bw + 1);
  const double* Ub = Lb + (size_t)nc * (bw + 1);
  double* x = Ainv + (size_t)blockIdx.y * ncp * ncp + j;          // x(i) = x[i * ncp]
  for (int i = 0; i < nc; ++i) {                                   // y(i) = e_j(i) - sum_{k = i-bw}^{i-1} L(i,k) y(k); y(i) = 0 for i < j
    double t = i == j ? 1.0 : 0.0;
    if (i > j) {
      const int k0 = max(max(i - bw, 0), j);
      for (int k = k0; k < i; ++k) t = fma(-Lb[(size_t)i * (bw + 1) + (k - i + bw)], x[(size_t)k * ncp], t);
    }
    x[(size_t)i * ncp] = t;
  }
  for (int i = nc - 1; i >= 0; --i) {                              // x(i) = (y(i) - sum_{k = i+1}^{i+bw} U(i,k) x(k)) / U(i,i)
    double t = x[(size_t)i * ncp];
    const int k1 = min(i + bw, nc - 1);
    for (int k = i + 1; k <= k1; ++k) t = fma(-Ub[(size_t)i * (bw + 1) + (k - i)], x[(size_t)k * ncp], t);
    x[(size_t)i * ncp] = t / Ub[(size_t)i * (bw + 1)];
  }
}

// Rc[n][node] = P^T r from the per-tile moments written by sweep_line_kernel<.., TWO>: part[n][tile][kind 0..3][segment 0..7],
// kinds 0/1 = sum_j (1 - j/16) S0/S1 -> coarse row of the tile's first row, 2/3 = sum_j (j/16) S0/S1 -> the next coarse row.
// A segment s (8 points from i = 8 s) lies in radial cell c = s / 2 at offset o0 = 8 (s & 1): its weight sum towards the cell's
// right node is (o0 S0 + S1)/16, towards its left node S0 minus that.
__device__ __forceinline__ double gather_node(const double* __restrict__ pn, int idx, int tiles_x, int tiles_y, int ncx) {
  const int nsegs = tiles_x * 8;
  const int p = idx % ncx + 1, q = idx / ncx + 1;
  double tot = 0.0;
#pragma unroll
  for (int zc = 0; zc < 2; ++zc) {            // z-cell q-1 (kinds 2,3) then z-cell q (kinds 0,1)
    const int ty = q - 1 + zc, kb = zc == 0 ? 2 : 0;
    if (ty < 0 || ty >= tiles_y) continue;
#pragma unroll
    for (int ss = 0; ss < 4; ++ss) {          // segments 2p-2, 2p-1 (left cell), 2p, 2p+1 (right cell)
      const int s = 2 * p - 2 + ss;
      if (s < 0 || s >= nsegs) continue;
      const double* t = pn + (size_t)(ty * tiles_x + s / 8) * 32 + (s & 7);
      const double s0 = t[kb * 8], s1 = t[(kb + 1) * 8];
      const double right = ((double)(8 * (s & 1)) * s0 + s1) * (1.0 / HR);
      tot += ss < 2 ? right : s0 - right;
    }
  }
  return tot;
}
static __global__ void coarse_gather_kernel(const double* __restrict__ part, double* __restrict__ Rc, const int* __restrict__ done,
                                     int ntiles, int tiles_x, int tiles_y, int ncx, int ncz, int ncp) {
  const int n = blockIdx.x;
  if (done && done[n]) return;
  const double* pn = part + (size_t)n * ntiles * 32;
  for (int idx = threadIdx.x; idx < ncx * ncz; idx += blockDim.x) Rc[(size_t)n * ncp + idx] = gather_node(pn, idx, tiles_x, tiles_y, ncx);
}

// Cv[n] = scale * Ainv Rc[n] for the whole batch: C[i][n] = sum_k Ainv[i][k] Rc[n][k], one small fp64 GEMM (ncp x nbatch x ncp)
// on the FP64 tensor cores (mma.sync m8n8k4: DMMA).  Both operands are k-contiguous in memory, which is exactly the
// row-major A / column-major B fragment layout of the instruction (thread (g = lane/4, t = lane%4) holds A[g][t] and
// B[t][g]), so the fragments are loaded straight from global memory (2 MB + 2 MB, L2-resident, re-use through L1) with no
// shared-memory staging.  Ainv and Rc
// are padded to ncp (multiple of 64) with zeros: no edge handling in k or i; rows n >= nbatch are clamped and discarded.
// Written into the rimmed layout Cv[n][pz][px] (interior only).
constexpr int GM = 32, GN = 32;
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
constexpr int GSPLIT = 4;   // the k range is split over GSPLIT blocks (more warps in flight: the loop is latency-bound), partial
                            // products summed in a fixed order by coarse_finish_kernel (deterministic, no atomics)
static __global__ void __launch_bounds__(128) coarse_gemm_kernel(const double* __restrict__ Ainv, const double* __restrict__ Rc,
                                                                 double* __restrict__ Ck, int nb, int ncp) {
  // warp tile 16 (i) x 16 (n) = 2 x 2 accumulator fragments, block = 4 warps = 32 x 32; grid (ncp/32, nbatch/32, GSPLIT)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int i0 = blockIdx.x * GM + (warp & 1) * 16, n0 = blockIdx.y * GN + (warp >> 1) * 16;
  const int klen = ncp / GSPLIT, kbeg = blockIdx.z * klen;      // ncp is a multiple of 64
  double c[2][2][2] = {};
  const double* ap[2]; const double* bp[2];
#pragma unroll
  for (int x = 0; x < 2; ++x) ap[x] = Ainv + (size_t)(i0 + 8 * x + g) * ncp + kbeg + t;
#pragma unroll
  for (int y = 0; y < 2; ++y) bp[y] = Rc + (size_t)min(n0 + 8 * y + g, nb - 1) * ncp + kbeg + t;
#pragma unroll 16
  for (int k0 = 0; k0 < klen; k0 += 4) {
    double a[2], b[2];
#pragma unroll
    for (int x = 0; x < 2; ++x) a[x] = __ldg(ap[x] + k0);
#pragma unroll
    for (int y = 0; y < 2; ++y) b[y] = __ldg(bp[y] + k0);
#pragma unroll
    for (int x = 0; x < 2; ++x)
#pragma unroll
      for (int y = 0; y < 2; ++y) dmma_m8n8k4(c[x][y], a[x], b[y]);
  }
  double* out = Ck + (size_t)blockIdx.z * nb * ncp;
#pragma unroll
  for (int y = 0; y < 2; ++y)
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const int n = n0 + 8 * y + 2 * t + v;
      if (n >= nb) continue;
#pragma unroll
      for (int x = 0; x < 2; ++x) out[(size_t)n * ncp + i0 + 8 * x + g] = c[x][y][v];
    }
}
// Cv[n][node] = scale * (sum of the GSPLIT partial products), written into the rimmed layout (interior only).
static __global__ void coarse_finish_kernel(const double* __restrict__ Ck, double* __restrict__ Cv, const int* __restrict__ done,
                                            double scale, int nb, int nc, int ncp, int ncx, int px, int pzpx) {
  const int n = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc || (done && done[n])) return;
  double t = 0.0;
#pragma unroll
  for (int q = 0; q < GSPLIT; ++q) t += Ck[((size_t)q * nb + n) * ncp + i];
  Cv[(size_t)n * pzpx + (size_t)(i / ncx + 1) * px + (i % ncx + 1) + COL0] = scale * t;
}

// The same product for a handful of solves (the spectral probes run on 1 solve): one warp per coarse unknown i, lanes over k,
// shuffle reduction (the DMMA kernel would walk its 128 k-steps with 4 warps: ~50 us of pure latency).
static __global__ void __launch_bounds__(256) coarse_matvec_kernel(const double* __restrict__ Ainv, long long ainv_stride,
                                                                   const double* __restrict__ Rc, double* __restrict__ Cv,
                                                                   const int* __restrict__ done, double scale, const double* __restrict__ scale_ps,
                                                                   int nb, int nc, int ncp, int ncx, int px, int pzpx) {
  // ainv_stride != 0: one coarse operator per solve (time series); scale_ps: per-solve factor (per-solve Chebyshev weight)
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, n = blockIdx.y;
  if (i >= nc || (done && done[n])) return;
  const double* ar = Ainv + (size_t)n * ainv_stride + (size_t)i * ncp; const double* rc = Rc + (size_t)n * ncp;
  double t = 0.0;
  for (int k0 = lane; k0 < ncp; k0 += 256) {     // all loads of a step in flight before the sum (same order as a plain loop)
    double av[8], rv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { const bool in = k0 + 32 * q < ncp; av[q] = in ? __ldg(ar + k0 + 32 * q) : 0.0; rv[q] = in ? __ldg(rc + k0 + 32 * q) : 0.0; }
#pragma unroll
    for (int q = 0; q < 8; ++q) t = fma(av[q], rv[q], t);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) Cv[(size_t)n * pzpx + (size_t)(i / ncx + 1) * px + (i % ncx + 1) + COL0] = (scale_ps ? scale_ps[n] : scale) * t;
}
// Gather and product in one launch for a handful of solves with a shared operator (single solves, the spectral probes: every
// kernel of such a sweep is latency, not work): every block first forms the whole P^T r of its solve in shared memory (the same
// sums in the same order as coarse_gather_kernel), then its 8 rows of the product.
constexpr int kMaxNcSmem = 2048;
static __global__ void __launch_bounds__(256) coarse_gather_matvec_kernel(const double* __restrict__ Ainv, const double* __restrict__ part,
                                                                          double* __restrict__ Cv, const int* __restrict__ done, double scale,
                                                                          int ntiles, int tiles_x, int tiles_y, int nc, int ncp, int ncx,
                                                                          int px, int pzpx) {
  __shared__ double rc[kMaxNcSmem];
  const int n = blockIdx.y;
  if (done && done[n]) return;
  const double* pn = part + (size_t)n * ntiles * 32;
  for (int idx = threadIdx.x; idx < nc; idx += 256) rc[idx] = gather_node(pn, idx, tiles_x, tiles_y, ncx);
  __syncthreads();
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= nc) return;
  const double* ar = Ainv + (size_t)i * ncp;
  double t = 0.0;
  // (ncp is a multiple of 64; Rc is zero-padded to it.)  The row comes from L2: all loads of a step are issued before the sum,
  // which keeps its order
  for (int k0 = lane; k0 < ncp; k0 += 256) {
    double av[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) av[q] = k0 + 32 * q < ncp ? __ldg(ar + k0 + 32 * q) : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int k = k0 + 32 * q; t = fma(av[q], k < nc ? rc[k] : 0.0, t); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) Cv[(size_t)n * pzpx + (size_t)(i / ncx + 1) * px + (i % ncx + 1) + COL0] = scale * t;
}
// per-solve factor of the coarse correction, -omega_k(rho_n) gamma (one operator per solve: the sweep kernel computes the same
// omega from the same rho)
template <class T>
__global__ void coarse_scale_kernel(const T* __restrict__ rho_ps, int cheb_k, double gamma, double* __restrict__ scale_ps, int nb) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < nb) scale_ps[n] = -(double)(T)cheb_omega(cheb_k, (double)rho_ps[n]) * gamma;
}

// x += P Cv on the interior points (bilinear; the rim of Cv is zero, which is the Dirichlet condition of the correction).
// Thread = one 8-point radial segment of one row (it lies inside one coarse cell: HR = 16): the correction is linear along
// it, base + slope * e, from the four coarse values of the cell interpolated in z.  64 bytes read and written per thread.
template <class T>
__global__ void __launch_bounds__(256) prolong_add_kernel(T* __restrict__ x, const double* __restrict__ Cv, const int* __restrict__ done,
                                                          int nx, int ny, int px, int pzpx) {
  const int n = blockIdx.z;
  if (done && done[n]) return;
  const int nseg = (nx + 7) / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = idx / nseg, sgm = idx % nseg;
  if (j < 1 || j > ny - 2) return;
  const double* cv = Cv + (size_t)n * pzpx;
  const int q0 = j / HZ, p0 = sgm / 2; const double tz = (double)(j % HZ) * (1.0 / HZ);
  cv += COL0;
  const double c00 = cv[q0 * px + p0], c01 = cv[q0 * px + p0 + 1], c10 = cv[(q0 + 1) * px + p0], c11 = cv[(q0 + 1) * px + p0 + 1];
  const double vl = fma(tz, c10 - c00, c00), vr = fma(tz, c11 - c01, c01);
  const double sl = (vr - vl) * (1.0 / HR), base = fma(sl, (double)(8 * (sgm & 1)), vl);
  T* row = x + ((size_t)n * ny + j) * nx + 8 * sgm;
  if (8 * sgm + 8 <= nx - 1 && sgm > 0 && sizeof(T) == 8) {      // whole segment interior: vector accesses
    double2* r2 = reinterpret_cast<double2*>(row);
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
      double2 v = r2[qd];
      v.x += fma(sl, (double)(2 * qd), base); v.y += fma(sl, (double)(2 * qd + 1), base);
      r2[qd] = v;
    }
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = 8 * sgm + e;
      if (i >= 1 && i <= nx - 2) row[e] = (T)((double)row[e] + fma(sl, (double)e, base));
    }
  }
}

// a <- a - b (power iteration on M^-1 L: with f = 0 one Jacobi sweep gives b = a - M^-1 L a)
template <class T>
__global__ void diff_kernel(T* __restrict__ a, const T* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = a[i] - b[i];
}

}  // namespace tl
}  // namespace xee
