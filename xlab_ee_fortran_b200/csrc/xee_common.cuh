// xee_common.cuh — shared device helpers for the B200 (sm_100a) elliptic-solve kernels.
//
// Arithmetic policies
//   STRICT : every multiply, add and divide is a separately rounded IEEE operation in the
//            reference's left-to-right order (xtt-lib-fortran/elliptic_tools.f90:77-85,
//            :190, :238).  The __d*_rn / __f*_rn intrinsics are never contracted into FMAs
//            by nvcc, so the iterates are bit-identical to gfortran -O0 on x86-64.
//   FAST   : FMA chains and a precomputed reciprocal of -coe5.  Same fixed point (L psi = f);
//            the iterates differ from STRICT by rounding only.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "../../include/xee_b200.h"

namespace xee {

template <class T>
struct Rn;
template <>
struct Rn<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double rcp(double a) { return __drcp_rn(a); }
  static __device__ __forceinline__ double abs(double a) { return fabs(a); }
  static __host__ __device__ __forceinline__ double huge() { return 1.7976931348623157e308; }
};
template <>
struct Rn<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float rcp(float a) { return __frcp_rn(a); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __host__ __device__ __forceinline__ float huge() { return 3.4028234663852886e38f; }
};

// acc + c*p in the chosen policy.
template <class T, int ARITH>
__device__ __forceinline__ T madd(T acc, T c, T p) {
  if (ARITH == XEE_ARITH_STRICT) return Rn<T>::add(acc, Rn<T>::mul(c, p));
  return Rn<T>::fma(c, p, acc);
}

// 9-point apply, slots 1..3 at j+1, 4..6 at j, 7..9 at j-1 (elliptic_tools.f90:19-21, 77-85).
// p[0..2] = psi(i-1..i+1, j+1), p[3..5] = row j, p[6..8] = row j-1.
template <class T, int ARITH>
__device__ __forceinline__ T apply9(const T (&c)[9], const T (&p)[9]) {
  T s = Rn<T>::mul(c[0], p[0]);
#pragma unroll
  for (int k = 1; k < 9; ++k) s = madd<T, ARITH>(s, c[k], p[k]);
  return s;
}

// Weighted-Jacobi update  psi + alpha*r/(-coe5)  (elliptic_tools.f90:238).
//   STRICT: (alpha*r)/(-c5) with a true division.   FAST: fma(alpha*rcp, r, psi), rcp = 1/(-c5).
template <class T, int ARITH>
__device__ __forceinline__ T jacobi_update(T center, T r, T alpha, T c5, T rcp) {
  if (ARITH == XEE_ARITH_STRICT) return Rn<T>::add(center, Rn<T>::div(Rn<T>::mul(alpha, r), -c5));
  return Rn<T>::fma(alpha * rcp, r, center);
}

// Deterministic block-wide sum of one double per thread (fixed shuffle tree, fixed smem order).
// Result valid in thread 0.  `warps` = blockDim/32 (<= 32).
__device__ __forceinline__ double block_sum(double v, double* smem_warp /*[32]*/, int tid, int warps) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) smem_warp[tid >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (tid == 0)
    for (int w = 0; w < warps; ++w) s += smem_warp[w];
  __syncthreads();
  return s;
}

}  // namespace xee
