// Times sweep_tb_kernel (v4) on the bench shape for one set of tuning knobs:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DXEE_TB_P=8 -DXEE_TB_G=4 ... tb_probe.cu -o tb_probe
//   ./tb_probe <depth> [nbatch] [chunk] [nx ny]
#include <vector>
#include "../../xlab_ee_fortran_b200/csrc/xee_sweep_tb.cuh"
using namespace xee;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static Enc enc;
static CUtensorMap mk(const double* base, int nx, int ny, int nb) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nb};
  const cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
  const cuuint32_t box[3] = {tb::W, tb::H, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}
__global__ void fill(double* p, size_t n, double s) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = s * (double)((i * 2654435761u) % 1000) / 1000.0; }
__global__ void fillcoe(double* c, size_t nn) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nn; i += (size_t)gridDim.x * blockDim.x) {
    for (int k = 0; k < 9; ++k) c[k * nn + i] = k == 4 ? -4.0 : 0.5;
    c[9 * nn + i] = 0.25;
  }
}
int main(int argc, char** argv) {
  const int depth = argc > 1 ? atoi(argv[1]) : 4, nb = argc > 2 ? atoi(argv[2]) : 512, chunk = argc > 3 ? atoi(argv[3]) : 32;
  const int nx = argc > 5 ? atoi(argv[4]) : 512, ny = argc > 5 ? atoi(argv[5]) : 256;
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q)); enc = (Enc)fp;
  const size_t nn = (size_t)nx * ny, tot = nn * nb;
  double *x[4], *f, *coe;
  for (auto& p : x) { CK(cudaMalloc(&p, tot * 8)); fill<<<1024, 256>>>(p, tot, 1.0); }
  CK(cudaMalloc(&f, tot * 8)); fill<<<1024, 256>>>(f, tot, 0.1);
  CK(cudaMalloc(&coe, nn * 10 * 8)); fillcoe<<<256, 256>>>(coe, nn);
  CUtensorMap mx[4]; for (int k = 0; k < 4; ++k) mx[k] = mk(x[k], nx, ny, nb);
  CUtensorMap mf = mk(f, nx, ny, nb);
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  TbArgs<double> A{};
  A.coe = coe; A.field_stride = (long long)nn; A.nx = nx; A.ny = ny; A.nbatch = nb; A.nsweeps = depth; A.tbh = depth; A.alpha = 1.0;
  for (int k = 0; k < depth; ++k) A.omega[k] = 1.2;
  const int sx = tb::W - 2 * depth, sy = tb::H - 2 * depth;
  A.tiles_x = nx <= tb::W ? 1 : (nx - tb::W + sx - 1) / sx + 1; A.tiles_y = ny <= tb::H ? 1 : (ny - tb::H + sy - 1) / sy + 1;
  A.chunk = chunk; A.nchunks = (nb + chunk - 1) / chunk;
  const long long units = (long long)A.tiles_x * A.tiles_y * A.nchunks;
  const int grid = (int)std::min<long long>(sms, units);
  auto kern = sweep_tb_kernel<double, XEE_ARITH_FAST, MODE_CHEBYSHEV, false>;
  const int smem = tb::Cfg<double>::SMEM_BYTES;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 20;
  for (int w = 0; w < 3 + reps; ++w) {
    if (w == 3) CK(cudaEventRecord(e0));
    const int pin = w & 1, po = pin ^ 1;
    A.out_new = x[2 * po]; A.out_prev = x[2 * po + 1];
    kern<<<grid, tb::NT, smem>>>(A, mx[2 * pin], mx[2 * pin + 1], mf);
  }
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / reps;
  const double alg = (double)nb * (nx - 2) * (ny - 2) * 8.0 * (4 + 9.0 / nb) * depth;
  printf("P=%d G=%d RG=%d NSTAGE=%d depth=%d nb=%d chunk=%d grid=%d tiles=%dx%d smem=%d : %.1f us/pass, %.1f us/sweep, alg %.0f GB/s\n", tb::P, tb::G, tb::RG,
         tb::NSTAGE, depth, nb, chunk, grid, A.tiles_x, A.tiles_y, smem, us, us / depth, alg / (us * 1e-6) / 1e9);
  return 0;
}
