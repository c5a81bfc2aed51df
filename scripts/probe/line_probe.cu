// Times sweep_line_kernel (v5) on the bench shape:  ./line_probe [nbatch] [chunk] [nx ny] [tstore]
#include <vector>
#include "../../xlab_ee_fortran_b200/csrc/xee_kernels.cuh"
#include "../../xlab_ee_fortran_b200/csrc/xee_sweep_line.cuh"
using namespace xee;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static Enc enc;
static CUtensorMap mk(const double* base, int nx, int ny, int nb, int bw, int bh) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nb};
  const cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}
__global__ void fill(double* p, size_t n, double s) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = s * (double)((i * 2654435761u) % 1000) / 1000.0; }
__global__ void fillcoe(double* c, size_t nn) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nn; i += (size_t)gridDim.x * blockDim.x)
    for (int k = 0; k < 9; ++k) c[k * nn + i] = k == 4 ? -4.0 : 0.5;
}
static CUtensorMap mk_out(const double* base, int nx, int ny, int nb) {
  CUtensorMap m;
  const cuuint64_t dims[4] = {(cuuint64_t)ln::TW, (cuuint64_t)(nx / ln::TW), (cuuint64_t)ny, (cuuint64_t)nb};
  const cuuint64_t strides[3] = {(cuuint64_t)ln::TW * 8, (cuuint64_t)nx * 8, (cuuint64_t)nx * ny * 8};
  const cuuint32_t box[4] = {(cuuint32_t)ln::Cfg<double>::FW, 1, (cuuint32_t)ln::TH, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode(out) failed %d\n", (int)r); exit(1); }
  return m;
}
int main(int argc, char** argv) {
  const int nb = argc > 1 ? atoi(argv[1]) : 512, chunk = argc > 2 ? atoi(argv[2]) : 32;
  const int nx = argc > 4 ? atoi(argv[3]) : 512, ny = argc > 4 ? atoi(argv[4]) : 256;
  const int tstore = argc > 5 ? atoi(argv[5]) : 1;
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q)); enc = (Enc)fp;
  const size_t nn = (size_t)nx * ny, tot = nn * nb;
  double *x[2], *f, *coe, *fac; unsigned char* pack;
  for (auto& p : x) { CK(cudaMalloc(&p, tot * 8)); fill<<<1024, 256>>>(p, tot, 1.0); }
  CK(cudaMalloc(&f, tot * 8)); fill<<<1024, 256>>>(f, tot, 0.1);
  CK(cudaMalloc(&coe, nn * 10 * 8)); fillcoe<<<256, 256>>>(coe, nn);
  CK(cudaMalloc(&fac, nn * kLineFacPlanes * 8));
  line_factor_kernel<double><<<dim3((nx / 32 + 31) / 32 + 1, ny), 32>>>(coe, fac, nx, ny);
  const int tlx = (nx + ln::TW - 1) / ln::TW, tly = (ny + ln::TH - 1) / ln::TH;
  CK(cudaMalloc(&pack, (size_t)tlx * tly * line_pack_tile_bytes<double>()));
  line_pack_kernel<double><<<tlx * tly, ln::NT>>>(coe, fac, pack, nx, ny, tlx);
  CUtensorMap mxh[2], mxp[2];
  for (int k = 0; k < 2; ++k) { mxh[k] = mk(x[k], nx, ny, nb, ln::Cfg<double>::XW, ln::TH + 2); mxp[k] = mk(x[k], nx, ny, nb, ln::Cfg<double>::FW, ln::TH); }
  CUtensorMap mf = mk(f, nx, ny, nb, ln::Cfg<double>::FW, ln::TH);
  CUtensorMap mo[2] = {mk_out(x[0], nx, ny, nb), mk_out(x[1], nx, ny, nb)};
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  LineArgs<double> A{};
  A.pack = pack; A.pack_set_stride = 0; A.field_stride = (long long)nn; A.nx = nx; A.ny = ny; A.nbatch = nb; A.alpha = 1.0; A.omega = 1.2;
  A.tiles_x = (nx + ln::TW - 1) / ln::TW; A.tiles_y = (ny + ln::TH - 1) / ln::TH;
  A.chunk = chunk; A.nchunks = (nb + chunk - 1) / chunk; A.tstore = tstore;
  const long long units = (long long)A.tiles_x * A.tiles_y * A.nchunks;
  const int grid = (int)std::min<long long>((long long)sms * ln::CTAS_PER_SM, units);
  auto kern = sweep_line_kernel<double, true, false>;
  const int smem = ln::Cfg<double>::SMEM_BYTES;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 20;
  for (int w = 0; w < 3 + reps; ++w) {
    if (w == 3) CK(cudaEventRecord(e0));
    const int src = w & 1;
    A.dst = x[src ^ 1];
    kern<<<grid, ln::NT, smem>>>(A, mxh[src], mxp[src ^ 1], mf, mo[src ^ 1], mf);
  }
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / reps;
  const double alg = (double)nb * (nx - 2) * (ny - 2) * 8.0 * (4 + 11.0 / nb);
  CK(cudaGetLastError());
  printf("tstore=%d NSTAGE=%d nb=%d chunk=%d grid=%d tiles=%dx%d smem=%d : %.1f us/sweep, alg %.0f GB/s\n", tstore, ln::NSTAGE, nb, chunk, grid, A.tiles_x, A.tiles_y, smem, us, alg / (us * 1e-6) / 1e9);
  return 0;
}
