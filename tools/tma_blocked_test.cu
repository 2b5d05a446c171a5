// Does a {8 floats, 4 units, 16 rows} TMA box of the blocked U layout land in shared memory as the same image as a
// row-major [32 columns x 16 rows] box (both SWIZZLE_128B_ATOM_32B)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/tma_blocked_test tools/tma_blocked_test.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include <cstdlib>
#include "../mms_answer_selection_b200/csrc/tc/umma.cuh"
using namespace umma;

__global__ void k(const __grid_constant__ CUtensorMap mb, const __grid_constant__ CUtensorMap mr, float* out, int which) {
  __shared__ __align__(1024) uint8_t sm[4096];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, which == 3 ? 4096 : 2048);
    if (which & 1) tma_load_5d(sm, &mb, &bar, 0, 0, 0, 0, 0);
    if (which & 2) tma_load_5d(sm + 2048, &mr, &bar, 0, 0, 0, 0, 0);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int rows = 64, cols = 64, units = cols / 8, groups = rows / 32;
  std::vector<float> hb(rows * cols), hr(rows * cols);
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) {
    hr[r * cols + c] = r * 100 + c;
    hb[((r / 32) * units + c / 8) * 256 + (r % 32) * 8 + c % 8] = r * 100 + c;
  }
  float *db, *dr, *dout; cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dr, hr.size() * 4); cudaMalloc(&dout, 4096);
  cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap mb, mr;
  { const int order = getenv("ORDER") ? atoi(getenv("ORDER")) : 0, promo = getenv("PROMO") ? atoi(getenv("PROMO")) : 1,
              swz = getenv("SWZ") ? atoi(getenv("SWZ")) : 1;
    cuuint64_t d[5] = {8, (cuuint64_t)units, 32, (cuuint64_t)groups, 1}, s[4] = {1024, 32, (cuuint64_t)units * 1024, (cuuint64_t)groups * units * 1024};
    cuuint32_t b[5] = {8, 4, 16, 1, 1}, e[5] = {1, 1, 1, 1, 1};
    if (order == 1) { d[1] = 32; d[2] = units; s[0] = 32; s[1] = 1024; b[1] = 16; b[2] = 4; }
    CUresult r = ((EncodeFn)fn)(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, db, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                swz == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swz == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode blocked (order %d promo %d swz %d): %d\n", order, promo, swz, (int)r); }
  { cuuint64_t d[5] = {(cuuint64_t)cols, (cuuint64_t)rows, 1, 1, 1}, s[4] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * rows * 4, (cuuint64_t)cols * rows * 4, (cuuint64_t)cols * rows * 4};
    cuuint32_t b[5] = {32, 16, 1, 1, 1}, e[5] = {1, 1, 1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, dr, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rowmajor: %d\n", (int)r); }
  int which = 3;
  if (getenv("WHICH")) which = atoi(getenv("WHICH"));
  k<<<1, 128>>>(mb, mr, dout, which);
  printf("kernel (which=%d): %s\n", which, cudaGetErrorString(cudaDeviceSynchronize()));
  std::vector<float> o(1024); cudaMemcpy(o.data(), dout, 4096, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 512; ++i) if (o[i] != o[512 + i]) ++bad;
  printf("mismatching floats: %d of 512\n", bad);
  for (int r = 0; r < 6; ++r) { printf("row %d blocked :", r); for (int c = 0; c < 32; c += 4) printf(" %6.0f", o[r * 32 + c]); printf("\n");
                                printf("row %d rowmajor:", r); for (int c = 0; c < 32; c += 4) printf(" %6.0f", o[512 + r * 32 + c]); printf("\n"); }
  return 0;
}
