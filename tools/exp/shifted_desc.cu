// Experiment (B200): does tcgen05.mma read a SWIZZLE_128B K-major A tile correctly when the descriptor's start address is
// shifted by whole 128-byte rows (not 1024-byte aligned) and the 8-row groups are 1280 bytes apart (SBO = 10 rows)?
// That is what an implicit 3x3 convolution needs to reuse ONE halo tile (18 x 10 pixels x 64 channels) for all nine taps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/exp/shifted_desc.bin tools/exp/shifted_desc.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gill_b200/csrc/ptx.cuh"
using namespace gb;

constexpr int ROWS = 192;

__global__ void __launch_bounds__(128) exp_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                  float* out, int shift_rows, int sbo) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                      // ROWS x 128 B
  uint8_t* sb = smem + ROWS * 128;         // 64 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar[0], ROWS * 128 + 64 * 128);
    tma_load_2d(sa, &ta, &bar[0], 0, 0);
    tma_load_2d(sb, &tb, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, 64, false, false);
    for (int k = 0; k < 4; ++k)
      umma_f16(tmem, make_smem_desc_sw128(smem_u32(sa) + shift_rows * 128 + k * 32, 16, sbo),
               make_smem_desc_sw128(smem_u32(sb) + k * 32, 16, 1024), idesc, k != 0 ? 1u : 0u);
    umma_commit(&bar[1]);
  }
  __syncwarp();
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  const uint32_t ta_ = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t r0[32], r1[32];
  tmem_ld_32x32b_x32(ta_, r0);
  tmem_ld_32x32b_x32(ta_ + 32, r1);
  tmem_wait_ld();
  for (int c = 0; c < 32; ++c) {
    out[tid * 64 + c] = __uint_as_float(r0[c]);
    out[tid * 64 + 32 + c] = __uint_as_float(r1[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  std::vector<__half> X(ROWS * 64), W(64 * 64);
  srand(1);
  for (auto& v : X) v = __float2half(static_cast<float>(rand() % 7 - 3));
  for (auto& v : W) v = __float2half(static_cast<float>(rand() % 5 - 2));
  __half *dX, *dW;
  float* dO;
  cudaMalloc(&dX, X.size() * 2);
  cudaMalloc(&dW, W.size() * 2);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  {
    cuuint64_t gd[2] = {64, ROWS}, gs[1] = {128};
    cuuint32_t bx[2] = {64, ROWS}, es[2] = {1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dX, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t gd2[2] = {64, 64};
    cuuint32_t bx2[2] = {64, 64};
    CUresult r2 = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dW, gd2, gs, bx2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r, (int)r2); return 1; }
  }
  const int smem_bytes = ROWS * 128 + 64 * 128 + 256 + 1024;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int cases[][2] = {{0, 1024}, {1, 1024}, {3, 1024}, {8, 1024}, {0, 1280}, {1, 1280}, {11, 1280}, {21, 1280}, {5, 2304}};
  std::vector<float> O(128 * 64);
  for (auto& cs : cases) {
    const int shift = cs[0], sbo = cs[1];
    cudaMemset(dO, 0, 128 * 64 * 4);
    exp_kernel<<<1, 128, smem_bytes>>>(ta, tb, dO, shift, sbo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("shift %d sbo %d: CUDA error %s\n", shift, sbo, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < 128; ++m) {
      const int r = shift + (m / 8) * (sbo / 128) + m % 8;
      for (int n = 0; n < 64; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += __half2float(X[r * 64 + k]) * __half2float(W[n * 64 + k]);
        const double d = fabs(ref - O[m * 64 + n]);
        if (d > maxerr) maxerr = d;
        if (d > 1e-3) ++bad;
      }
    }
    printf("shift_rows %2d  SBO %4d B: max |err| %.3g, %d / 8192 outputs wrong -> %s\n", shift, sbo, maxerr, bad,
           bad == 0 ? "EXACT" : "MISMATCH");
  }
  return 0;
}
