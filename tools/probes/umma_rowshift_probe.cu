// Probe (GPU): can a tcgen05.mma read a SWIZZLE_128B K-major operand tile that TMA wrote, starting at a row that is NOT a
// multiple of 8 (start address + r * 128 B), and what must the descriptor's base-offset field (bits 49-51) hold?
// Needed for sharing one (128 + 2)-row activation tile between the three horizontal taps of a 3x3 convolution.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I ../../deep-exemplar-based-video-colorization_b200/csrc \
//        umma_rowshift_probe.cu -L../../deep-exemplar-based-video-colorization_b200/lib -ldvc -o umma_rowshift_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <vector>

#include "tc_common.cuh"

using namespace dvc;

constexpr int ROWS = 144;  // A rows loaded (box), >= 128 + max shift
constexpr int BN = 64;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                       float* out, int shift, int base_off) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                  // ROWS x 128 B
  uint8_t* sB = smem + ROWS * 128;     // BN x 128 B   (ROWS * 128 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + BN * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1), tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(slot, 64);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    tc::mbar_arrive_expect_tx(bar, ROWS * 128 + BN * 128);
    tc::tma_load_2d(sA, &tmA, bar, 0, 0);
    tc::tma_load_2d(sB, &tmB, bar, 0, 0);
    tc::mbar_wait(bar, 0);
    tc::tc_fence_after();
    constexpr uint32_t IDESC = tc::umma_idesc(0u, 128, BN);
    uint64_t dA = tc::umma_desc_k128(tc::smem_u32(sA) + shift * 128) | ((uint64_t)(base_off & 7) << 49);
    uint64_t dB = tc::umma_desc_k128(tc::smem_u32(sB));
    for (int kk = 0; kk < 4; ++kk) tc::umma_ss<false>(tmem, dA + (uint64_t)(kk * 2), dB + (uint64_t)(kk * 2), IDESC, kk ? 1u : 0u);
    tc::umma_commit(done);
  }
  __syncwarp();
  tc::mbar_wait(done, 0);
  tc::tc_fence_after();
  for (int c = 0; c < BN / 32; ++c) {
    uint32_t r[32];
    tc::tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    tc::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * BN + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> hA(ROWS * 64), hB(BN * 64);
  std::vector<float> fA(ROWS * 64), fB(BN * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) fA[i] = (float)((rand() % 17) - 8), hA[i] = __float2half(fA[i]);
  for (size_t i = 0; i < hB.size(); ++i) fB[i] = (float)((rand() % 9) - 4), hB[i] = __float2half(fB[i]);
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2), cudaMalloc(&dB, hB.size() * 2), cudaMalloc(&dO, 128 * BN * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (encode_tmap_2d(&tmA, dA, ROWS, 64, ROWS, 64, 2) || encode_tmap_2d(&tmB, dB, BN, 64, BN, 64, 2)) {
    printf("tensor map encode failed\n");
    return 1;
  }
  const int smem = ROWS * 128 + BN * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(128 * BN);
  for (int shift = 0; shift <= 10; ++shift)
    for (int bo = 0; bo < 8; ++bo) {
      if (!(bo == 0 || bo == (shift & 7) || bo == ((8 - (shift & 7)) & 7))) continue;
      cudaMemset(dO, 0, 128 * BN * 4);
      probe_kernel<<<1, 128, smem>>>(tmA, tmB, dO, shift, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("shift %d base_offset %d: CUDA error %s\n", shift, bo, cudaGetErrorString(e));
        return 2;
      }
      cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < BN; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += fA[(m + shift) * 64 + k] * fB[n * 64 + k];
          if (fabsf(ref - out[m * BN + n]) > 1e-3f) ++bad;
        }
      printf("shift %2d rows, base_offset %d: %s (%d / %d wrong)\n", shift, bo, bad ? "MISMATCH" : "exact", bad, 128 * BN);
    }
  return 0;
}
