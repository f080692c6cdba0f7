// Microbenchmark: cycles per tcgen05.mma (M=128, N, K=16, fp16, SWIZZLE_NONE K-major descriptors as used by
// conv1d_umma.cu) when one thread issues a long back-to-back chain.  nvcc -gencode arch=compute_100a,code=sm_100a
//   -I ims_toucan_prosody_variance_b200/csrc tools/umma_rate.cu -o gpurun_out/umma_rate
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include "common.cuh"

namespace tb200 {
void set_error(const char*, ...) {}
int fail(int c, const char*, ...) { return c; }
}
using namespace tb200;

__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int rows_a, int distinct, int nacc, int swz, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_instr_desc(N, false);
    const uint32_t lbo_a = rows_a * 16, lbo_b = N * 16;
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    // swz: 0 = SWIZZLE_NONE (SBO 128 B), 1 = SWIZZLE_128B (layout_type 2 in bits 61..63, SBO 1024 B); timing only
    const uint32_t hi = swz ? (smem_desc_hi(1024) | (2u << 29)) : smem_desc_hi(128);
    // 8 precomputed descriptor pairs, fully unrolled body: 8 MMAs per loop trip, ~2 instructions per MMA
    uint32_t al[8], bl[8], dc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      al[j] = swz ? smem_desc_lo(a0 + j * 1024, 16) : smem_desc_lo(a0 + j * 16, lbo_a);
      bl[j] = swz ? smem_desc_lo(b0, 16) : smem_desc_lo(b0, lbo_b);
      dc[j] = tmem_base + (uint32_t)((j % nacc) * N);
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) umma_ss_lohi<false>(dc[j], al[j], hi, bl[j], hi, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 2000;
  for (int swz : {0, 1}) {
    for (int N : {32, 64, 128, 256}) {
      for (int nacc : {1, 2, 8}) {
        if (nacc * N > 512) continue;
        rate_kernel<<<148, 128, 96 * 1024>>>(N, iters, 128, 8, nacc, swz, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("swizzle %d  N %3d  accumulators %d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal math %d)  %s\n", swz, N, nacc,
               (double)h[0] / iters, (double)h[1] / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  }
  return 0;
}
