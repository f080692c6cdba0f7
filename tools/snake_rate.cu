// Microbenchmark: throughput of the anti-aliased SnakeBeta streaming filter (snake_stream.cuh) when its input is a
// shared-memory tile (the fused residual-pair kernel's staging) instead of global memory: cycles per element per SM
// as a function of the number of staging warps.  No MMA, no global traffic in the timed loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DTB200_NO_AA_CONSTANT -I ims_toucan_prosody_variance_b200/csrc \
//        tools/snake_rate.cu -o build/snake_rate
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <vector>
#include "snake_stream.cuh"

namespace tb200 {
void set_error(const char*, ...) {}
int fail(int c, const char*, ...) { return c; }
}
using namespace tb200;

// X tile: C rows (channels) of `pitch` elements, time origin tx0 (multiple of 8).  A1 tile: [C/8][R][8] halves.
template <bool XF16, int MAXWARPS>
__global__ void __launch_bounds__(MAXWARPS * 32, 1) snake_kernel(int C, int R, int nwarps_used, int iters, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int esz = XF16 ? 2 : 4;
  const int rows_x = R + 64;
  const int pitch = ((rows_x * esz + 127) / 128 * 128 + 16) / esz;   // bytes = 16 (mod 128): conflict-free lane = channel reads
  uint8_t* X = smem;
  __half* A = reinterpret_cast<__half*>(smem + (size_t)C * pitch * esz);
  // fill X with a smooth deterministic signal
  for (int i = threadIdx.x; i < C * pitch; i += blockDim.x) {
    const float v = __sinf(0.01f * i) * 0.7f;
    if (XF16) reinterpret_cast<__half*>(X)[i] = __float2half(v);
    else reinterpret_cast<float*>(X)[i] = v;
  }
  __syncthreads();
  const int ncb = C / 32;
  // every warp streams rows [0, R) of channel block (warp % ncb): identical work per warp, so the time per tile is the
  // time of one segment of R rows with nwarps_used warps resident (steady throughput vs warm-up overhead separated
  // by varying R)
  const long long t0 = clock64();
  if (warp < nwarps_used) {
    for (int it = 0; it < iters; ++it) {
      {
        const int cb = warp % ncb;
        const int t_lo = 26;
        const int r_beg = 0, r_end = R;
        const int c = cb * 32 + lane;
        const float ea = 1.0f + 0.01f * c, ib = 0.9f;
        __half* dst = A + ((size_t)(c / 8) * R) * 8 + (c % 8);
        // time origin of the X tile: element index = t - tx0 with tx0 = 0 here; the tile starts 16 elements into the row
        aa_channel_task<__half, false, XF16, true>(X, (long long)c * pitch + 16 - 0, ea, ib, t_lo, t_lo + r_beg, t_lo + r_end, 1 << 30, dst);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nwarps_used * 32) : "memory");
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (threadIdx.x == 1 && blockIdx.x == 0) out[1] = (long long)__half_as_ushort(A[5]);
}

template <bool XF16, int MAXWARPS>
static void run(int C, int R, int nw, long long* d) {
  const int esz = XF16 ? 2 : 4;
  const int rows_x = R + 64;
  const int pitch = ((rows_x * esz + 127) / 128 * 128 + 16) / esz;
  const size_t smem = (size_t)C * pitch * esz + (size_t)C * R * 2 + 256;
  if (smem > 227 * 1024) return;
  auto k = snake_kernel<XF16, MAXWARPS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 20;
  k<<<148, MAXWARPS * 32, smem>>>(C, R, nw, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double per_tile = (double)h[0] / iters;
  printf("x %s  C %3d  segment rows %4d  warps %2d (block %2d warps, %3d regs cap): %8.0f cycles/segment  %.3f cycles/element/SM  %s\n",
         XF16 ? "f16" : "f32", C, R, nw, MAXWARPS, 65536 / (MAXWARPS * 32), per_tile, per_tile / ((double)32 * nw * R),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int C = 64;
  for (int R : {48, 96, 192, 384}) {
    run<true, 8>(C, R, 4, d);
    run<true, 8>(C, R, 8, d);
    run<true, 12>(C, R, 12, d);
    run<true, 16>(C, R, 16, d);
    run<true, 20>(C, R, 20, d);
    run<true, 24>(C, R, 24, d);
    run<false, 16>(C, R, 16, d);
  }
  return 0;
}
