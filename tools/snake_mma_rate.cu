// Microbenchmark + numerics check of the tensor-path anti-aliased SnakeBeta (snake_mma.cuh, mma.sync FIRs) against the
// CUDA-core streaming filter (snake_stream.cuh) on the same shared-memory tile: cycles per element per SM for a range
// of warp counts and segment lengths, and the max abs difference of the fp16 operand tiles they write.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DTB200_NO_AA_CONSTANT -I ims_toucan_prosody_variance_b200/csrc -I tools \
//        tools/snake_mma_rate.cu -o build/snake_mma_rate
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cmath>
#include <vector>
#include "snake_mma.cuh"

namespace tb200 {
void set_error(const char*, ...) {}
int fail(int c, const char*, ...) { return c; }
}
using namespace tb200;

constexpr int kLead = 32;   // tile column of time 0

// X tile: C rows of `pitch` halves; A tile: [C/8][R][8] halves.  MODE 0: CUDA-core filter (32 channels per warp),
// MODE 1: mma.sync filter (16 channels per warp).  Every warp runs rows [0, R) of channel block (warp % nblocks).
template <int MODE, int MAXWARPS>
__global__ void __launch_bounds__(MAXWARPS * 32, 1) snake_kernel(int C, int R, int nw, int iters, long long* out, __half* dump) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_x = R + 96;
  const int pitch = ((rows_x * 2 + 127) / 128 * 128 + 16) / 2;
  __half* X = reinterpret_cast<__half*>(smem);
  __half* A = X + (size_t)C * pitch;
  float2* eaib = reinterpret_cast<float2*>(A + (size_t)C * R);
  for (int i = threadIdx.x; i < C * pitch; i += blockDim.x) {
    const int c = i / pitch, t = i % pitch;
    X[i] = __float2half(__sinf(0.05f * t + 0.3f * c) * 0.9f + 0.4f * __sinf(1.3f * t + c));
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) eaib[i] = make_float2(1.0f + 0.02f * i, 0.9f - 0.003f * i);
  for (int i = threadIdx.x; i < C * R; i += blockDim.x) A[i] = __float2half(0.f);
  __syncthreads();
  AaMmaTaps T;
  aa_mma_taps(lane, T);
  const int t_lo = 26;
  const long long t0 = clock64();
  if (warp < nw) {
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {
        const int cb = warp % (C / 32), c = cb * 32 + lane;
        __half* dst = A + ((size_t)(c / 8) * R) * 8 + (c % 8);
        aa_channel_task<__half, false, true, true>(X, (long long)c * pitch + kLead, eaib[c].x, eaib[c].y, t_lo, t_lo, t_lo + R, 1 << 30, dst);
      } else {
        const int cb = warp % (C / 16);
        aa_mma_task(X, pitch, -kLead, eaib, cb * 16, T, t_lo, t_lo, t_lo + R, A, R, lane);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (blockIdx.x == 0 && dump)
    for (int i = threadIdx.x; i < C * R; i += blockDim.x) dump[i] = A[i];
}

template <int MODE, int MAXWARPS>
static double run(int C, int R, int nw, long long* d, __half* dump, std::vector<__half>* host) {
  const int rows_x = R + 96;
  const int pitch = ((rows_x * 2 + 127) / 128 * 128 + 16) / 2;
  const size_t smem = (size_t)C * pitch * 2 + (size_t)C * R * 2 + C * 8 + 256;
  if (smem > 227 * 1024) return 0;
  auto k = snake_kernel<MODE, MAXWARPS>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 20;
  k<<<148, MAXWARPS * 32, smem>>>(C, R, nw, iters, d, dump);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  if (host) { host->resize((size_t)C * R); cudaMemcpy(host->data(), dump, (size_t)C * R * 2, cudaMemcpyDeviceToHost); }
  const double per_seg = (double)h[0] / iters;
  const int ch_per_warp = MODE ? 16 : 32;
  printf("%s  C %3d  rows %4d  warps %2d (block %2d): %8.0f cycles/segment  %.3f cycles/element/SM  %s\n", MODE ? "mma " : "simt", C, R, nw,
         MAXWARPS, per_seg, per_seg / ((double)ch_per_warp * nw * R), e == cudaSuccess ? "" : cudaGetErrorString(e));
  return per_seg;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  __half* dump;
  cudaMalloc(&dump, 128 * 1024 * 2);
  const int C = 64;
  {   // numerics: same tile through both paths
    std::vector<__half> a, b;
    run<0, 8>(C, 384, 8, d, dump, &a);
    run<1, 8>(C, 384, 8, d, dump, &b);
    double mx = 0, rms = 0, ref = 0;
    for (size_t i = 0; i < a.size(); ++i) {
      const double x = __half2float(a[i]), y = __half2float(b[i]);
      mx = fmax(mx, fabs(x - y)); rms += (x - y) * (x - y); ref += x * x;
    }
    printf("numerics: max abs diff %.3e, rel rms %.3e (%.1f dB), rms of reference %.3f, %zu values\n", mx, sqrt(rms / ref),
           -10 * log10(rms / ref), sqrt(ref / a.size()), a.size());
  }
  for (int R : {96, 192, 384}) {
    run<0, 8>(C, R, 8, d, nullptr, nullptr);
    run<1, 8>(C, R, 4, d, nullptr, nullptr);
    run<1, 8>(C, R, 8, d, nullptr, nullptr);
    run<1, 12>(C, R, 12, d, nullptr, nullptr);
    run<1, 16>(C, R, 16, d, nullptr, nullptr);
  }
  return 0;
}
