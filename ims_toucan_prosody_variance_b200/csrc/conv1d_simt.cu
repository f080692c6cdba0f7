// conv1d_simt.cu -- fp32 CUDA-core implementation of tb200_conv1d (TB200_PREC_FP32_SIMT).
//
// Same GEMM view, prologue and epilogue as conv1d_umma.cu, but fp32 FMA on the CUDA cores and raw
// torch-layout weights: the exact-parity mode (mel rel-L1 <= 1e-3 needs more mantissa than fp16
// through 18 flow blocks) and the independent cross-check of the tensor-core path.
//
// CTA tile: 64 time steps x 64 GEMM-N columns, 256 threads, each thread a 4(time) x 4(n) register
// tile; input channels are walked in chunks of 8 with the activated input rows (halo included) and
// the matching weight slab staged in shared memory.
#include "conv_common.cuh"

namespace tb200 {

constexpr int kSimtThreads = 256;
constexpr int kSimtTT = 64;   // time steps per CTA
constexpr int kSimtTN = 64;   // GEMM-N columns per CTA
constexpr int kSimtCK = 8;    // input channels per chunk

struct SimtStore {
  float* base;  // [CK][Rs]
  int Rs;
  __device__ __forceinline__ void operator()(int g, int r, const float (&v)[1]) const { base[g * Rs + r] = v[0]; }
};

// weight of GEMM column n, input channel ci, tap j in the raw torch layout
__device__ __forceinline__ float raw_weight(const ConvArgs& a, const float* w, int n, int ci, int j, int K) {
  if (a.up > 0) {  // (Cin, Cout, 2u): n = co*u + phase; tap 0 -> k = phase, tap 1 -> k = phase + u
    const int co = n / a.up, ph = n - co * a.up;
    return __ldg(w + ((long long)ci * a.Cout + co) * (2 * a.up) + ph + j * a.up);
  }
  return __ldg(w + ((long long)n * a.Cin + ci) * K + j);
}

__global__ void __launch_bounds__(kSimtThreads) conv1d_simt_kernel(const __grid_constant__ ConvArgs a, int K) {
  extern __shared__ float sm[];
  const int Rs = kSimtTT + a.R - kTileM;          // staged rows: 64 + halo
  float* sx = sm;                                  // [CK][Rs]
  float* sw = sx + kSimtCK * Rs;                   // [CK][ntaps][TN]
  float* scratch = sw + kSimtCK * a.ntaps * kSimtTN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_t = a.tiles_per_utt * (kTileM / kSimtTT);
  const int b = blockIdx.x / tiles_t;
  const int t0 = (blockIdx.x - b * tiles_t) * kSimtTT;
  const int n0 = blockIdx.y * kSimtTN;
  const int len = a.len_in ? min(__ldg(a.len_in + b), a.L_in_max) : a.L_in_max;   // never beyond the declared extent
  const int rows = len + (a.up > 0 ? 1 : 0);
  if (t0 >= rows || len <= 0) return;
  const int len_out = a.up > 0 ? len * a.up : len;

  const int tt = (threadIdx.x & 15) * 4;   // time offset of this thread's 4 rows
  const int tn = (threadIdx.x >> 4) * 4;   // n offset of this thread's 4 columns
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* w = reinterpret_cast<const float*>(a.w);
  for (int c0 = 0; c0 < a.Cin; c0 += kSimtCK) {
    __syncthreads();
    SimtStore st{sx, Rs};
    if (a.act == TB200_ACT_AA_SNAKEBETA)
      stage_aa_snake<1, false>(a, b, t0 - a.halo_l, Rs, c0, kSimtCK, len, st, scratch, warp, kSimtThreads / 32, lane);
    else
      stage_pointwise<1>(a, b, t0 - a.halo_l, Rs, c0, kSimtCK, len, st, warp, kSimtThreads / 32, lane);
    for (int i = threadIdx.x; i < kSimtCK * a.ntaps * kSimtTN; i += kSimtThreads) {
      const int nn = i % kSimtTN;
      const int j = (i / kSimtTN) % a.ntaps;
      const int cc = i / (kSimtTN * a.ntaps);
      const int n = n0 + nn, ci = c0 + cc;
      sw[i] = (n < a.N_total && ci < a.Cin) ? raw_weight(a, w, n, ci, j, K) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int cc = 0; cc < kSimtCK; ++cc) {
      const float* xr = sx + cc * Rs + a.halo_l + tt;
      for (int j = 0; j < a.ntaps; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(sw + (cc * a.ntaps + j) * kSimtTN + tn);
        const float* xp = xr + a.tap_off[j];
        const float x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3];
        acc[0][0] = fmaf(x0, wv.x, acc[0][0]); acc[0][1] = fmaf(x0, wv.y, acc[0][1]);
        acc[0][2] = fmaf(x0, wv.z, acc[0][2]); acc[0][3] = fmaf(x0, wv.w, acc[0][3]);
        acc[1][0] = fmaf(x1, wv.x, acc[1][0]); acc[1][1] = fmaf(x1, wv.y, acc[1][1]);
        acc[1][2] = fmaf(x1, wv.z, acc[1][2]); acc[1][3] = fmaf(x1, wv.w, acc[1][3]);
        acc[2][0] = fmaf(x2, wv.x, acc[2][0]); acc[2][1] = fmaf(x2, wv.y, acc[2][1]);
        acc[2][2] = fmaf(x2, wv.z, acc[2][2]); acc[2][3] = fmaf(x2, wv.w, acc[2][3]);
        acc[3][0] = fmaf(x3, wv.x, acc[3][0]); acc[3][1] = fmaf(x3, wv.y, acc[3][1]);
        acc[3][2] = fmaf(x3, wv.z, acc[3][2]); acc[3][3] = fmaf(x3, wv.w, acc[3][3]);
      }
    }
  }

#pragma unroll
  for (int jn = 0; jn < 4; ++jn) {
    const int n = n0 + tn + jn;
    if (n >= a.N_total) continue;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int m = t0 + tt + it;
      int co, t;
      if (a.up > 0) {
        co = n / a.up;
        t = m * a.up + (n - co * a.up) - a.up_pad;
      } else {
        co = n;
        t = m;
      }
      if (t < 0 || t >= len_out || m >= rows) continue;
      const long long yidx = (long long)b * a.y_bs + (long long)co * a.y_ld + t;
      const long long ridx = (long long)b * a.r_bs + (long long)co * a.r_ld + t;
      store_y(a, yidx, finish(acc[it][jn], a, co, ridx, yidx));
    }
  }
}

int fill_conv_args(const tb200_conv1d_params* p, int precision, ConvArgs& a);  // api.cu

int conv1d_simt(const tb200_conv1d_params* p, cudaStream_t stream) {
  ConvArgs a;
  int rc = fill_conv_args(p, TB200_PREC_FP32_SIMT, a);
  if (rc) return rc;
  const int Rs = kSimtTT + a.R - kTileM;
  const int smem = (kSimtCK * Rs + kSimtCK * a.ntaps * kSimtTN + (kSimtThreads / 32) * 2 * kAaScratch) * 4;
  if (smem > 200 * 1024) return fail(TB200_E_NOSMEM, "conv1d(simt): %d bytes of shared memory", smem);
  static int configured_per_dev[kMaxDeviceSlots] = {};   // largest dynamic shared-memory limit set so far, per device
  const int slot = current_device_slot();
  if (slot < 0) return fail(TB200_E_NODEVICE, "conv1d(simt): no current CUDA device");
  int& configured = configured_per_dev[slot];
  if (smem > configured) {
    TB200_CUDA_CHECK(cudaFuncSetAttribute(conv1d_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  dim3 grid(a.B * a.tiles_per_utt * (kTileM / kSimtTT), (a.N_total + kSimtTN - 1) / kSimtTN);
  conv1d_simt_kernel<<<grid, kSimtThreads, smem, stream>>>(a, p->K);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace tb200
