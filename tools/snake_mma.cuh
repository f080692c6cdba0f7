// snake_mma.cuh -- BigVGAN's anti-aliased SnakeBeta (alias_free_torch.Activation1d(SnakeBeta): 2x kaiser-sinc up-sampler,
// snake, 2x down-sampler; AMP.py:45-57, Snake.py:56-69 of the reference) with both 12-tap FIRs on the warp-level tensor
// path (mma.sync m16n8k16, fp16 operands, fp32 accumulation) instead of 24 CUDA-core FMAs per element.
//
// Formulation (one warp = 16 channels, streaming along time in blocks of 8 input steps; everything stays in registers):
//   up:    U^T[ch][u]  = X^T[ch][t_in] * Gup[t_in][u]     A = 16 channels x 16 input steps straight from the [c][t] fp16
//                                                         shared-memory tile (one 32-bit load per register), B = constant
//                                                         band matrix: a 16-step window yields 16 up-samples (2 n-blocks)
//   snake: on the 8 accumulator values a thread holds (its channels g, g+8), rounded to fp16
//   down:  Y^T[ch][t]  = S^T[ch][u] * Gdn[u][t]           A = the packed accumulators of four consecutive up-blocks (the
//                                                         C fragment of one MMA is the A fragment of the next), B constant
//   store: 8x8 transposes (movmatrix) turn (channel, time-pair) registers into (time, channel-pair) words of the K-major
//          operand tile.
// The filter taps are split into fp16 hi + lo parts (two MMAs per product): with single fp16 taps the waveform SNR drops
// from 54.8 to 47.9 dB (tests/sim_fir_precision.py); with the split it is the rounding of the signal alone.
// 8 HMMA per 128 elements: 0.125 SM-cycles per element at the measured 2 cycles per HMMA per SM (tools/hmma_rate.cu).
// Interior segments only (every touched input inside the utterance).
//
// MEASURED, NOT ADOPTED (tools/snake_mma_rate.cu, profiles/r2_snake_mma_rate.txt): the results agree with the CUDA-core
// filter to 71.5 dB (the fp16 rounding of the signal), but the stage runs at 0.33-0.41 SM-cycles per element against
// 0.365 for the CUDA-core filter at the same 8 warps: 8 HMMA (64 sub-partition cycles) + 8 MUFU.SIN (64) + 6 packed
// conversions per 128 elements leave no room under the ~160 cycles the CUDA-core version needs.  Kept as a tool.
#pragma once
#include "snake_stream.cuh"

namespace tb200 {

__device__ __forceinline__ void hmma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t v) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) { return f16x2_sat(lo, hi); }

// This lane's constant B fragments: [matrix][hi|lo][register].
struct AaMmaTaps {
  uint32_t up[2][2][2];   // up-sampler: n-block 0 / 1 of a 16-step input window
  uint32_t dn[2][2][2];   // down-sampler: k-step 0 (up-blocks 2i-3, 2i-2) / 1 (up-blocks 2i-1, 2i)
};

__device__ __forceinline__ float aa_tap_rt(int k) {   // runtime index (set-up only)
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) v = (k == i) ? aa_tap(i) : v;
  return v;
}
// Gup[k][n] (block b): coefficient of x[t0 - 4 + k] in up-sample 2 t0 + 8 b + n;  u[2t] = 2 sum_j f[2j+1] x[t+2-j],
// u[2t+1] = 2 sum_j f[2j] x[t+3-j]  (conv_transpose1d(stride 2) of the replicate-padded input, cropped [15:-15]).
__device__ __forceinline__ float aa_gup(int b, int k, int n) {
  const int j = (n >> 1) + ((n & 1) ? 7 : 6) - k + 4 * b;
  if (j < 0 || j > 5) return 0.f;
  return 2.f * aa_tap_rt((n & 1) ? 2 * j : 2 * j + 1);
}
// Gdn[k][n] (k-step s): coefficient of s[v0 + k] in y[T0 + n], y[t] = sum_kk f[kk] s[2t + kk - 5]; v0 = 2 T0 - 8 / + 8.
__device__ __forceinline__ float aa_gdn(int s, int k, int n) {
  const int kk = k - 2 * n + (s ? 13 : -3);
  return aa_tap_rt(kk);
}
__device__ __forceinline__ void aa_mma_taps(int lane, AaMmaTaps& T) {
  const int n = lane >> 2, q = lane & 3;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = 2 * q + 8 * r;
      float v[2][2] = {{aa_gup(m, k, n), aa_gup(m, k + 1, n)}, {aa_gdn(m, k, n), aa_gdn(m, k + 1, n)}};
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const __half h0 = __float2half_rn(v[w][0]), h1 = __float2half_rn(v[w][1]);
        const __half l0 = __float2half_rn(v[w][0] - __half2float(h0)), l1 = __float2half_rn(v[w][1] - __half2float(h1));
        const uint32_t hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        const uint32_t lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
        if (w == 0) { T.up[m][0][r] = hi; T.up[m][1][r] = lo; }
        else { T.dn[m][0][r] = hi; T.dn[m][1][r] = lo; }
      }
    }
}

// One (16-channel block, row segment) task of a warp.  xs: fp16 [c][t] shared-memory tile, element (c, t) at
// xs[c * pitch + t - col0_t] (pitch even, col0_t even); c0: first channel of the block; ea_ib[c] = (e^alpha, 1/(e^beta+1e-9));
// output rows [t_beg, t_end) go to the K-major operand tile dst_tile ([C/8][Rp][8] halves, row 0 <-> time t_lo).
// Reads x on [t_beg' - 12, t_end + 27) (t_beg' = t_beg & ~1): all of it must lie inside the utterance and the tile.
__device__ __forceinline__ void aa_mma_task(const __half* xs, int pitch, int col0_t, const float2* ea_ib, int c0,
                                            const AaMmaTaps& T, int t_lo, int t_beg, int t_end, __half* dst_tile, int Rp,
                                            int lane) {
  const int g = lane >> 2, q = lane & 3;
  const float2 p0 = ea_ib[c0 + g], p1 = ea_ib[c0 + g + 8];
  const int o = (t_beg & ~1) - 16;                                  // input block i covers [o + 8i, o + 8i + 8)
  const int n_it = (t_end - 1 - o) / 8 + 2;                         // last iteration (emits the block holding t_end - 1)
  const uint32_t* r0 = reinterpret_cast<const uint32_t*>(xs + (long long)(c0 + g) * pitch + (o - 4 - col0_t + 2 * q));
  const uint32_t* r1 = reinterpret_cast<const uint32_t*>(xs + (long long)(c0 + g + 8) * pitch + (o - 4 - col0_t + 2 * q));
  // word index w of r0/r1 <-> time o - 4 + 2q + 2w.  Iteration 0 is a dummy (its window starts 8 steps later, so the
  // task reads nothing before t_beg' - 12): it only fills the history.
  uint32_t a0 = r0[4], a1 = r1[4];
  uint32_t hist[5][2];                                              // packed up-blocks 2i-5 .. 2i-1
#pragma unroll
  for (int h = 0; h < 5; ++h) hist[h][0] = hist[h][1] = 0u;
  __half* d0 = dst_tile + ((long long)(c0 >> 3) * Rp - t_lo) * 8 + 2 * q;   // + t * 8: channels c0 + 2q, +1 at time t
  __half* d1 = d0 + (long long)Rp * 8;                                       // channels c0 + 8 + 2q, +1
#pragma unroll 1
  for (int i = 1; i <= n_it; ++i) {
    const uint32_t a2 = r0[4 * i + 4], a3 = r1[4 * i + 4];          // window steps 8..15: times o + 8i + 4 + 2q (+1)
    // eight independent MMAs: the up-sampler of input block i (hi / lo taps, two n-blocks) and the down-sampler of
    // output block i - 2 (up-blocks 2i-5 .. 2i-2, all in the history), so that no MMA waits for another
    float uh[2][4], ul[2][4], yh[2][4], yl[2][4];
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) uh[b][e] = ul[b][e] = yh[b][e] = yl[b][e] = 0.f;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      hmma16816(uh[b], a0, a1, a2, a3, T.up[b][0][0], T.up[b][0][1]);
      hmma16816(ul[b], a0, a1, a2, a3, T.up[b][1][0], T.up[b][1][1]);
    }
    hmma16816(yh[0], hist[0][0], hist[0][1], hist[1][0], hist[1][1], T.dn[0][0][0], T.dn[0][0][1]);
    hmma16816(yh[1], hist[2][0], hist[2][1], hist[3][0], hist[3][1], T.dn[1][0][0], T.dn[1][0][1]);
    hmma16816(yl[0], hist[0][0], hist[0][1], hist[1][0], hist[1][1], T.dn[0][1][0], T.dn[0][1][1]);
    hmma16816(yl[1], hist[2][0], hist[2][1], hist[3][0], hist[3][1], T.dn[1][1][0], T.dn[1][1][1]);
    a0 = a2; a1 = a3;
    uint32_t pk[2][2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      pk[b][0] = pack_f16x2(snake_value(uh[b][0] + ul[b][0], p0.x, p0.y), snake_value(uh[b][1] + ul[b][1], p0.x, p0.y));   // channel g
      pk[b][1] = pack_f16x2(snake_value(uh[b][2] + ul[b][2], p1.x, p1.y), snake_value(uh[b][3] + ul[b][3], p1.x, p1.y));   // channel g + 8
    }
    if (i >= 4) {
      // y[0], y[1]: channel g, times T0 + 2q, +1;  y[2], y[3]: channel g + 8.  Transposed: time T0 + g, channels 2q, 2q + 1.
      float y[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) y[e] = (yh[0][e] + yh[1][e]) + (yl[0][e] + yl[1][e]);
      const uint32_t w0 = movmatrix_trans(pack_f16x2(y[0], y[1]));
      const uint32_t w1 = movmatrix_trans(pack_f16x2(y[2], y[3]));
      const int t = o + 8 * (i - 2) + g;
      if (t >= t_beg && t < t_end) {
        *reinterpret_cast<uint32_t*>(d0 + (long long)t * 8) = w0;
        *reinterpret_cast<uint32_t*>(d1 + (long long)t * 8) = w1;
      }
    }
#pragma unroll
    for (int h = 0; h < 3; ++h) { hist[h][0] = hist[h + 2][0]; hist[h][1] = hist[h + 2][1]; }
    hist[3][0] = pk[0][0]; hist[3][1] = pk[0][1];
    hist[4][0] = pk[1][0]; hist[4][1] = pk[1][1];
  }
}

}  // namespace tb200
