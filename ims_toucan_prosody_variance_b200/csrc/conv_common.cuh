// conv_common.cuh -- geometry shared by the SIMT and tcgen05 implementations of tb200_conv1d
// and by the weight packer.
#pragma once
#include "common.cuh"

namespace tb200 {

constexpr int kTileM = 128;   // output rows (time steps) per tile == TMEM lanes
constexpr int kMaxTaps = 32;

// Everything a conv kernel needs, derived on the host from tb200_conv1d_params.
struct ConvArgs {
  const void* x;
  const int* len_in;
  void* y;
  const void* residual;
  const float* bias;
  const float* alpha;
  const float* beta;
  const void* w;
  long long x_bs, y_bs, r_bs;
  int x_ld, y_ld, r_ld;
  int B, Cin, Cin_pad, Cout, L_in_max;
  int ntaps;
  int tap_off[kMaxTaps];  // A-tile row read by tap j for output row m is  m + tap_off[j]
  int halo_l;             // -min(tap_off)  (>= 0)
  int R;                  // rows of the staged input tile: 128 + max(tap_off) - min(tap_off)
  int up, up_pad;         // ConvTranspose stride u (0 = regular conv) and its padding u/2
  int N_total;            // GEMM N: Cout (regular) or Cout*u (transposed: n = co*u + phase)
  int NT, n_ntiles;       // N tile (multiple of 16, <= 256)
  int KC, n_kchunks;      // input-channel chunk per weight block
  int n_chunks;           // n_ntiles * ntaps * n_kchunks weight blocks, each KC*NT elements
  int chunk_bytes;
  int ring_slots;         // weight blocks resident in smem at once
  int resident;           // all blocks fit: load once per CTA
  int act;
  float slope;
  int x_f16, y_f16, r_f16;
  int out_act;
  float out_alpha, res_beta;
  int accumulate;
  int tiles_per_utt, total_tiles;
  int tmem_cols;
  int a_bytes;            // bytes of one staged input buffer (one channel panel)
  int S;                  // 128-row sub-tiles (accumulators) per CTA tile
  int a_bufs, acc_bufs;   // staged-input / accumulator buffers (pipeline depth)
  int n_panels;           // input-channel panels staged one at a time (1 = all channels at once)
  int aa_fast;            // lane=channel snake staging allowed (alignment / channel-count preconditions)
  int pw_vec;             // vectorised pointwise staging allowed (alignment preconditions)
  int epi_fast;           // plain epilogue allowed (see epilogue_plain)
  int epi_up;             // plain transposed-conv epilogue allowed (see epilogue_up): 1 scalar stores, 2 aligned pairs
  int l2_prefetch;        // next-tile L2 prefetch (tuning knob, off by default)
  long long* trace;       // per-tile clock64() stamps of CTA 0 (TB200_TRACE debugging aid) or nullptr
  int n_prod;             // producer warps (the other worker warps run the epilogue)
};

struct ConvGeom {
  int Cin_pad, N_total, NT, n_ntiles, KC, n_kchunks, n_chunks, ntaps, elem_bytes, epc;
  long long chunk_elems, packed_bytes;
};

// Tiling policy shared by packer and kernels.  precision: TB200_PREC_F16 / TB200_PREC_TF32;
// the SIMT path uses the F32 image with the same ordering (elem 4 bytes, epc 4).
inline ConvGeom conv_geom(int Cin, int Cout, int K, int up, int precision) {
  ConvGeom g;
  g.elem_bytes = (precision == TB200_PREC_F16) ? 2 : 4;
  g.epc = 16 / g.elem_bytes;                 // elements per 16-byte K chunk
  int kstep = 2 * g.epc;                     // UMMA K per instruction (16 fp16 / 8 tf32)
  g.Cin_pad = (Cin + kstep - 1) / kstep * kstep;
  g.N_total = up > 0 ? Cout * up : Cout;
  int n16 = (g.N_total + 15) / 16 * 16;
  g.n_ntiles = (n16 + 255) / 256;
  g.NT = ((n16 + g.n_ntiles - 1) / g.n_ntiles + 15) / 16 * 16;
  g.ntaps = up > 0 ? 2 : K;
  // channel chunk: largest multiple of kstep dividing Cin_pad with KC*NT*elem <= 16 KB
  int best = kstep;
  for (int kc = kstep; kc <= g.Cin_pad; kc += kstep) {
    if (g.Cin_pad % kc) continue;
    if ((long long)kc * g.NT * g.elem_bytes <= 16384) best = kc;
  }
  g.KC = best;
  g.n_kchunks = g.Cin_pad / g.KC;
  g.n_chunks = g.n_ntiles * g.ntaps * g.n_kchunks;
  g.chunk_elems = (long long)g.KC * g.NT;
  g.packed_bytes = g.chunk_elems * g.n_chunks * g.elem_bytes;
  return g;
}

constexpr int kAaRows = 26;     // output rows per warp pass of the anti-aliased prologue
constexpr int kAaScratch = 72;  // floats per scratch line (64 used)

#ifdef __CUDACC__
__device__ __forceinline__ float load_x(const void* x, bool f16, long long idx) {
  return f16 ? __half2float(__ldg(reinterpret_cast<const __half*>(x) + idx))
             : __ldg(reinterpret_cast<const float*>(x) + idx);
}

__device__ __forceinline__ float clamp_f16(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
// fp32 -> fp16 round-to-nearest-even with overflow saturating to +-65504 in the conversion itself
// (F2FP.SATFINITE: one instruction instead of two FMNMX + F2FP; same result as clamp_f16 for every finite input).
__device__ __forceinline__ __half f16_sat(float v) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return __ushort_as_half(h);
}
__device__ __forceinline__ uint32_t f16x2_sat(float lo, float hi) {   // {lo in bits 0..15, hi in bits 16..31}
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ float apply_pointwise(float v, int act, float slope) {
  switch (act) {
    case TB200_ACT_LEAKY_RELU: return v > 0.f ? v : v * slope;
    case TB200_ACT_RELU: return fmaxf(v, 0.f);
    case TB200_ACT_SWISH: return v / (1.f + expf(-v));
    case TB200_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// ---------------------------------------------------------------------------------------------
// producers: stage ACT(x) rows [t_lo, t_lo + R) x channel groups [g0, g0 + ng) of utterance b.
// E = channels per group; st(g - g0, r, v[E]) stores one group-row.  Rows outside [0, len) and
// channels >= Cin are staged as zeros (the conv's zero padding at the utterance's own ends).
// ---------------------------------------------------------------------------------------------
template <int E, typename Store>
__device__ __forceinline__ void stage_pointwise(const ConvArgs& a, int b, int t_lo, int R, int g0, int ng, int len,
                                                const Store& st, int warp, int nwarps, int lane) {
  const int nrb = (R + 31) / 32;
  const long long xb = (long long)b * a.x_bs;
  for (int task = warp; task < ng * nrb; task += nwarps) {
    const int g = task / nrb, rb = task - g * nrb;
    const int r = rb * 32 + lane;
    const int t = t_lo + r;
    const bool valid = (r < R) && (t >= 0) && (t < len);
    float v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = (g0 + g) * E + e;
      float xv = 0.f;
      if (valid && c < a.Cin) xv = apply_pointwise(load_x(a.x, a.x_f16, xb + (long long)c * a.x_ld + t), a.act, a.slope);
      v[e] = xv;
    }
    if (r < R) st(g, r, v);
  }
}

// BigVGAN anti-aliased SnakeBeta (alias_free_torch.Activation1d(SnakeBeta), AMP.py:45-57):
//   u[2t]   = 2 sum_q x[clamp(t-3+q)] f[11-2q],   u[2t+1] = 2 sum_q x[clamp(t-2+q)] f[10-2q]   (q = 0..5)
//   s[m]    = u[m] + 1/(e^beta + 1e-9) sin^2(u[m] e^alpha),        m clamped to [0, 2 len - 1]
//   out[t]  = sum_k f[k] s[2t - 5 + k]                                                  (k = 0..11)
// (replicate pad 5|5 -> 2*conv_transpose(stride 2)[15:-15] -> snake -> replicate pad 5|6 -> conv(stride 2)).
// A warp handles one channel group for 26 consecutive rows per pass: all 32 lanes produce the 64
// s values those 26 outputs need into a per-warp scratch line, then lanes 0..25 run the 12-tap
// down filter.  scratch: nwarps * 2 * kAaScratch floats.
#ifndef TB200_NO_AA_CONSTANT
template <int E, bool kFast, typename Store>
__device__ __forceinline__ void stage_aa_snake(const ConvArgs& a, int b, int t_lo, int R, int g0, int ng, int len,
                                               const Store& st, float* scratch, int warp, int nwarps, int lane) {
  const int nrb = (R + kAaRows - 1) / kAaRows;
  const long long xb = (long long)b * a.x_bs;
  float* sc = scratch + warp * (2 * kAaScratch);
  float f[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) f[k] = c_aa_filter[k];
  int flip = 0;
  for (int task = warp; task < ng * nrb; task += nwarps) {
    const int g = task / nrb, rb = task - g * nrb;
    const int r0 = rb * kAaRows;
    const int tp = t_lo + r0 - 3 + lane;  // time index whose (s[2tp], s[2tp+1]) this lane produces
    const int tc = min(max(tp, 0), len - 1);
    const int r = r0 + lane;              // output row of this lane (lanes < 26)
    const int t = t_lo + r;
    const bool out_lane = lane < kAaRows && r < R;
    const bool valid = out_lane && t >= 0 && t < len;
    const bool interior = (tc - 3 >= 0) && (tc + 3 < len);
    float v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = (g0 + g) * E + e;
      float outv = 0.f;
      if (c < a.Cin) {  // warp-uniform
        const long long row = xb + (long long)c * a.x_ld;
        float xw[7];
        if (interior) {
#pragma unroll
          for (int q = 0; q < 7; ++q) xw[q] = load_x(a.x, a.x_f16, row + tc - 3 + q);
        } else {
#pragma unroll
          for (int q = 0; q < 7; ++q) xw[q] = load_x(a.x, a.x_f16, row + min(max(tc - 3 + q, 0), len - 1));
        }
        float u0 = 0.f, u1 = 0.f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          u0 = fmaf(xw[q], f[11 - 2 * q], u0);
          u1 = fmaf(xw[q + 1], f[10 - 2 * q], u1);
        }
        u0 *= 2.f;
        u1 *= 2.f;
        float s0, s1;
        if (kFast) {
          const float ea = __expf(__ldg(a.alpha + c));
          const float ib = 1.0f / (__expf(__ldg(a.beta + c)) + 1e-9f);
          const float z0 = __sinf(u0 * ea), z1 = __sinf(u1 * ea);
          s0 = fmaf(ib * z0, z0, u0);
          s1 = fmaf(ib * z1, z1, u1);
        } else {
          const float ea = expf(__ldg(a.alpha + c));
          const float ib = 1.0f / (expf(__ldg(a.beta + c)) + 1e-9f);
          const float z0 = sinf(u0 * ea), z1 = sinf(u1 * ea);
          s0 = u0 + ib * (z0 * z0);
          s1 = u1 + ib * (z1 * z1);
        }
        if (tp < 0) s1 = s0;     // m clamped to 0
        if (tp >= len) s0 = s1;  // m clamped to 2 len - 1
        float* line = sc + flip * kAaScratch;
        flip ^= 1;
        *reinterpret_cast<float2*>(line + 2 * lane) = make_float2(s0, s1);
        __syncwarp();
        if (valid) {
          // s[2t-5+k] sits at line[2 lane + 1 + k]  (line[0] <-> m = 2 (t_lo + r0 - 3))
          float sv[14];
#pragma unroll
          for (int q = 0; q < 7; ++q) {
            const float2 p2 = *reinterpret_cast<const float2*>(line + 2 * lane + 2 * q);
            sv[2 * q] = p2.x;
            sv[2 * q + 1] = p2.y;
          }
#pragma unroll
          for (int k = 0; k < 12; ++k) outv = fmaf(f[k], sv[k + 1], outv);
        }
      }
      v[e] = outv;
    }
    if (out_lane) st(g, r, v);
  }
  __syncwarp();
}
#endif  // TB200_NO_AA_CONSTANT

// ---------------------------------------------------------------------------------------------
// epilogue arithmetic shared by both implementations
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float finish(float acc, const ConvArgs& a, int co, long long ridx, long long yidx) {
  float v = acc + (a.bias ? __ldg(a.bias + co) : 0.f);
  if (a.out_act == TB200_OUT_TANH) v = tanhf(v);
  else if (a.out_act == TB200_OUT_RELU) v = fmaxf(v, 0.f);
  v *= a.out_alpha;
  if (a.residual)
    v = fmaf(a.res_beta, a.r_f16 ? __half2float(reinterpret_cast<const __half*>(a.residual)[ridx])
                                 : __ldg(reinterpret_cast<const float*>(a.residual) + ridx), v);
  if (a.accumulate)
    v += a.y_f16 ? __half2float(reinterpret_cast<const __half*>(a.y)[yidx]) : reinterpret_cast<const float*>(a.y)[yidx];
  return v;
}

__device__ __forceinline__ void store_y(const ConvArgs& a, long long yidx, float v) {
  if (a.y_f16) reinterpret_cast<__half*>(a.y)[yidx] = f16_sat(v);
  else reinterpret_cast<float*>(a.y)[yidx] = v;
}
#endif  // __CUDACC__

}  // namespace tb200
