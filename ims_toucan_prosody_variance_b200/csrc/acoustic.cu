// acoustic.cu -- the non-GEMM kernels of the ToucanTTS acoustic model (Conformer encoder/decoder,
// variance predictors, PostNet, Glow PostFlow).  Everything here is fp32 and works on NCL tensors
// (x[b][c][t], t contiguous) with per-utterance lengths: positions >= len[b] are never read as data
// and never written, so a batched call equals per-utterance batch-1 reference calls.
// The dense contractions of the same modules run through tb200_conv1d (conv1d_umma.cu).
#include "common.cuh"

namespace tb200 {

__device__ __forceinline__ int utt_len(const int* len, int b, int L_max) {
  return len ? min(__ldg(len + b), L_max) : L_max;
}

// ---------------------------------------------------------------------------------------------
// channel_norm: LayerNorm over channels (LayerNorm.py:17, eps 1e-12; biased variance) or
// ConditionalLayerNorm (ConditionalLayerNorm.py:52-67: (x-mean)/VARIANCE, no epsilon, scale/bias per
// utterance).  grid (time tiles of 32, B), block (32, 8): lane = time (coalesced), the 8 warps split
// the channels; values stay in registers between the mean and the variance pass (two-pass, like
// torch).  C <= 8 * kCnMaxPerWarp.
// ---------------------------------------------------------------------------------------------
constexpr int kCnWarps = 8;
constexpr int kCnMaxPerWarp = 32;

__global__ void __launch_bounds__(32 * kCnWarps) channel_norm_kernel(
    const float* __restrict__ x, long long x_bs, int x_ld, float* __restrict__ y, long long y_bs, int y_ld,
    const int* __restrict__ len, int C, int L_max, const float* __restrict__ gamma, const float* __restrict__ beta,
    long long gb_bs, int mode, float eps) {
  __shared__ float red[kCnWarps][33];
  const int b = blockIdx.y;
  const int L = utt_len(len, b, L_max);
  const int t0 = blockIdx.x * 32;
  if (t0 >= L) return;
  const int lane = threadIdx.x, w = threadIdx.y;
  const int t = t0 + lane;
  const bool ok = t < L;
  const float* xb = x + (long long)b * x_bs + (ok ? t : t0);
  float v[kCnMaxPerWarp];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kCnMaxPerWarp; ++i) {
    const int c = w + i * kCnWarps;
    v[i] = (c < C) ? __ldg(xb + (long long)c * x_ld) : 0.f;
    s += v[i];
  }
  red[w][lane] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kCnWarps; ++i) tot += red[i][lane];
  const float mean = tot / (float)C;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kCnMaxPerWarp; ++i) {
    const int c = w + i * kCnWarps;
    const float d = v[i] - mean;
    if (c < C) q = fmaf(d, d, q);
  }
  red[w][lane] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int i = 0; i < kCnWarps; ++i) qt += red[i][lane];
  const float var = qt / (float)C;
  const float inv = (mode == 0) ? rsqrtf(var + eps) : 1.0f / var;
  if (!ok) return;
  const float* g = gamma + (long long)b * gb_bs;
  const float* be = beta + (long long)b * gb_bs;
  float* yb = y + (long long)b * y_bs + t;
#pragma unroll
  for (int i = 0; i < kCnMaxPerWarp; ++i) {
    const int c = w + i * kCnWarps;
    if (c < C) yb[(long long)c * y_ld] = fmaf((v[i] - mean) * inv, __ldg(g + c), __ldg(be + c));
  }
}

// ---------------------------------------------------------------------------------------------
// group_norm (PostNet.py:45,55: GroupNorm(32,256) / GroupNorm(20,80), eps 1e-5) over
// (channels of the group) x (the utterance's own frames), then optional tanh and residual add
// (`mel + postnet(mel)`, InferenceToucanTTS.py:241).  grid (groups, B), two-pass statistics.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256) group_norm_kernel(
    const float* __restrict__ x, long long x_bs, int x_ld, float* __restrict__ y, long long y_bs, int y_ld,
    const float* __restrict__ residual, long long r_bs, int r_ld, const int* __restrict__ len, int C, int L_max,
    int groups, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int out_act) {
  __shared__ float red[8];
  const int b = blockIdx.y, g = blockIdx.x;
  const int L = utt_len(len, b, L_max);
  if (L <= 0) return;
  const int cpg = C / groups;
  const float* xb = x + (long long)b * x_bs + (long long)g * cpg * x_ld;
  const int n = cpg * L;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int c = i / L, t = i - c * L;
    s += __ldg(xb + (long long)c * x_ld + t);
  }
  const float mean = block_sum_256(s, red) / (float)n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int c = i / L, t = i - c * L;
    const float d = __ldg(xb + (long long)c * x_ld + t) - mean;
    q = fmaf(d, d, q);
  }
  const float var = block_sum_256(q, red) / (float)n;
  const float inv = rsqrtf(var + eps);
  float* yb = y + (long long)b * y_bs + (long long)g * cpg * y_ld;
  const float* rb = residual ? residual + (long long)b * r_bs + (long long)g * cpg * r_ld : nullptr;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int c = i / L, t = i - c * L;
    const int ch = g * cpg + c;
    float v = fmaf((__ldg(xb + (long long)c * x_ld + t) - mean) * inv, __ldg(gamma + ch), __ldg(beta + ch));
    if (out_act == TB200_OUT_TANH) v = tanhf(v);
    if (rb) v += __ldg(rb + (long long)c * r_ld + t);
    yb[(long long)c * y_ld + t] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// glu_dwconv: the middle of the Conformer convolution module (Convolution.py:43-52):
//   h = a * sigmoid(g)            a = x[:, :C], g = x[:, C:]     (GLU over channels)
//   h = depthwise_conv_k(h) + bias   zero padded at the utterance's own ends
//   h = BatchNorm1d(eval)(h);  y = h * sigmoid(h)                (Swish)
// grid (time tiles of 256, C, B), block 256.
// ---------------------------------------------------------------------------------------------
constexpr int kDwThreads = 256;
constexpr int kDwPer = 4;                       // consecutive outputs per thread (register sliding window)
constexpr int kDwTile = kDwThreads * kDwPer;    // 1024 time steps per CTA
constexpr int kDwMaxK = 63;

__global__ void __launch_bounds__(kDwThreads) glu_dwconv_kernel(
    const float* __restrict__ x, long long x_bs, int x_ld, float* __restrict__ y, long long y_bs, int y_ld,
    const int* __restrict__ len, int C, int L_max, const float* __restrict__ w, const float* __restrict__ bias, int K,
    const float* __restrict__ bn_mean, const float* __restrict__ bn_var, const float* __restrict__ bn_gamma,
    const float* __restrict__ bn_beta, float bn_eps) {
  __shared__ __align__(16) float h[kDwTile + kDwMaxK + 9];
  __shared__ float wk[kDwMaxK + 1];
  const int b = blockIdx.z, c = blockIdx.y;
  const int L = utt_len(len, b, L_max);
  const int t0 = blockIdx.x * kDwTile;
  if (t0 >= L) return;
  const int pad = (K - 1) / 2;
  const float* xa = x + (long long)b * x_bs + (long long)c * x_ld;
  const float* xg = xa + (long long)C * x_ld;
  for (int i = threadIdx.x; i < kDwTile + kDwMaxK + 9; i += kDwThreads) {
    const int t = t0 - pad + i;
    float v = 0.f;
    if (i < kDwTile + K - 1 && t >= 0 && t < L) {
      const float a = __ldg(xa + t), g = __ldg(xg + t);
      v = a * (1.0f / (1.0f + expf(-g)));
    }
    h[i] = v;
  }
  if (threadIdx.x <= kDwMaxK) wk[threadIdx.x] = threadIdx.x < K ? __ldg(w + (long long)c * K + threadIdx.x) : 0.f;  // zero padded
  __syncthreads();
  const int o0 = threadIdx.x * kDwPer;            // first output of this thread inside the tile
  if (t0 + o0 >= L) return;
  float acc[kDwPer];
#pragma unroll
  for (int i = 0; i < kDwPer; ++i) acc[i] = 0.f;
  // 8-value register window over h[o0 + 4q .. o0 + 4q + 7]: four taps per 16-byte shared-memory load (conflict
  // free: consecutive lanes read consecutive 16-byte groups), 16 FMAs per load
  float4 lo = *reinterpret_cast<const float4*>(h + o0);
  for (int q = 0; q * 4 < K; ++q) {
    const float4 hi = *reinterpret_cast<const float4*>(h + o0 + 4 * q + 4);
    const float win[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float wj = wk[4 * q + jj];
#pragma unroll
      for (int i = 0; i < kDwPer; ++i) acc[i] = fmaf(wj, win[i + jj], acc[i]);
    }
    lo = hi;
  }
  const float bc = __ldg(bias + c), mu = __ldg(bn_mean + c), rs = rsqrtf(__ldg(bn_var + c) + bn_eps);
  const float ga = __ldg(bn_gamma + c), be = __ldg(bn_beta + c);
  float* yo = y + (long long)b * y_bs + (long long)c * y_ld + t0 + o0;
#pragma unroll
  for (int i = 0; i < kDwPer; ++i) {
    if (t0 + o0 + i < L) {
      const float n = (acc[i] + bc - mu) * rs * ga + be;
      yo[i] = n * (1.0f / (1.0f + expf(-n)));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// relpos_attention: RelPositionMultiHeadedAttention core (Attention.py:159-198, rel_shift
// :138-157, forward_attention :66-92), flash style: the (T, 2T-1) score tensors never exist.
//   score(i,j) = ((q_i + u_h) . k_j + (q_i + v_h) . p_{i-j}) / sqrt(dk),  j < len[b]
//   out_i = sum_j softmax_j(score(i,.)) v_j
// p_r = linear_pos(PE(r)) for relative position r = i - j lives in column (pos_center - r) of `pos`
// (rel_shift maps bd[i][T-1-i+j] and table row k <-> relative position T-1-k: together r = i - j).
// grid (query tiles of 64, H, B), block 256: thread (ty,tx) owns scores (ty+16a, tx+16b), a,b<4.
// ---------------------------------------------------------------------------------------------
constexpr int kAtQ = 64, kAtK = 64;

template <int DK>
struct AttnSmem {
  float qu[DK][kAtQ];
  float qv[DK][kAtQ];
  float k[DK][kAtK];
  float p[DK][2 * kAtK];
  float vt[kAtK][DK + 1];
  float s[kAtQ][kAtK + 1];
};

template <int DK>
__global__ void __launch_bounds__(256) relpos_attention_kernel(
    const float* __restrict__ qkv, long long qkv_bs, int qkv_ld, const float* __restrict__ pos, int pos_ld,
    int pos_center, int pos_cols, const float* __restrict__ bias_u, const float* __restrict__ bias_v,
    const int* __restrict__ len, int H, int L_max, float* __restrict__ out, long long out_bs, int out_ld) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  AttnSmem<DK>& sm = *reinterpret_cast<AttnSmem<DK>*>(smem_raw);
  constexpr int E = DK / 16;  // output dims per thread
  const int b = blockIdx.z, hh = blockIdx.y;
  const int L = utt_len(len, b, L_max);
  const int i0 = blockIdx.x * kAtQ;
  if (i0 >= L) return;
  const int D = H * DK;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* qb = qkv + (long long)b * qkv_bs + (long long)(hh * DK) * qkv_ld;
  const float* kb = qb + (long long)D * qkv_ld;
  const float* vb = kb + (long long)D * qkv_ld;
  const float* pb = pos + (long long)(hh * DK) * pos_ld;
  const float scale = rsqrtf((float)DK);

  // queries (+ biases), staged once
  for (int i = tid; i < DK * kAtQ; i += 256) {
    const int d = i / kAtQ, r = i - d * kAtQ;
    const int t = i0 + r;
    const float q = (t < L) ? __ldg(qb + (long long)d * qkv_ld + t) : 0.f;
    sm.qu[d][r] = q + __ldg(bias_u + hh * DK + d);
    sm.qv[d][r] = q + __ldg(bias_v + hh * DK + d);
  }

  float m_run[4], l_run[4], o[4][E];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) o[a][e] = 0.f;
  }

  for (int j0 = 0; j0 < L; j0 += kAtK) {
    __syncthreads();  // previous tile fully consumed (also orders the query staging before first use)
    for (int i = tid; i < DK * kAtK; i += 256) {
      const int d = i / kAtK, r = i - d * kAtK;
      const int t = j0 + r;
      const bool ok = t < L;
      sm.k[d][r] = ok ? __ldg(kb + (long long)d * qkv_ld + t) : 0.f;
      sm.vt[r][d] = ok ? __ldg(vb + (long long)d * qkv_ld + t) : 0.f;
    }
    // band of relative positions: local column c <-> r = (i0 - j0) + 63 - c
    const int col_lo = pos_center - (i0 - j0) - (kAtK - 1);
    for (int i = tid; i < DK * 2 * kAtK; i += 256) {
      const int d = i / (2 * kAtK), c = i - d * (2 * kAtK);
      const int col = col_lo + c;
      sm.p[d][c] = (c < 2 * kAtK - 1 && col >= 0 && col < pos_cols) ? __ldg(pb + (long long)d * pos_ld + col) : 0.f;
    }
    __syncthreads();

    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    const int pbase = (kAtK - 1) - ty + tx;  // + 16 (b - a)
#pragma unroll 4
    for (int d = 0; d < DK; ++d) {
      float qu[4], qv[4], kk[4], pp[7];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        qu[a] = sm.qu[d][ty + 16 * a];
        qv[a] = sm.qv[d][ty + 16 * a];
        kk[a] = sm.k[d][tx + 16 * a];
      }
#pragma unroll
      for (int z = 0; z < 7; ++z) pp[z] = sm.p[d][pbase + 16 * (z - 3)];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(qu[a], kk[c], fmaf(qv[a], pp[c - a + 3], acc[a][c]));
    }

    // online softmax over this key tile; a row's 64 scores sit in the 16 lanes sharing ty
    float corr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool ok = (j0 + tx + 16 * c) < L;
        acc[a][c] = ok ? acc[a][c] * scale : -INFINITY;
        mx = fmaxf(mx, acc[a][c]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[a], mx);  // finite: every tile has at least one valid key
      corr[a] = __expf(m_run[a] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float pexp = __expf(acc[a][c] - m_new);
        sm.s[ty + 16 * a][tx + 16 * c] = pexp;
        rs += pexp;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[a] = l_run[a] * corr[a] + rs;
      m_run[a] = m_new;
#pragma unroll
      for (int e = 0; e < E; ++e) o[a][e] *= corr[a];
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < kAtK; ++j) {
      float pr[4], vv[E];
#pragma unroll
      for (int a = 0; a < 4; ++a) pr[a] = sm.s[ty + 16 * a][j];
#pragma unroll
      for (int e = 0; e < E; ++e) vv[e] = sm.vt[j][tx + 16 * e];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int e = 0; e < E; ++e) o[a][e] = fmaf(pr[a], vv[e], o[a][e]);
    }
  }

  // normalise and write: stage through smem so the NCL store is coalesced along time
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float inv = 1.0f / l_run[a];
#pragma unroll
    for (int e = 0; e < E; ++e) sm.qu[tx + 16 * e][ty + 16 * a] = o[a][e] * inv;
  }
  __syncthreads();
  float* ob = out + (long long)b * out_bs + (long long)(hh * DK) * out_ld;
  for (int i = tid; i < DK * kAtQ; i += 256) {
    const int d = i / kAtQ, r = i - d * kAtQ;
    const int t = i0 + r;
    if (t < L) ob[(long long)d * out_ld + t] = sm.qu[d][r];
  }
}

// ---------------------------------------------------------------------------------------------
// small elementwise / layout kernels
// ---------------------------------------------------------------------------------------------

// y[b][c][t] = (x[b][c][t] (or 0) + vec[b][c] (or 0)) * scale     t < len[b]
__global__ void rowvec_affine_kernel(const float* __restrict__ x, long long x_bs, int x_ld, float* __restrict__ y,
                                     long long y_bs, int y_ld, const int* __restrict__ len, int C, int L_max,
                                     const float* __restrict__ vec, long long vec_bs, float scale) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int L = utt_len(len, b, L_max);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  float v = x ? __ldg(x + (long long)b * x_bs + (long long)c * x_ld + t) : 0.f;
  if (vec) v += __ldg(vec + (long long)b * vec_bs + c);
  y[(long long)b * y_bs + (long long)c * y_ld + t] = v * scale;
}

// (B, L, C) row-major  <->  (B, C, L) NCL, through a 32x32 smem tile.  to_ncl: in is (B,L,C).
__global__ void transpose_kernel(const float* __restrict__ in, long long in_bs, int in_ld, float* __restrict__ out,
                                 long long out_bs, int out_ld, const int* __restrict__ len, int C, int L_max, int to_ncl) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int L = utt_len(len, b, L_max);
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (t0 >= L) return;
  const int lx = threadIdx.x, ly = threadIdx.y;  // block (32, 8)
  if (to_ncl) {
    // in[b][t][c]: c contiguous
    for (int r = ly; r < 32; r += 8) {
      const int t = t0 + r, c = c0 + lx;
      tile[r][lx] = (t < L && c < C) ? __ldg(in + (long long)b * in_bs + (long long)t * in_ld + c) : 0.f;
    }
    __syncthreads();
    for (int r = ly; r < 32; r += 8) {
      const int c = c0 + r, t = t0 + lx;
      if (t < L && c < C) out[(long long)b * out_bs + (long long)c * out_ld + t] = tile[lx][r];
    }
  } else {
    for (int r = ly; r < 32; r += 8) {
      const int c = c0 + r, t = t0 + lx;
      tile[r][lx] = (t < L && c < C) ? __ldg(in + (long long)b * in_bs + (long long)c * in_ld + t) : 0.f;
    }
    __syncthreads();
    for (int r = ly; r < 32; r += 8) {
      const int t = t0 + r, c = c0 + lx;
      if (t < L && c < C) out[(long long)b * out_bs + (long long)t * out_ld + c] = tile[lx][r];
    }
  }
}

// glow_utils.py:28-53 with n_sqz = 2 on NCL tensors.
//   squeeze:   y[b][s*C + c][tau] = x[b][c][2 tau + s],  tau < len[b]/2   (odd tail dropped)
//   unsqueeze: y[b][c][2 tau + s] = x[b][s*C + c][tau],  tau < len2[b]    (len = squeezed lengths)
__global__ void squeeze2_kernel(const float* __restrict__ x, long long x_bs, int x_ld, float* __restrict__ y,
                                long long y_bs, int y_ld, const int* __restrict__ len, int C, int L_max, int inverse) {
  // a thread moves 4 consecutive positions of the UNSQUEEZED axis = 2 positions of each squeezed row:
  // 16-byte accesses on the unsqueezed side, 8-byte accesses on the two squeezed rows, all coalesced
  const int b = blockIdx.z, c = blockIdx.y;
  const int L = utt_len(len, b, L_max);
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int n = inverse ? 2 * L : 2 * (L / 2);   // valid unsqueezed positions (len = squeezed / unsqueezed lengths)
  if (i >= n) return;
  const long long un = (long long)b * (inverse ? y_bs : x_bs) + (long long)c * (inverse ? y_ld : x_ld) + i;
  const long long s0 = (long long)b * (inverse ? x_bs : y_bs) + (long long)c * (inverse ? x_ld : y_ld) + (i >> 1);
  const long long s1 = s0 + (long long)C * (inverse ? x_ld : y_ld);
  const bool full = i + 4 <= n && (((inverse ? y_ld : x_ld) | (inverse ? x_ld : y_ld)) & 3) == 0 &&
                    (((inverse ? y_bs : x_bs) | (inverse ? x_bs : y_bs)) & 3) == 0;
  if (!inverse) {
    if (full && ((reinterpret_cast<uintptr_t>(x + un) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y + s0) & 7) == 0) &&
        ((reinterpret_cast<uintptr_t>(y + s1) & 7) == 0)) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + un));
      *reinterpret_cast<float2*>(y + s0) = make_float2(v.x, v.z);
      *reinterpret_cast<float2*>(y + s1) = make_float2(v.y, v.w);
    } else {
      for (int k = 0; k < 4 && i + k < n; ++k) y[((k & 1) ? s1 : s0) + (k >> 1)] = __ldg(x + un + k);
    }
  } else {
    if (full && ((reinterpret_cast<uintptr_t>(y + un) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x + s0) & 7) == 0) &&
        ((reinterpret_cast<uintptr_t>(x + s1) & 7) == 0)) {
      const float2 e = __ldg(reinterpret_cast<const float2*>(x + s0)), o = __ldg(reinterpret_cast<const float2*>(x + s1));
      *reinterpret_cast<float4*>(y + un) = make_float4(e.x, o.x, e.y, o.y);
    } else {
      for (int k = 0; k < 4 && i + k < n; ++k) y[un + k] = __ldg(x + ((k & 1) ? s1 : s0) + (k >> 1));
    }
  }
}

// WaveNet gate (wavenet.py:29-35, 102-111): y[c] = tanh(a[c]) * sigmoid(a[c + H]).
// VEC = 4: one 16-byte load per operand row and one 16-byte store per thread (pitches and bases 16-byte aligned).
template <int VEC>
__global__ void wn_gate_kernel(const float* __restrict__ a, long long a_bs, int a_ld, float* __restrict__ y, long long y_bs,
                               int y_ld, const int* __restrict__ len, int Hc, int L_max) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int L = utt_len(len, b, L_max);
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (t >= L) return;
  const float* ta = a + (long long)b * a_bs + (long long)c * a_ld + t;
  const float* sa = ta + (long long)Hc * a_ld;
  float* yo = y + (long long)b * y_bs + (long long)c * y_ld + t;
  if (VEC == 4 && t + 4 <= L) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(ta)), g = __ldg(reinterpret_cast<const float4*>(sa));
    float4 o;
    o.x = tanhf(u.x) * (1.0f / (1.0f + expf(-g.x)));
    o.y = tanhf(u.y) * (1.0f / (1.0f + expf(-g.y)));
    o.z = tanhf(u.z) * (1.0f / (1.0f + expf(-g.z)));
    o.w = tanhf(u.w) * (1.0f / (1.0f + expf(-g.w)));
    *reinterpret_cast<float4*>(yo) = o;
  } else {
    for (int i = 0; i < VEC && t + i < L; ++i) yo[i] = tanhf(__ldg(ta + i)) * (1.0f / (1.0f + expf(-__ldg(sa + i))));
  }
}

// The three elementwise steps that close one reversed flow block (Glow.py:260-263, 116-128, 30-32):
//   coupling^-1 : x1 <- (x1 - m) * exp(-logs)          (m, logs) = end(wn(...)), x1 = x[C/2:]
//   invconv^-1  : channel a*(C/2) + 2 m + r  <->  group q = 2a + r of 4;  x_p <- sum_q Winv[p][q] x_q
//   actnorm^-1  : x <- (x - bias) * exp(-logs_an)
// in place on x (B, C, L2); one thread per (m, t) touches exactly its 4 channels.
__global__ void flow_close_kernel(float* __restrict__ x, long long x_bs, int x_ld, const float* __restrict__ ml,
                                  long long ml_bs, int ml_ld, const int* __restrict__ len, int C, int L_max,
                                  const float* __restrict__ w_inv, const float* __restrict__ an_bias,
                                  const float* __restrict__ an_logs) {
  const int b = blockIdx.z, m = blockIdx.y;
  const int L = utt_len(len, b, L_max);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  const int half = C / 2;
  float* xb = x + (long long)b * x_bs + t;
  const float* mb = ml + (long long)b * ml_bs + t;
  float v[4];
  int ch[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int a = q >> 1, r = q & 1;
    ch[q] = a * half + 2 * m + r;
    v[q] = xb[(long long)ch[q] * x_ld];
    if (a == 1) {
      const int k = 2 * m + r;
      const float mean = __ldg(mb + (long long)k * ml_ld);
      const float logs = __ldg(mb + (long long)(half + k) * ml_ld);
      v[q] = (v[q] - mean) * expf(-logs);
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc = fmaf(__ldg(w_inv + p * 4 + q), v[q], acc);
    const int c = ch[p];
    xb[(long long)c * x_ld] = (acc - __ldg(an_bias + c)) * expf(-__ldg(an_logs + c));
  }
}

// F.normalize(x, dim=1) on (B, C): x / max(||x||_2, 1e-12).  One warp per row.
__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = x[(long long)b * C + c];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  const float d = fmaxf(sqrtf(s), 1e-12f);
  for (int c = lane; c < C; c += 32) y[(long long)b * C + c] = x[(long long)b * C + c] / d;
}

// The conditioning MLPs of ConditionalLayerNorm (ConditionalLayerNorm.py:27-50):
//   out[n][b] = W4 tanh(W2 tanh(W0 e_b + b0) + b2) + b4     (E -> E -> Cc -> Cc), n = 0..N-1 stacked MLPs.
// grid (N, ceil(B / kClnBT)), block 256: a block runs one MLP for kClnBT utterances, so every weight row is read once per
// kClnBT utterances; a warp owns output rows (lanes read consecutive columns: coalesced 128-byte requests, then a
// butterfly sum).  fp32 dot products, exact tanhf; the summation order differs from a sequential dot product only by
// the lane interleave (|error| ~ 1e-7 relative).
constexpr int kClnBT = 8;

__device__ __forceinline__ void cln_layer(const float* __restrict__ W, const float* __restrict__ bias, int n_out, int n_in,
                                          const float* __restrict__ in_s, int in_pitch, float* __restrict__ out_s, int out_pitch,
                                          float* __restrict__ out_g, long long out_g_stride, int nb, bool act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o = warp; o < n_out; o += nwarps) {
    float acc[kClnBT];
#pragma unroll
    for (int j = 0; j < kClnBT; ++j) acc[j] = 0.f;
    const float* row = W + (long long)o * n_in;
    for (int i = lane; i < n_in; i += 32) {
      const float w = __ldg(row + i);
#pragma unroll
      for (int j = 0; j < kClnBT; ++j) acc[j] = fmaf(w, in_s[j * in_pitch + i], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kClnBT; ++j) acc[j] = warp_sum(acc[j]);
    if (lane < kClnBT && lane < nb) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < kClnBT; ++j) v = (lane == j) ? acc[j] : v;
      v += __ldg(bias + o);
      if (act) v = tanhf(v);
      if (out_s) out_s[lane * out_pitch + o] = v;
      else out_g[(long long)lane * out_g_stride + o] = v;
    }
  }
}

__global__ void __launch_bounds__(256) cln_mlp_kernel(const float* __restrict__ e, int E, int Cc,
                                                      const float* __restrict__ w0, const float* __restrict__ b0,
                                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                                      const float* __restrict__ w4, const float* __restrict__ b4,
                                                      float* __restrict__ out, int B) {
  extern __shared__ float sh[];  // e[kClnBT][E], h0[kClnBT][E], h1[kClnBT][Cc]
  float* se = sh;
  float* h0 = se + kClnBT * E;
  float* h1 = h0 + kClnBT * E;
  const int n = blockIdx.x, b_lo = blockIdx.y * kClnBT;
  const int nb = min(kClnBT, B - b_lo);
  for (int i = threadIdx.x; i < kClnBT * E; i += blockDim.x) {
    const int j = i / E, c = i - j * E;
    se[i] = j < nb ? e[(long long)(b_lo + j) * E + c] : 0.f;
  }
  __syncthreads();
  cln_layer(w0 + (long long)n * E * E, b0 + (long long)n * E, E, E, se, E, h0, E, nullptr, 0, kClnBT, true);
  __syncthreads();
  cln_layer(w2 + (long long)n * Cc * E, b2 + (long long)n * Cc, Cc, E, h0, E, h1, Cc, nullptr, 0, kClnBT, true);
  __syncthreads();
  cln_layer(w4 + (long long)n * Cc * Cc, b4 + (long long)n * Cc, Cc, Cc, h1, Cc, nullptr, 0,
            out + ((long long)n * B + b_lo) * Cc, Cc, nb, false);
}

template <int DK>
static int launch_attention(const float* qkv, int64_t qkv_bs, int32_t qkv_ld, const float* pos, int32_t pos_ld,
                            int32_t pos_center, int32_t pos_cols, const float* bias_u, const float* bias_v,
                            const int32_t* len, int32_t B, int32_t H, int32_t L_max, float* out, int64_t out_bs,
                            int32_t out_ld, cudaStream_t s) {
  auto kern = relpos_attention_kernel<DK>;
  const int smem = (int)sizeof(AttnSmem<DK>);
  static bool configured_per_dev[kMaxDeviceSlots] = {};   // per kernel instantiation (DK) and device
  const int slot = current_device_slot();
  if (slot < 0) return fail(TB200_E_NODEVICE, "relpos_attention: no current CUDA device");
  if (!configured_per_dev[slot]) {
    TB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_per_dev[slot] = true;
  }
  dim3 grid((L_max + kAtQ - 1) / kAtQ, H, B);
  kern<<<grid, 256, smem, s>>>(qkv, qkv_bs, qkv_ld, pos, pos_ld, pos_center, pos_cols, bias_u, bias_v, len, H, L_max, out,
                              out_bs, out_ld);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace tb200

using namespace tb200;

extern "C" {

int tb200_channel_norm(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                       const int32_t* len, int32_t B, int32_t C, int32_t L_max, const float* gamma, const float* beta,
                       int64_t gb_bs, int32_t mode, float eps, void* stream) {
  if (!x || !y || !gamma || !beta) return fail(TB200_E_BADARG, "channel_norm: null pointer");
  if (B <= 0 || C <= 0 || L_max <= 0 || C > kCnWarps * kCnMaxPerWarp || mode < 0 || mode > 1)
    return fail(TB200_E_BADARG, "channel_norm: bad shape (C must be <= %d)", kCnWarps * kCnMaxPerWarp);
  // (a two-steps-per-lane variant with 8 channels per warp and 768-thread blocks measured 39 % of the HBM peak against
  // 54 % for this one: more, smaller blocks do not overlap better, the cross-warp reductions grow)
  dim3 grid((L_max + 31) / 32, B), block(32, kCnWarps);
  channel_norm_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(x, x_bs, x_ld, y, y_bs, y_ld, len, C, L_max,
                                                                              gamma, beta, gb_bs, mode, eps);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_group_norm(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                     const float* residual, int64_t r_bs, int32_t r_ld, const int32_t* len, int32_t B, int32_t C,
                     int32_t L_max, int32_t groups, const float* gamma, const float* beta, float eps, int32_t out_act,
                     void* stream) {
  if (!x || !y || !gamma || !beta) return fail(TB200_E_BADARG, "group_norm: null pointer");
  if (B <= 0 || C <= 0 || L_max <= 0 || groups <= 0 || C % groups) return fail(TB200_E_BADARG, "group_norm: bad shape");
  dim3 grid(groups, B);
  group_norm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_bs, x_ld, y, y_bs, y_ld, residual, r_bs, r_ld,
                                                                          len, C, L_max, groups, gamma, beta, eps, out_act);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_glu_dwconv(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld, const int32_t* len,
                     int32_t B, int32_t C, int32_t L_max, const float* w, const float* bias, int32_t K,
                     const float* bn_mean, const float* bn_var, const float* bn_gamma, const float* bn_beta, float bn_eps,
                     void* stream) {
  if (!x || !y || !w || !bias || !bn_mean || !bn_var || !bn_gamma || !bn_beta) return fail(TB200_E_BADARG, "glu_dwconv: null pointer");
  if (B <= 0 || C <= 0 || L_max <= 0 || K < 1 || K > kDwMaxK || !(K & 1)) return fail(TB200_E_BADARG, "glu_dwconv: K must be odd and <= %d", kDwMaxK);
  if (C > 65535 || B > 65535) return fail(TB200_E_BADARG, "glu_dwconv: grid too large");
  dim3 grid((L_max + kDwTile - 1) / kDwTile, C, B);
  glu_dwconv_kernel<<<grid, kDwThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, x_bs, x_ld, y, y_bs, y_ld, len, C, L_max, w, bias,
                                                                              K, bn_mean, bn_var, bn_gamma, bn_beta, bn_eps);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_relpos_attention(const float* qkv, int64_t qkv_bs, int32_t qkv_ld, const float* pos, int32_t pos_ld,
                           int32_t pos_center, int32_t pos_cols, const float* bias_u, const float* bias_v,
                           const int32_t* len, int32_t B, int32_t H, int32_t dk, int32_t L_max, float* out,
                           int64_t out_bs, int32_t out_ld, void* stream) {
  if (!qkv || !pos || !bias_u || !bias_v || !out) return fail(TB200_E_BADARG, "relpos_attention: null pointer");
  if (B <= 0 || H <= 0 || L_max <= 0 || B > 65535 || H > 65535) return fail(TB200_E_BADARG, "relpos_attention: bad shape");
  if (pos_center - (L_max - 1) < 0 || pos_center + (L_max - 1) >= pos_cols)
    return fail(TB200_E_BADARG, "relpos_attention: positional table (%d columns, centre %d) too short for L=%d", pos_cols,
                pos_center, L_max);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dk) {
    case 32: return launch_attention<32>(qkv, qkv_bs, qkv_ld, pos, pos_ld, pos_center, pos_cols, bias_u, bias_v, len, B, H, L_max, out, out_bs, out_ld, s);
    case 48: return launch_attention<48>(qkv, qkv_bs, qkv_ld, pos, pos_ld, pos_center, pos_cols, bias_u, bias_v, len, B, H, L_max, out, out_bs, out_ld, s);
    case 64: return launch_attention<64>(qkv, qkv_bs, qkv_ld, pos, pos_ld, pos_center, pos_cols, bias_u, bias_v, len, B, H, L_max, out, out_bs, out_ld, s);
    default: return fail(TB200_E_BADARG, "relpos_attention: head size %d not in {32,48,64}", dk);
  }
}

static int ew_grid(int L_max, int C, int B, dim3& grid) {
  if (C > 65535 || B > 65535) return fail(TB200_E_BADARG, "elementwise: grid too large");
  grid = dim3((L_max + 255) / 256, C, B);
  return 0;
}

int tb200_rowvec_affine(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                        const int32_t* len, int32_t B, int32_t C, int32_t L_max, const float* vec, int64_t vec_bs,
                        float scale, void* stream) {
  if (!y || (!x && !vec)) return fail(TB200_E_BADARG, "rowvec_affine: null pointer");
  if (B <= 0 || C <= 0 || L_max <= 0) return fail(TB200_E_BADARG, "rowvec_affine: bad shape");
  dim3 grid;
  if (int rc = ew_grid(L_max, C, B, grid)) return rc;
  rowvec_affine_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_bs, x_ld, y, y_bs, y_ld, len, C, L_max, vec,
                                                                             vec_bs, scale);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_transpose(const float* in, int64_t in_bs, int32_t in_ld, float* out, int64_t out_bs, int32_t out_ld,
                    const int32_t* len, int32_t B, int32_t C, int32_t L_max, int32_t to_ncl, void* stream) {
  if (!in || !out) return fail(TB200_E_BADARG, "transpose: null pointer");
  if (B <= 0 || C <= 0 || L_max <= 0 || B > 65535) return fail(TB200_E_BADARG, "transpose: bad shape");
  dim3 grid((L_max + 31) / 32, (C + 31) / 32, B), block(32, 8);
  transpose_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(in, in_bs, in_ld, out, out_bs, out_ld, len, C, L_max,
                                                                           to_ncl);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_squeeze2(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld, const int32_t* len,
                   int32_t B, int32_t C, int32_t L_max, int32_t inverse, void* stream) {
  if (!x || !y) return fail(TB200_E_BADARG, "squeeze2: null pointer");
  if (B <= 0 || C <= 0 || L_max <= 0) return fail(TB200_E_BADARG, "squeeze2: bad shape");
  if (C > 65535 || B > 65535) return fail(TB200_E_BADARG, "elementwise: grid too large");
  const int n = inverse ? 2 * L_max : L_max;     // unsqueezed positions, 4 per thread
  dim3 grid((n + 1023) / 1024, C, B);
  squeeze2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_bs, x_ld, y, y_bs, y_ld, len, C, L_max, inverse);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_wn_gate(const float* a, int64_t a_bs, int32_t a_ld, float* y, int64_t y_bs, int32_t y_ld, const int32_t* len,
                  int32_t B, int32_t hidden, int32_t L_max, void* stream) {
  if (!a || !y) return fail(TB200_E_BADARG, "wn_gate: null pointer");
  if (B <= 0 || hidden <= 0 || L_max <= 0) return fail(TB200_E_BADARG, "wn_gate: bad shape");
  if (hidden > 65535 || B > 65535) return fail(TB200_E_BADARG, "elementwise: grid too large");
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && a_ld % 4 == 0 && y_ld % 4 == 0 &&
                   a_bs % 4 == 0 && y_bs % 4 == 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec) {
    dim3 grid((L_max + 1023) / 1024, hidden, B);
    wn_gate_kernel<4><<<grid, 256, 0, s>>>(a, a_bs, a_ld, y, y_bs, y_ld, len, hidden, L_max);
  } else {
    dim3 grid((L_max + 255) / 256, hidden, B);
    wn_gate_kernel<1><<<grid, 256, 0, s>>>(a, a_bs, a_ld, y, y_bs, y_ld, len, hidden, L_max);
  }
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_flow_close(float* x, int64_t x_bs, int32_t x_ld, const float* ml, int64_t ml_bs, int32_t ml_ld,
                     const int32_t* len, int32_t B, int32_t C, int32_t L_max, const float* w_inv, const float* an_bias,
                     const float* an_logs, void* stream) {
  if (!x || !ml || !w_inv || !an_bias || !an_logs) return fail(TB200_E_BADARG, "flow_close: null pointer");
  if (B <= 0 || C <= 0 || (C & 3) || L_max <= 0) return fail(TB200_E_BADARG, "flow_close: C must be a multiple of 4");
  dim3 grid;
  if (int rc = ew_grid(L_max, C / 4, B, grid)) return rc;
  flow_close_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_bs, x_ld, ml, ml_bs, ml_ld, len, C, L_max, w_inv,
                                                                          an_bias, an_logs);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_l2_normalize(const float* x, float* y, int32_t B, int32_t C, void* stream) {
  if (!x || !y || B <= 0 || C <= 0) return fail(TB200_E_BADARG, "l2_normalize: bad argument");
  l2_normalize_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(x, y, B, C);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_cln_mlp(const float* e, int32_t B, int32_t E, int32_t Cc, int32_t N, const float* w0, const float* b0,
                  const float* w2, const float* b2, const float* w4, const float* b4, float* out, void* stream) {
  if (!e || !w0 || !b0 || !w2 || !b2 || !w4 || !b4 || !out) return fail(TB200_E_BADARG, "cln_mlp: null pointer");
  if (B <= 0 || E <= 0 || Cc <= 0 || N <= 0 || B > 65535) return fail(TB200_E_BADARG, "cln_mlp: bad shape");
  dim3 grid(N, (B + kClnBT - 1) / kClnBT);
  const size_t smem = (size_t)kClnBT * (2 * E + Cc) * sizeof(float);
  if (smem > 48 * 1024) return fail(TB200_E_BADARG, "cln_mlp: E=%d, Cc=%d need %zu bytes of shared memory (max 48 KB)", E, Cc, smem);
  cln_mlp_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(e, E, Cc, w0, b0, w2, b2, w4, b4, out, B);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
