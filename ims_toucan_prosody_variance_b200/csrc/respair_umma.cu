// respair_umma.cu -- one residual pair of the vocoder generators as ONE kernel:
//
//     y = out_alpha * (conv2(ACT2(conv1(ACT1(x)) + b1)) + b2) + res_beta * x   (+ y_old)
//
//   conv1: kernel K, dilation d; conv2: kernel K, dilation 1 (both "same" padded at the utterance's own ends).
//   Reference: `xt = a1(x); xt = c1(xt); xt = a2(xt); xt = c2(xt); x = xt + x`
//     BigVGAN  AMPBlock1.forward   TrainingInterfaces/Spectrogram_to_Wave/BigVGAN/AMP.py:51-60
//     HiFiGAN  HiFiGANResidualBlock.forward   Layers/ResidualBlock.py:93-97
//   ACT = LeakyReLU or BigVGAN's anti-aliased SnakeBeta (alias_free_torch.Activation1d).
//
// The tensor between the two convolutions never leaves the SM.  Per time tile (GEMM view as in conv1d_umma.cu:
// M = time, 128 rows per accumulator; N = channels; K = channels per tap; taps = row-shifted descriptors of one
// staged operand tile):
//
//   X   raw input rows of the tile (+ halos), brought in by the TMA engine (cp.async.bulk, one row per channel)
//   A1  = ACT1(X) in the K-major operand layout          (producer warps, from shared memory -- no global loads)
//   acc1 = conv1(A1)  in TMEM                            (tcgen05.mma, one elected thread)
//   leaky: A2 = LeakyReLU(acc1 + b1) written straight from the accumulator rows (thread = time step, 16 channels =
//          two 16-byte operand groups)
//   snake: SCR = fp16(acc1 + b1) as [channel][time] (epilogue warps), then A2 = SNAKE2(SCR) by the producer warps
//          with the same streaming filter that builds A1 (lane = channel, sequential in time)
//   acc2 = conv2(A2) in TMEM;  y = out_alpha (acc2 + b2) + res_beta x (+ y_old)   (epilogue warps, x re-read from L2)
//
// Tile geometry (h = 6 rows of anti-aliasing halo for snake, 0 for leaky; p2 = (K-1)/2; p1 = p2 d; NR = S*128):
//   conv1 computes u1 rows [t0 - p2 - h, +NR); ACT2 yields rows [t0 - p2, + NR - 2h); conv2's first
//   T_out = NR - 2h - 2 p2 rows are the tile's output [t0, t0 + T_out).  A1 rows = NR + 2 p1.
//
// Pipeline: every buffer is single; tile j's conv1 side overlaps tile j-1's conv2 side:
//   producers: P1(j) P2(j-1) | MMA: M1(j) M2(j-1) | epilogue: E1(j) E2(j-1) | X loader one tile ahead.
// Warp roles: producers [0, NP), epilogue [NP, NP+NE) (any 4 consecutive warps cover the 4 TMEM lane quarters),
// then the MMA issuer, the X loader and the weight loader (bulk copies; resident when both convs fit, else a ring).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "snake_stream.cuh"

namespace tb200 {

struct PairArgs {
  const void* x;
  void* y;
  const int* len;
  const float *bias1, *bias2, *alpha1, *beta1, *alpha2, *beta2;
  const void *w1, *w2;
  long long x_bs, y_bs;
  int x_ld, y_ld;
  int B, C, L_max, K, dil;
  int x_f16, y_f16;
  float slope, out_alpha, res_beta;
  int accumulate;
  // residual of the output epilogue: x itself for a pair; any (B, C, L) tensor or none for a single staged conv
  const void* res;
  long long r_bs;
  int r_ld, r_f16;
  int single;             // 1: ONE convolution (conv1 only) through the same TMA-fed pipeline, operand tiles double buffered
  // tile geometry
  int S, NR, h, p1, p2, T_out;
  int R1, Rp1;            // A1 rows used / plane pitch (rows)
  int R2v, Rp2;           // A2 rows written by ACT2 / plane pitch (rows, >= NR + K - 1)
  int PX, PS;             // X / SCR row pitch in elements (pitch bytes = 16 mod 32: conflict-free lane = channel reads)
  int KC, n_kchunks, chunk_bytes, n_chunks;   // per conv: n_chunks = K * n_kchunks blocks of KC x C halves
  int resident, ring_slots;
  int tiles_per_utt, total_tiles;
  int tmem_cols;
  int x_buf_bytes, x_bufs;   // bytes of one X buffer; 1 or 2 buffers (2: single mode only)
  int a1_off, a2_off, x_off, scr_off, w_off, bar_off, bias_off, tmem_off, stage_off, smem_total;
  long long* trace;
  int dbg_skip;           // TB200_PAIR_SKIP (profiling only): 1 no MMAs, 2 no conv2 epilogue body, 4 no conv1 epilogue body
};

// Producer warps of the snake variant (the rest of the 12 worker warps run the epilogues).  The streaming filter is
// bound by the FMA pipe, which one warp per scheduler nearly saturates (tools/snake_rate.cu: 4 warps 0.45-0.49 cycles
// per element per SM, 8 warps 0.39-0.44).  Measured: the fused pair (two snake stages per tile) is 5-15 % faster with
// 8 + 4 at K >= 7; the single-conv mode (one snake stage, epilogue with residual loads) is 8 % faster with 4 + 8; the
// whole BigVGAN step is the same within noise (75.3-75.6 ms).  Default: 8 + 4.
#ifndef TB200_PAIR_NP_SNAKE
#define TB200_PAIR_NP_SNAKE 8
#endif
template <bool SNAKE>
struct PairRoles {
  static constexpr int kProd = TB200_PAIR_NP_SNAKE > 0 && SNAKE ? TB200_PAIR_NP_SNAKE : 4;   // 15 warps: up to 128 registers per thread
  static constexpr int kEpi = 12 - kProd;
  static constexpr int kMma = kProd + kEpi;
  static constexpr int kXLoad = kMma + 1;
  static constexpr int kWLoad = kMma + 2;
  static constexpr int kThreads = (kMma + 3) * 32;
};

enum {  // mbarrier slots (w_full / w_empty rings follow)
  BX_FULL = 0, BX_EMPTY, BA1_FULL, BA1_EMPTY, BACC1_FULL, BACC1_EMPTY, BSCR_FULL, BSCR_EMPTY,
  BA2_FULL, BA2_EMPTY, BACC2_FULL, BACC2_EMPTY, BX2_FULL, BX2_EMPTY, BNUM
};
constexpr int kPairMaxRing = 64;

// Optional timeline of CTA 0 (TB200_TRACE=1): clock64() stamps, 16 per tile.
//  0/1 P1 begin/end  2/3 P2 begin/end  4/5 E1 begin/end  6/7 E2 begin/end  8/9 M1 begin/issued  10/11 M2 begin/issued
//  12 X tile landed (seen by the producers)
constexpr int kPairTraceTiles = 64;
constexpr int kPairTraceLen = kPairTraceTiles * 16 + 160;
__device__ __forceinline__ void ptrace(const PairArgs& a, int slot, uint32_t j) {
  if (a.trace && blockIdx.x == 0 && j < (uint32_t)kPairTraceTiles) a.trace[j * 16 + slot] = clock64();
}

__device__ __forceinline__ void pair_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct PairTile {
  int b, t0, len, tile;
};
__device__ __forceinline__ bool pair_tile(const PairArgs& a, int tile, PairTile& ti) {
  ti.tile = tile;
  ti.b = tile / a.tiles_per_utt;
  ti.t0 = (tile - ti.b * a.tiles_per_utt) * a.T_out;
  ti.len = a.len ? min(__ldg(a.len + ti.b), a.L_max) : a.L_max;
  return ti.t0 < ti.len;
}

// Every role walks the CTA's tiles in the same order: stage-1 work of tile j, then stage-2 work of tile j-1.
template <typename F1, typename F2>
__device__ __forceinline__ void pair_schedule(const PairArgs& a, F1&& stage1, F2&& stage2) {
  // (one call site per stage: the stage bodies are inlined exactly once)
  PairTile prev{0, 0, 0, 0};
  bool have_prev = false;
  uint32_t j = 0;
  int tile = blockIdx.x;
#pragma unroll 1
  for (;;) {
    PairTile cur{0, 0, 0, 0};
    bool have_cur = false;
#pragma unroll 1
    while (tile < a.total_tiles && !have_cur) {
      have_cur = pair_tile(a, tile, cur);
      tile += gridDim.x;
    }
    if (have_cur) stage1(cur, j);
    if (have_prev) stage2(prev, j - 1);
    if (!have_cur) break;
    prev = cur;
    have_prev = true;
    ++j;
  }
}

// ---------------------------------------------------------------------------------------------
// producers
// ---------------------------------------------------------------------------------------------
// One copy of the streaming filter per (EDGE, XF16), shared by both activations of the pair: the steady loop alone is
// ~8 KB of SASS, and with one inlined copy per call site (the first build: 17 K instructions per kernel) the warps
// running ACT1 and ACT2 side by side thrashed the instruction caches (ncu: `no_instruction` the second largest stall).
template <bool EDGE, bool XF16>
__device__ __noinline__ void pair_aa_task(const void* src, long long row, float ea, float ib, int t_lo, int t_beg, int t_end,
                                          int len, __half* dst) {
  aa_channel_task<__half, EDGE, XF16, true>(src, row, ea, ib, t_lo, t_beg, t_end, len, dst);
}

// Anti-aliased SnakeBeta of a [channel][time] shared-memory tile into the K-major operand tile dst ([C/8][Rp][8]):
// warp pw of np takes one (32-channel block, row segment); src element of channel c at time t = src[c*pitch + t - col0_t].
// ea_ib: per channel (e^alpha, 1 / (e^beta + 1e-9)), computed once per CTA into shared memory.
template <bool XF16>
__device__ __forceinline__ void pair_snake_stage(const void* src, int pitch, int col0_t, const float2* ea_ib, int C, int t_lo,
                                                 int R, int Rp, int len, __half* dst_tile, int pw, int np, int lane) {
  const int ncb = C >> 5;
  const int nseg = np / ncb;
  // Segments in whole 8-step blocks of the filter, the same number (+-1) for every warp: interior boundaries sit where
  // (t + 6) % 8 == 0, so only the tile's first and last segment run partial (checked) blocks.
  const int ph0 = (t_lo + 6) & 7;                  // rows [0, R) span blocks 0 .. nblk-1 of 8 rows starting at row -ph0
  const int nblk = (R + ph0 + 7) >> 3;
  auto seg_start = [&](int sgm) {
    if (sgm <= 0) return 0;
    if (sgm >= nseg) return R;
    return min(max(((sgm * nblk + nseg / 2) / nseg) * 8 - ph0, 0), R);
  };
  const int cb = pw / nseg, seg = pw - cb * nseg;
  if (cb >= ncb) return;
  const int r_beg = seg_start(seg), r_end = seg_start(seg + 1);
  if (r_beg >= r_end) return;
  const int cl = cb * 32 + lane;
  const int t_beg = t_lo + r_beg, t_end = t_lo + r_end;
  __half* dst = dst_tile + ((long long)(cl >> 3) * Rp) * 8 + (cl & 7);
  if (t_beg >= len || t_end <= 0) {   // entirely outside the utterance: the conv's zero padding
    for (int t = t_beg; t < t_end; ++t) dst[(long long)(t - t_lo) * 8] = __float2half(0.f);
    return;
  }
  const long long row = (long long)cl * pitch - col0_t;
  const bool edge = (((t_beg - 9) & ~7) < 0) || (t_end + 32 > len);   // warp-uniform
  const float2 pr = ea_ib[cl];
  const float ea = pr.x, ib = pr.y;
  if (edge) pair_aa_task<true, XF16>(src, row, ea, ib, t_lo, t_beg, t_end, len, dst);
  else pair_aa_task<false, XF16>(src, row, ea, ib, t_lo, t_beg, t_end, len, dst);
}

// LeakyReLU of the X tile into A1: a thread takes 8 channels (one operand group) x 2 consecutive time steps.
template <bool XF16>
__device__ __forceinline__ void pair_leaky_stage(const void* X, int PX, int tx0, int C, int ta0, int R1, int Rp1, int len,
                                                 float slope, __half* A1, int pw, int np, int lane) {
  const int t_e0 = ta0 & ~1;                       // first (even) time step of pair 0
  const int npairs = (ta0 + R1 - t_e0 + 1) >> 1;
  const int nchunks = (npairs + 31) >> 5;
  const int ngroups = C >> 3;
  const __half2 slope2 = __float2half2_rn(slope);
#pragma unroll 1
  for (int task = pw; task < ngroups * nchunks; task += np) {
    const int g = task / nchunks, ch = task - g * nchunks;
    const int pr = ch * 32 + lane;
    if (pr >= npairs) continue;
    const int t = t_e0 + 2 * pr;
    const int r = t - ta0;                         // row of the first step (may be -1)
    const bool v0 = t >= 0 && t < len, v1 = t + 1 >= 0 && t + 1 < len;
    uint32_t w[8];                                 // per channel: (step t, step t+1) as half2
    if constexpr (XF16) {
      const __half* xp = reinterpret_cast<const __half*>(X) + (long long)(g * 8) * PX + (t - tx0);
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = *reinterpret_cast<const uint32_t*>(xp + (long long)e * PX);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        __half2 hv = *reinterpret_cast<__half2*>(&w[e]);
        hv = __hmax2(hv, __hmul2(hv, slope2));     // 0 <= slope <= 1 (host-checked)
        w[e] = *reinterpret_cast<uint32_t*>(&hv);
        if (!v0) w[e] &= 0xffff0000u;
        if (!v1) w[e] &= 0x0000ffffu;
      }
    } else {
      const float* xp = reinterpret_cast<const float*>(X) + (long long)(g * 8) * PX + (t - tx0);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 f = *reinterpret_cast<const float2*>(xp + (long long)e * PX);
        const float a0 = v0 ? fmaxf(f.x, f.x * slope) : 0.f, a1 = v1 ? fmaxf(f.y, f.y * slope) : 0.f;
        w[e] = f16x2_sat(a0, a1);
      }
    }
    uint4 lo, hi;                                  // rows t and t+1: 8 channels each
    lo.x = __byte_perm(w[0], w[1], 0x5410); lo.y = __byte_perm(w[2], w[3], 0x5410);
    lo.z = __byte_perm(w[4], w[5], 0x5410); lo.w = __byte_perm(w[6], w[7], 0x5410);
    hi.x = __byte_perm(w[0], w[1], 0x7632); hi.y = __byte_perm(w[2], w[3], 0x7632);
    hi.z = __byte_perm(w[4], w[5], 0x7632); hi.w = __byte_perm(w[6], w[7], 0x7632);
    __half* dst = A1 + ((long long)g * Rp1 + r) * 8;
    if (r >= 0 && r < R1) *reinterpret_cast<uint4*>(dst) = lo;
    if (r + 1 >= 0 && r + 1 < R1) *reinterpret_cast<uint4*>(dst + 8) = hi;
  }
}

// ---------------------------------------------------------------------------------------------
// waits of the roles that are idle most of the time (loaders, MMA issuer, epilogue between phases): poll with a
// back-off, so that their try_wait loops do not take issue slots from the warps that compute (the plain polling
// loops were 17 % of all executed instructions of a snake pair)
// ---------------------------------------------------------------------------------------------
// (__noinline__, loops not unrolled: one small copy each -- inlined at every wait site they were 3.5 K instructions)
#ifndef TB200_PAIR_SLEEP_NS
#define TB200_PAIR_SLEEP_NS 128
#endif
__device__ __noinline__ void pair_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    __nanosleep(TB200_PAIR_SLEEP_NS);
    if (mbar_try_wait(bar, parity)) return;
  }
  mbar_timeout();
}
__device__ __noinline__ void pair_wait(uint64_t* bar, uint32_t parity) {   // producers: usually already complete
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    if (mbar_try_wait(bar, parity)) return;
  }
  mbar_timeout();
}

// ---------------------------------------------------------------------------------------------
// epilogue of conv2:  y = out_alpha * (acc2 + b2) + res_beta * x (+ y_old)
// An accumulator row is one time step (thread) x 16 channels; global memory is [channel][time].  Each item
// (128-row sub-tile, 16-channel slab; a warp owns 32 rows) is transposed through a per-warp shared-memory stage
// ([16 channels][32 rows] fp32, row pitch 36 floats: conflict-free both ways) so that every thread then owns 8
// consecutive time steps of 2 channels: 16-byte loads of the residual / old y and 16-byte stores instead of 2-byte ones
// (16x fewer memory instructions), and the residual vectors of the next kD items are in flight while an item is finished
// (the scalar version paid one L2 latency per pair of items: 15-20 K cycles per tile).
// Tile origins are multiples of 8 rows (T_out % 8 == 0), so the 8-row groups are 16-byte aligned.
// ---------------------------------------------------------------------------------------------
constexpr int kStagePitch = 36;
constexpr int kStageFloats = 16 * kStagePitch;

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4 (&v)[1], float (&f)[8]) {   // 8 halves
  const __half2* h = reinterpret_cast<const __half2*>(&v[0]);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack8(const uint4 (&v)[2], float (&f)[8]) {   // 8 floats
  f[0] = __uint_as_float(v[0].x); f[1] = __uint_as_float(v[0].y); f[2] = __uint_as_float(v[0].z); f[3] = __uint_as_float(v[0].w);
  f[4] = __uint_as_float(v[1].x); f[5] = __uint_as_float(v[1].y); f[6] = __uint_as_float(v[1].z); f[7] = __uint_as_float(v[1].w);
}

// the 8-row group that straddles the end of the utterance (or of the tile): element by element (rare)
template <bool XF16, bool YF16>
__device__ __noinline__ void pair_e2_tail(const char* xsrc, char* ydst, float4 fa, float4 fb, int n, float res_beta, bool accumulate) {
  const float f[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll 1
  for (int e = 0; e < n; ++e) {
    const float xv = !xsrc ? 0.f : XF16 ? __half2float(reinterpret_cast<const __half*>(xsrc)[e]) : reinterpret_cast<const float*>(xsrc)[e];
    float val = fmaf(res_beta, xv, f[e]);
    if (accumulate) val += YF16 ? __half2float(reinterpret_cast<const __half*>(ydst)[e]) : reinterpret_cast<const float*>(ydst)[e];
    if constexpr (YF16) reinterpret_cast<__half*>(ydst)[e] = f16_sat(val);
    else reinterpret_cast<float*>(ydst)[e] = val;
  }
}

template <bool XF16, bool YF16, bool ACC, bool RES = true>
__device__ __forceinline__ void pair_e2(const PairArgs& a, const PairTile& ti, int nsub, float* stg, const float* bias2_s,
                                        uint32_t tm, int q, int lane, int slab0, int slab_step, int slabs) {
  // items whose residual (and, for the multi-receptive-field sum, old y) vectors are in flight
  constexpr int kD = XF16 ? (ACC ? 2 : 3) : 1;
  constexpr int XV = XF16 ? 1 : 2, YV = YF16 ? 1 : 2;   // 16-byte vectors per 8 elements
  constexpr int XE = XF16 ? 2 : 4, YE = YF16 ? 2 : 4;   // element sizes
  const int nsl = (slabs - slab0 + slab_step - 1) / slab_step;
  const int nitems = nsub * nsl;
  const int limit = min(a.T_out, ti.len - ti.t0);   // valid output rows of this tile
  const int c0 = lane >> 2, g = lane & 3;           // this thread's slots of an item: channels c0, c0 + 8; rows 8g .. 8g+7 of the warp's 32
  const int r8 = q * 32 + 8 * g;
  const char* xrow = reinterpret_cast<const char*>(a.res) + ((long long)ti.b * a.r_bs + ti.t0 + r8) * XE;
  char* yrow = reinterpret_cast<char*>(a.y) + ((long long)ti.b * a.y_bs + ti.t0 + r8) * YE;
  const float out_alpha = a.out_alpha, res_beta = a.res_beta;
  constexpr bool accumulate = ACC;
  constexpr int PV = (RES ? XV : 0) + (ACC ? YV : 0) + ((RES || ACC) ? 0 : 1);   // residual vectors, then the old y vectors
  constexpr int YO = RES ? XV : 0;                  // first old-y vector
  uint4 pre[kD][2][PV];
  auto geom = [&](int k, int& sub, int& s) {
    sub = k / nsl;
    s = slab0 + (k - sub * nsl) * slab_step;
  };
  auto prefetch = [&](int k, uint4 (&p)[2][PV]) {
    int sub, s;
    geom(k, sub, s);
    const bool full = sub * kTileM + r8 + 8 <= limit;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if constexpr (RES) {
        const char* src = xrow + ((long long)(s * 16 + c0 + 8 * i) * a.r_ld + sub * kTileM) * XE;
#pragma unroll
        for (int v = 0; v < XV; ++v) p[i][v] = full ? ldg128(src + 16 * v) : make_uint4(0u, 0u, 0u, 0u);
      }
      if constexpr (ACC) {
        const char* ysrc = yrow + ((long long)(s * 16 + c0 + 8 * i) * a.y_ld + sub * kTileM) * YE;
#pragma unroll
        for (int v = 0; v < YV; ++v)
          p[i][YO + v] = full ? *reinterpret_cast<const uint4*>(ysrc + 16 * v) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  auto process = [&](int k, const uint4 (&p)[2][PV]) {
    int sub, s;
    geom(k, sub, s);
    uint32_t v[16];
    __syncwarp();                                    // the previous item's stage reads are done
    tmem_ld_x16(tm + (uint32_t)(sub * a.C + s * 16), v);
    float bias[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias2_s + s * 16 + 4 * i);
      bias[4 * i] = b4.x; bias[4 * i + 1] = b4.y; bias[4 * i + 2] = b4.z; bias[4 * i + 3] = b4.w;
    }
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) stg[i * kStagePitch + lane] = fmaf(__uint_as_float(v[i]), out_alpha, bias[i]);
    __syncwarp();
    const int o8 = sub * kTileM + r8;
    if (o8 >= limit) return;
    const bool full = o8 + 8 <= limit;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int cl = c0 + 8 * i;
      const long long crow = (long long)(s * 16 + cl);
      char* ydst = yrow + (crow * a.y_ld + sub * kTileM) * YE;
      float f[8];
      {
        const float4 s0 = *reinterpret_cast<const float4*>(stg + cl * kStagePitch + 8 * g);
        const float4 s1 = *reinterpret_cast<const float4*>(stg + cl * kStagePitch + 8 * g + 4);
        f[0] = s0.x; f[1] = s0.y; f[2] = s0.z; f[3] = s0.w; f[4] = s1.x; f[5] = s1.y; f[6] = s1.z; f[7] = s1.w;
      }
      if (full) {
        if constexpr (RES) {
          float r[8];
          uint4 xv[XV];
#pragma unroll
          for (int w = 0; w < XV; ++w) xv[w] = p[i][w];
          unpack8(xv, r);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaf(res_beta, r[e], f[e]);
        }
        if constexpr (ACC) {   // the multi-receptive-field sum: last pair of the 2nd / 3rd residual block only
          uint4 yv[YV];
#pragma unroll
          for (int w = 0; w < YV; ++w) yv[w] = p[i][YO + w];
          float yo[8];
          unpack8(yv, yo);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += yo[e];
        }
        if constexpr (YF16) {
          *reinterpret_cast<uint4*>(ydst) = make_uint4(f16x2_sat(f[0], f[1]), f16x2_sat(f[2], f[3]), f16x2_sat(f[4], f[5]), f16x2_sat(f[6], f[7]));
        } else {
          *reinterpret_cast<float4*>(ydst) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(ydst + 16) = make_float4(f[4], f[5], f[6], f[7]);
        }
      } else {
        pair_e2_tail<XF16, YF16>(RES ? xrow + (crow * a.r_ld + sub * kTileM) * XE : nullptr, ydst, make_float4(f[0], f[1], f[2], f[3]),
                                 make_float4(f[4], f[5], f[6], f[7]), limit - o8, res_beta, accumulate);
      }
    }
  };
#pragma unroll
  for (int d = 0; d < kD; ++d)
    if (d < nitems) prefetch(d, pre[d]);
  for (int k0 = 0; k0 < nitems; k0 += kD) {
#pragma unroll
    for (int d = 0; d < kD; ++d) {
      const int k = k0 + d;
      if (k < nitems) {
        process(k, pre[d]);
        if (k + kD < nitems) prefetch(k + kD, pre[d]);
      }
    }
  }
}

// Any other combination of stream types (fp32 residual streams: the high-precision mode and the tests): one compact
// copy, thread = time step, element-wise global accesses (a warp reads / writes 32 consecutive time steps per channel).
// Kept out of line so that its registers and code do not weigh on the fp16 variants.
__device__ __noinline__ void pair_e2_generic(const PairArgs& a, const PairTile& ti, int nsub, const float* bias2_s, uint32_t tm,
                                             int q, int lane, int slab0, int slab_step, int slabs) {
  const int limit = min(a.T_out, ti.len - ti.t0);
  const int XE = a.r_f16 ? 2 : 4, YE = a.y_f16 ? 2 : 4;
#pragma unroll 1
  for (int sub = 0; sub < nsub; ++sub) {
    const int o = sub * kTileM + q * 32 + lane;
    const bool ok = o < limit;
    const long long t = ti.t0 + (ok ? o : 0);
#pragma unroll 1
    for (int s = slab0; s < slabs; s += slab_step) {
      uint32_t v[16];
      __syncwarp();
      tmem_ld_x16(tm + (uint32_t)(sub * a.C + s * 16), v);
      tmem_ld_wait();
      if (!ok) continue;
      const char* xp = a.res ? reinterpret_cast<const char*>(a.res) + ((long long)ti.b * a.r_bs + (long long)(s * 16) * a.r_ld + t) * XE : nullptr;
      char* yp = reinterpret_cast<char*>(a.y) + ((long long)ti.b * a.y_bs + (long long)(s * 16) * a.y_ld + t) * YE;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const float xv = !xp ? 0.f : a.r_f16 ? __half2float(*reinterpret_cast<const __half*>(xp)) : *reinterpret_cast<const float*>(xp);
        float val = fmaf(a.res_beta, xv, fmaf(__uint_as_float(v[i]), a.out_alpha, bias2_s[s * 16 + i]));
        if (a.accumulate) val += a.y_f16 ? __half2float(*reinterpret_cast<const __half*>(yp)) : *reinterpret_cast<const float*>(yp);
        if (a.y_f16) *reinterpret_cast<__half*>(yp) = f16_sat(val);
        else *reinterpret_cast<float*>(yp) = val;
        if (xp) xp += (long long)a.r_ld * XE;
        yp += (long long)a.y_ld * YE;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool SNAKE>
__global__ void __launch_bounds__(PairRoles<SNAKE>::kThreads, 1) respair_kernel(const __grid_constant__ PairArgs a) {
  using Rl = PairRoles<SNAKE>;
  constexpr int NP = Rl::kProd, NE = Rl::kEpi;
  extern __shared__ __align__(128) uint8_t smem[];
  __half* A1 = reinterpret_cast<__half*>(smem + a.a1_off);
  __half* A2 = reinterpret_cast<__half*>(smem + a.a2_off);
  uint8_t* Xbase = smem + a.x_off;
  __half* SCR = reinterpret_cast<__half*>(smem + a.scr_off);
  uint8_t* smW = smem + a.w_off;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.bar_off);
  uint64_t* w_full = bars + BNUM;
  uint64_t* w_empty = w_full + a.ring_slots;
  float* bias1_s = reinterpret_cast<float*>(smem + a.bias_off);
  float* bias2_s = bias1_s + a.C;                 // bias2 * out_alpha
  float2* snake1_s = reinterpret_cast<float2*>(bias2_s + a.C);   // (e^alpha, 1/(e^beta + 1e-9)) of ACT1, ACT2
  float2* snake2_s = snake1_s + a.C;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + a.tmem_off);
  float* stage_s = reinterpret_cast<float*>(smem + a.stage_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_kernel_start = clock64();

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bars + BX_FULL, 1);
      mbar_init(bars + BX_EMPTY, NP);
      mbar_init(bars + BX2_FULL, 1);       // single conv: the X tile is double buffered as well (tile j in buffer j & 1)
      mbar_init(bars + BX2_EMPTY, NP);
      mbar_init(bars + BA1_FULL, NP);
      mbar_init(bars + BA1_EMPTY, 1);
      mbar_init(bars + BACC1_FULL, 1);
      mbar_init(bars + BACC1_EMPTY, NE);
      mbar_init(bars + BSCR_FULL, NE);
      mbar_init(bars + BSCR_EMPTY, NP);
      mbar_init(bars + BA2_FULL, (SNAKE || a.single) ? NP : NE);
      mbar_init(bars + BA2_EMPTY, 1);
      mbar_init(bars + BACC2_FULL, 1);
      mbar_init(bars + BACC2_EMPTY, NE);
      for (int i = 0; i < a.ring_slots; ++i) {
        mbar_init(w_full + i, 1);
        mbar_init(w_empty + i, 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, a.tmem_cols);
    tmem_relinquish();
  }
  // operand tiles start as zeros: the rows of A2 behind the last one ACT2 writes are read by (discarded) output rows only,
  // the padding rows of the planes never; zero keeps them finite
  {
    uint4* z = reinterpret_cast<uint4*>(smem + a.a1_off);
    const int n16 = (a.x_off - a.a1_off) >> 4;    // A1 and A2 are adjacent
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = threadIdx.x; i < a.C; i += blockDim.x) {
    bias1_s[i] = a.bias1 ? __ldg(a.bias1 + i) : 0.f;
    const float* bout = a.single ? a.bias1 : a.bias2;       // bias of the conv whose accumulators the output epilogue drains
    bias2_s[i] = bout ? __ldg(bout + i) * a.out_alpha : 0.f;
    if constexpr (SNAKE) {
      snake1_s[i] = make_float2(__expf(__ldg(a.alpha1 + i)), 1.0f / (__expf(__ldg(a.beta1 + i)) + 1e-9f));
      if (!a.single) snake2_s[i] = make_float2(__expf(__ldg(a.alpha2 + i)), 1.0f / (__expf(__ldg(a.beta2 + i)) + 1e-9f));
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1_col = tmem_base, acc2_col = tmem_base + (uint32_t)(a.S * a.C);

  const int esz = a.x_f16 ? 2 : 4;
  // per-tile time origins
  auto tu0_of = [&](const PairTile& ti) { return ti.t0 - a.p2 - a.h; };      // first u1 (conv1 output) row
  auto ta0_of = [&](const PairTile& ti) { return ti.t0 - a.p2 - a.h - a.p1; };  // first A1 row
  auto tx0_of = [&](const PairTile& ti) { return SNAKE ? ((ta0_of(ti) - 9) & ~7) : (ta0_of(ti) & ~7); };   // X column 0
  // sub-tiles that matter: u1 rows are needed up to len + p2 + h, output rows up to min(T_out, len - t0)
  auto nsub1_of = [&](const PairTile& ti) { return min(a.S, (ti.len + a.p2 + a.h - tu0_of(ti) + kTileM - 1) / kTileM); };
  auto nsub2_of = [&](const PairTile& ti) { return min(a.S, (min(a.T_out, ti.len - ti.t0) + kTileM - 1) / kTileM); };

  if (warp < NP) {
    // ================================ producers ================================
    const int pw = warp;
    auto p1 = [&](const PairTile& ti, uint32_t j) {
      // single conv: the two operand tiles (A1, A2 regions) are the two buffers of one conv, tile j uses buffer j & 1
      const int buf = a.single ? (int)(j & 1) : 0;
      const uint32_t use = a.single ? (j >> 1) : j;
      const int xbuf = a.x_bufs == 2 ? buf : 0;
      const uint32_t xuse = a.x_bufs == 2 ? use : j;
      pair_wait(bars + (xbuf ? BX2_FULL : BX_FULL), xuse & 1);
      pair_wait(bars + (buf ? BA2_EMPTY : BA1_EMPTY), (use & 1) ^ 1);
      if (threadIdx.x == 0) { ptrace(a, 12, j); ptrace(a, 0, j); }
      const int ta0 = ta0_of(ti), tx0 = tx0_of(ti);
      __half* Adst = buf ? A2 : A1;
      const uint8_t* X = Xbase + (size_t)xbuf * a.x_buf_bytes;
      if constexpr (SNAKE) {
        if (a.x_f16) pair_snake_stage<true>(X, a.PX, tx0, snake1_s, a.C, ta0, a.R1, a.Rp1, ti.len, Adst, pw, NP, lane);
        else pair_snake_stage<false>(X, a.PX, tx0, snake1_s, a.C, ta0, a.R1, a.Rp1, ti.len, Adst, pw, NP, lane);
      } else {
        if (a.x_f16) pair_leaky_stage<true>(X, a.PX, tx0, a.C, ta0, a.R1, a.Rp1, ti.len, a.slope, Adst, pw, NP, lane);
        else pair_leaky_stage<false>(X, a.PX, tx0, a.C, ta0, a.R1, a.Rp1, ti.len, a.slope, Adst, pw, NP, lane);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        pair_mbar_arrive(bars + (buf ? BA2_FULL : BA1_FULL));
        pair_mbar_arrive(bars + (xbuf ? BX2_EMPTY : BX_EMPTY));
      }
      if (threadIdx.x == 0) ptrace(a, 1, j);
    };
    auto p2 = [&](const PairTile& ti, uint32_t j) {
      if (a.single) return;
      if constexpr (SNAKE) {
        pair_wait(bars + BSCR_FULL, j & 1);
        pair_wait(bars + BA2_EMPTY, (j & 1) ^ 1);
        if (threadIdx.x == 0) ptrace(a, 2, j);
        const int tu0 = tu0_of(ti);
        const int tsc0 = tu0 & ~7;               // SCR column 8 holds time tsc0
        pair_snake_stage<true>(SCR, a.PS, tsc0 - 8, snake2_s, a.C, tu0 + a.h, a.R2v, a.Rp2, ti.len, A2, pw, NP, lane);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          pair_mbar_arrive(bars + BA2_FULL);
          pair_mbar_arrive(bars + BSCR_EMPTY);
        }
        if (threadIdx.x == 0) ptrace(a, 3, j);
      }
    };
    pair_schedule(a, p1, p2);
  } else if (warp < NP + NE) {
    // ================================ epilogue ================================
    const int ew = warp - NP;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int slab_step = NE >> 2, slab0 = ew >> 2;
    const int slabs = a.C >> 4;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    auto e1 = [&](const PairTile& ti, uint32_t j) {
      if (a.single) return;                         // single conv: only the output epilogue below
      pair_wait_sleep(bars + BACC1_FULL, j & 1);
      if constexpr (SNAKE) pair_wait_sleep(bars + BSCR_EMPTY, (j & 1) ^ 1);
      else pair_wait_sleep(bars + BA2_EMPTY, (j & 1) ^ 1);
      tc_fence_after();
      if (ew == 0 && lane == 0) ptrace(a, 4, j);
      const int tu0 = tu0_of(ti);
      const int nsub = (a.dbg_skip & 4) ? 0 : nsub1_of(ti);
      for (int sub = 0; sub < nsub; ++sub) {
        const int r = sub * kTileM + q * 32 + lane;     // u1 row of this thread
        for (int s = slab0; s < slabs; s += slab_step) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld_x16(acc1_col + lane_base + (uint32_t)(sub * a.C + s * 16), v);
          float bias[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias1_s + s * 16 + 4 * i);
            bias[4 * i] = b4.x; bias[4 * i + 1] = b4.y; bias[4 * i + 2] = b4.z; bias[4 * i + 3] = b4.w;
          }
          tmem_ld_wait();
          const bool row_ok = r < a.NR;                  // the last sub-tile may be partial (NR is a multiple of 32)
          if constexpr (SNAKE) {
            // [channel][time] fp16 scratch for the streaming filter; column 8 <-> time tu0 & ~7
            __half* dst = SCR + (long long)(s * 16) * a.PS + (tu0 - (tu0 & ~7) + 8 + r);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __half hv = f16_sat(__uint_as_float(v[i]) + bias[i]);
              if (row_ok) dst[(long long)i * a.PS] = hv;
            }
          } else {
            const int t = tu0 + r;
            const bool valid = t >= 0 && t < ti.len;     // conv2 sees zeros outside the utterance
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float f0 = __uint_as_float(v[2 * i]) + bias[2 * i], f1 = __uint_as_float(v[2 * i + 1]) + bias[2 * i + 1];
              f0 = fmaxf(f0, f0 * a.slope);
              f1 = fmaxf(f1, f1 * a.slope);
              o[i] = valid ? f16x2_sat(f0, f1) : 0u;
            }
            __half* dst = A2 + ((long long)(2 * s) * a.Rp2 + r) * 8;
            if (row_ok) {
              *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(dst + (long long)a.Rp2 * 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
        }
      }
      tc_fence_before();
      if constexpr (!SNAKE) fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        pair_mbar_arrive(bars + BACC1_EMPTY);
        pair_mbar_arrive(bars + (SNAKE ? BSCR_FULL : BA2_FULL));
      }
      if (ew == 0 && lane == 0) ptrace(a, 5, j);
    };
    // output epilogue: conv2 of a pair, or the conv of the single mode (tile j in accumulator buffer j & 1).  One call
    // site, so that its variants are inlined once.
    auto e2 = [&](const PairTile& ti, uint32_t j) {
      const int buf = a.single ? (int)(j & 1) : 1;
      const uint32_t use = a.single ? (j >> 1) : j;
      pair_wait_sleep(bars + (buf ? BACC2_FULL : BACC1_FULL), use & 1);
      tc_fence_after();
      if (ew == 0 && lane == 0) ptrace(a, 6, j);
      const uint32_t tm = (buf ? acc2_col : acc1_col) + lane_base;
      const int nsub = a.single ? nsub1_of(ti) : nsub2_of(ti);
      float* stg = stage_s + ew * kStageFloats;
      if (a.dbg_skip & 2) {
      } else if (a.y_f16 && (a.r_f16 || !a.res)) {   // the generators' fp16 streams: the hot variants, inlined
        if (!a.res) pair_e2<true, true, false, false>(a, ti, nsub, stg, bias2_s, tm, q, lane, slab0, slab_step, slabs);
        else if (a.accumulate) pair_e2<true, true, true>(a, ti, nsub, stg, bias2_s, tm, q, lane, slab0, slab_step, slabs);
        else pair_e2<true, true, false>(a, ti, nsub, stg, bias2_s, tm, q, lane, slab0, slab_step, slabs);
      } else {
        pair_e2_generic(a, ti, nsub, bias2_s, tm, q, lane, slab0, slab_step, slabs);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) pair_mbar_arrive(bars + (buf ? BACC2_EMPTY : BACC1_EMPTY));
      if (ew == 0 && lane == 0) ptrace(a, 7, j);
    };
    pair_schedule(a, e1, e2);
  } else if (warp == Rl::kMma) {
    // ================================ MMA issuer ================================
    // (whole warp convergent, one elected lane issues: see conv1d_umma.cu)
    const uint32_t idesc = make_instr_desc(a.C, false);
    const uint32_t desc_hi = smem_desc_hi(128);
    const uint32_t lbo_b = (uint32_t)a.C * 16u;
    const uint32_t b_kstep = (2u * lbo_b) >> 4;
    const uint32_t smW_u = smem_u32(smW);
    const int nks = a.KC / 16;
    uint32_t cc = 0;                              // weight chunks consumed (ring position)
    bool first1 = true, first2 = true;
    // one conv over an operand tile: taps x channel chunks x sub-tiles x K steps
    auto conv = [&](uint32_t smA_u, int Rp, int tap_stride, int nsub, uint32_t acc_col, int chunk0, bool& first) {
      const uint32_t lbo_a = (uint32_t)Rp * 16u;
      const uint32_t a_kstep = (2u * lbo_a) >> 4;
#pragma unroll 1
      for (int jt = 0; jt < a.K; ++jt) {
        const uint32_t a_row = smA_u + (uint32_t)(jt * tap_stride) * 16u;
#pragma unroll 1
        for (int kc = 0; kc < a.n_kchunks; ++kc, ++cc) {
          int slot;
          if (a.resident) {
            slot = chunk0 + jt * a.n_kchunks + kc;
            if (first) pair_wait(w_full + slot, 0);
          } else {
            slot = cc % a.ring_slots;
            pair_wait(w_full + slot, (cc / a.ring_slots) & 1);
          }
          __syncwarp();
          tc_fence_after();
          const uint32_t b_base = smW_u + (uint32_t)slot * (uint32_t)a.chunk_bytes;
          const uint32_t a_base = a_row + (uint32_t)(kc * (a.KC / 8)) * lbo_a;
          const uint32_t fresh = (jt == 0 && kc == 0) ? 0u : 1u;
          const uint32_t a_lo0 = smem_desc_lo(a_base, lbo_a), b_lo0 = smem_desc_lo(b_base, lbo_b);
          if (elect_one()) {
#pragma unroll 1
            for (int sub = 0; sub < ((a.dbg_skip & 1) ? 0 : nsub); ++sub) {
              uint32_t a_lo = a_lo0 + (uint32_t)(sub * kTileM), b_lo = b_lo0;   // 16 bytes per row -> +1 per row
              const uint32_t d_col = acc_col + (uint32_t)(sub * a.C);
              umma_ss_lohi<false>(d_col, a_lo, desc_hi, b_lo, desc_hi, idesc, fresh);
              for (int ks = 1; ks < nks; ++ks) {
                a_lo += a_kstep;
                b_lo += b_kstep;
                umma_ss_lohi<false>(d_col, a_lo, desc_hi, b_lo, desc_hi, idesc, 1u);
              }
            }
            if (!a.resident) umma_commit(w_empty + slot);
          }
          __syncwarp();
        }
      }
      first = false;
    };
    auto m1 = [&](const PairTile& ti, uint32_t j) {
      const int buf = a.single ? (int)(j & 1) : 0;
      const uint32_t use = a.single ? (j >> 1) : j;
      pair_wait_sleep(bars + (buf ? BA2_FULL : BA1_FULL), use & 1);
      pair_wait_sleep(bars + (buf ? BACC2_EMPTY : BACC1_EMPTY), (use & 1) ^ 1);
      __syncwarp();
      if (lane == 0) ptrace(a, 8, j);
      tc_fence_after();
      conv(smem_u32(buf ? A2 : A1), a.Rp1, a.dil, nsub1_of(ti), buf ? acc2_col : acc1_col, 0, first1);
      if (elect_one()) {
        umma_commit(bars + (buf ? BA2_EMPTY : BA1_EMPTY));
        umma_commit(bars + (buf ? BACC2_FULL : BACC1_FULL));
      }
      __syncwarp();
      if (lane == 0) ptrace(a, 9, j);
    };
    auto m2 = [&](const PairTile& ti, uint32_t j) {
      if (a.single) return;
      pair_wait_sleep(bars + BA2_FULL, j & 1);
      pair_wait_sleep(bars + BACC2_EMPTY, (j & 1) ^ 1);
      __syncwarp();
      if (lane == 0) ptrace(a, 10, j);
      tc_fence_after();
      conv(smem_u32(A2), a.Rp2, 1, nsub2_of(ti), acc2_col, a.n_chunks, first2);
      if (elect_one()) {
        umma_commit(bars + BA2_EMPTY);
        umma_commit(bars + BACC2_FULL);
      }
      __syncwarp();
      if (lane == 0) ptrace(a, 11, j);
    };
    pair_schedule(a, m1, m2);
  } else if (warp == Rl::kXLoad) {
    // ================================ X loader ================================
    // one bulk copy per channel row: global [b][c][lo, hi) -> X[c][lo - tx0 ...]; lanes share the rows
    // rows [lo, hi) of the X tile of `ti` (16-byte multiples, inside the utterance's padded row)
    auto x_range = [&](const PairTile& ti, int& lo, int& hi) {
      const int ta0 = ta0_of(ti);
      const int tx_end = (ta0 + a.R1 + (SNAKE ? 6 : 0) + 7) & ~7;
      const int unit = 16 / esz;
      const int len_al = min((ti.len + unit - 1) / unit * unit, a.x_ld);
      lo = max(tx0_of(ti), 0);
      hi = min(tx_end, len_al);
    };
    auto xl = [&](const PairTile& ti, uint32_t j) {
      // The X tile is single: the load of tile j can only start when the producers are done with tile j-1 and is then on
      // their critical path.  Pull this CTA's NEXT tile into L2 while the current one is being staged, so that the
      // exposed part is an L2 hit (measured 6.2 K cycles per tile from DRAM).
      {
        PairTile nx;
        const int ntile = ti.tile + (int)gridDim.x;
        if (ntile < a.total_tiles && pair_tile(a, ntile, nx)) {
          int nlo, nhi;
          x_range(nx, nlo, nhi);
          const uint32_t nb = (uint32_t)(nhi - nlo) * (uint32_t)esz;
          const uint8_t* nsrc = reinterpret_cast<const uint8_t*>(a.x) + ((long long)nx.b * a.x_bs + nlo) * esz;
#pragma unroll 1
          for (int c = lane; c < a.C; c += 32)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nsrc + (long long)c * a.x_ld * esz), "r"(nb) : "memory");
        }
      }
      const int buf = a.x_bufs == 2 ? (int)(j & 1) : 0;
      const uint32_t use = a.x_bufs == 2 ? (j >> 1) : j;
      uint64_t* xfull = bars + (buf ? BX2_FULL : BX_FULL);
      pair_wait_sleep(bars + (buf ? BX2_EMPTY : BX_EMPTY), (use & 1) ^ 1);
      __syncwarp();
      const int tx0 = tx0_of(ti);
      int lo, hi;
      x_range(ti, lo, hi);
      const uint32_t nbytes = (uint32_t)(hi - lo) * (uint32_t)esz;
      if (lane == 0) mbar_arrive_expect_tx(xfull, nbytes * (uint32_t)a.C);
      __syncwarp();
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.x) + ((long long)ti.b * a.x_bs + lo) * esz;
      uint8_t* dst = Xbase + (size_t)buf * a.x_buf_bytes + (long long)(lo - tx0) * esz;
#pragma unroll 1
      for (int c = lane; c < a.C; c += 32)
        bulk_copy_g2s(dst + (long long)c * a.PX * esz, src + (long long)c * a.x_ld * esz, nbytes, xfull);
    };
    auto none = [&](const PairTile&, uint32_t) {};
    pair_schedule(a, xl, none);
  } else {
    // ================================ weight loader ================================
    const uint8_t* w1 = reinterpret_cast<const uint8_t*>(a.w1);
    const uint8_t* w2 = reinterpret_cast<const uint8_t*>(a.w2);
    if (a.resident) {
      // only when this CTA has a tile to run (a CTA must not exit with bulk copies in flight)
      bool any = false;
#pragma unroll 1
      for (int tile = blockIdx.x; tile < a.total_tiles && !any; tile += gridDim.x) {
        PairTile ti;
        any = pair_tile(a, tile, ti);
      }
      if (any) {
#pragma unroll 1
        for (int c = 0; c < (a.single ? 1 : 2) * a.n_chunks; ++c) {
          if (elect_one()) {
            const uint8_t* src = c < a.n_chunks ? w1 + (long long)c * a.chunk_bytes : w2 + (long long)(c - a.n_chunks) * a.chunk_bytes;
            mbar_arrive_expect_tx(w_full + c, a.chunk_bytes);
            bulk_copy_g2s(smW + (long long)c * a.chunk_bytes, src, a.chunk_bytes, w_full + c);
          }
          __syncwarp();
        }
      }
    } else {
      uint32_t cc = 0;
      auto stream = [&](const uint8_t* w) {
#pragma unroll 1
        for (int c = 0; c < a.n_chunks; ++c, ++cc) {
          const int slot = cc % a.ring_slots;
          pair_wait_sleep(w_empty + slot, ((cc / a.ring_slots) & 1) ^ 1);
          __syncwarp();
          if (elect_one()) {
            mbar_arrive_expect_tx(w_full + slot, a.chunk_bytes);
            bulk_copy_g2s(smW + (long long)slot * a.chunk_bytes, w + (long long)c * a.chunk_bytes, a.chunk_bytes, w_full + slot);
          }
          __syncwarp();
        }
      };
      auto s1 = [&](const PairTile&, uint32_t) { stream(w1); };
      auto s2 = [&](const PairTile&, uint32_t) {
        if (!a.single) stream(w2);
      };
      pair_schedule(a, s1, s2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (a.trace && threadIdx.x == 0 && blockIdx.x < 160) a.trace[kPairTraceTiles * 16 + blockIdx.x] = clock64() - t_kernel_start;
  if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct PairDevice {
  int sm_count = 0, max_smem = 0;
  bool configured[2] = {false, false};
};
static PairDevice g_pair_dev[64];
static long long* g_pair_trace = nullptr;
static int g_pair_trace_on = -1, g_pair_debug = -1, g_pair_max_s = -1, g_pair_skip = 0, g_pair_min_ring = 0;

static int pair_device(PairDevice*& d) {
  int dev = 0;
  TB200_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(TB200_E_BADARG, "respair: device index %d", dev);
  d = &g_pair_dev[dev];
  if (!d->sm_count) {
    TB200_CUDA_CHECK(cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, dev));
    TB200_CUDA_CHECK(cudaDeviceGetAttribute(&d->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  if (g_pair_trace_on < 0) {   // tuning / debugging knobs are read once per process
    g_pair_trace_on = getenv("TB200_TRACE") != nullptr;
    g_pair_debug = getenv("TB200_PLAN_DEBUG") != nullptr;
    const char* e = getenv("TB200_PAIR_MAX_S");
    g_pair_max_s = e ? atoi(e) : 0;
    e = getenv("TB200_PAIR_SKIP");
    g_pair_skip = e ? atoi(e) : 0;
    e = getenv("TB200_PAIR_MIN_RING");
    g_pair_min_ring = e ? atoi(e) : 0;
  }
  return 0;
}

// Row pitch (elements) of a [channel][time] tile read with lane = channel, 16 bytes per lane: pitch bytes = 16 (mod 32),
// so the eight lanes of a quarter-warp phase fall on eight different 16-byte bank groups.
static int round_pitch(int rows, int esz) {
  return (((rows * esz + 31) / 32) * 32 + 16) / esz;
}
static int plane_rows(int rows) { return rows % 4 == 0 ? rows + 1 : rows; }   // 16-byte groups of a row on distinct banks

// Largest tile (S sub-tiles of 128 rows) whose buffers fit; weights resident if both convs fit next to them.
static int plan_pair(PairArgs& a, bool snake, int smem_cap, int sm_count) {
  const bool single = a.single != 0;
  const int nconv = single ? 1 : 2;
  const ConvGeom g = conv_geom(a.C, a.C, a.K, 0, TB200_PREC_F16);
  if (g.n_ntiles != 1 || g.NT != a.C) return fail(TB200_E_BADARG, "respair: C=%d must be a multiple of 16, <= 256", a.C);
  a.KC = g.KC; a.n_kchunks = g.n_kchunks; a.n_chunks = g.ntaps * g.n_kchunks;
  a.chunk_bytes = (int)(g.chunk_elems * 2);
  a.p1 = (a.K - 1) / 2 * a.dil;
  a.h = (snake && !single) ? 6 : 0;      // rows of u1 the second activation needs on either side (pair only)
  a.p2 = single ? 0 : (a.K - 1) / 2;
  const int esz = a.x_f16 ? 2 : 4;
  // NR = rows of conv1 per tile: a multiple of 32 (an epilogue warp owns 32 accumulator rows), the last 128-row
  // sub-tile at least half used; the largest that fits wins (halo recompute and the per-segment warm-up of the
  // streaming filter are per tile).
  const int nr_max = (g_pair_max_s > 0 ? g_pair_max_s : 8) * kTileM;
  for (int NR = nr_max; NR >= 64; NR -= 32) {
    if (NR % kTileM != 0 && NR % kTileM < 64) continue;
    const int S = (NR + kTileM - 1) / kTileM;
    const int T_out = (NR - 2 * a.h - 2 * a.p2) & ~7;   // multiple of 8: tile origins stay 16-byte aligned
    if (T_out <= 0) continue;
    if (2 * S * a.C > 512) continue;
    if (NR > 64) {
      const int T_prev = (NR - 32 - 2 * a.h - 2 * a.p2) & ~7;
      if (T_prev >= a.L_max) continue;                                             // tile longer than the data
      if (NR > kTileM && (long long)a.B * ((a.L_max + T_out - 1) / T_out) < 2LL * sm_count) continue;   // keep every SM busy
    }
    // operand planes: the rows a conv reads past the last written one belong to discarded output rows only; they
    // alias the next plane / the buffer behind the tile (any finite or non-finite value is harmless there)
    const int R1 = NR + 2 * a.p1, Rp1 = plane_rows(R1);
    const int R2v = NR - 2 * a.h, Rp2 = single ? Rp1 : plane_rows(R2v);   // single conv: the A2 region is A1's second buffer
    const int a1_bytes = (a.C / 8) * Rp1 * 16, a2_bytes = (a.C / 8) * Rp2 * 16;
    const int PX = round_pitch(R1 + 32, esz), x_buf = a.C * PX * esz + 128;     // + look-ahead reads of the last row
    // One X buffer: its load is then exposed between two tiles of the single mode (~5 K cycles of ~25 K, the next tile is
    // prefetched into L2), but two buffers cost a third of the tile height and measured slower (C = 64, K = 3 snake:
    // 2.71 vs 2.51 ms per pair of launches): with every role busy all the time the SM's issue rate is the limit.
    const int x_bufs = 1;
    const int x_bytes = x_bufs * x_buf;
    const int PS = (snake && !single) ? round_pitch(NR + 16, 2) : 0, scr_bytes = PS ? a.C * PS * 2 + 128 : 0;
    const int stage_bytes = (snake ? PairRoles<true>::kEpi : PairRoles<false>::kEpi) * kStageFloats * 4;
    const int fixed = (BNUM + 2 * kPairMaxRing) * 8 + 6 * a.C * 4 + 16 + 256 + stage_bytes;
    const long long avail = (long long)smem_cap - a1_bytes - a2_bytes - x_bytes - scr_bytes - fixed;
    const long long w_total = (long long)nconv * a.n_chunks * a.chunk_bytes;
    int resident = 0, ring = 0;
    if (w_total <= avail && nconv * a.n_chunks <= kPairMaxRing) {
      resident = 1;
      ring = nconv * a.n_chunks;
    } else {
      // streamed weights: a chunk is consumed in a few hundred cycles and refilled from L2 in one to two thousand, so
      // the ring must hold a few chunks and >= 24 KB (a 2-slot ring of 2 KB chunks made the K = 11, C = 32 pairs 2x
      // slower than a smaller tile with resident weights; 3 x 8 KB at C = 64 measured as good as resident weights)
      long long slots = avail / a.chunk_bytes;
      if (slots > 16) slots = 16;
      if (slots > nconv * a.n_chunks) slots = nconv * a.n_chunks;
      if (slots < (g_pair_min_ring > 0 ? g_pair_min_ring : 3) || slots * a.chunk_bytes < 24 * 1024) continue;
      ring = (int)slots;
    }
    a.x_buf_bytes = x_buf; a.x_bufs = x_bufs;
    a.S = S; a.NR = NR; a.T_out = T_out; a.R1 = R1; a.Rp1 = Rp1; a.R2v = R2v; a.Rp2 = Rp2; a.PX = PX; a.PS = PS;
    a.resident = resident; a.ring_slots = ring;
    a.tiles_per_utt = (a.L_max + T_out - 1) / T_out;
    a.total_tiles = a.tiles_per_utt * a.B;
    int cols = 32;
    while (cols < 2 * S * a.C) cols <<= 1;
    a.tmem_cols = cols;
    int off = 0;
    a.a1_off = off; off += a1_bytes;
    a.a2_off = off; off += a2_bytes;
    a.x_off = off; off += x_bytes;
    a.scr_off = off; off += scr_bytes;
    a.w_off = off; off += ring * a.chunk_bytes;
    a.bar_off = off; off += (BNUM + 2 * ring) * 8;
    a.bias_off = off; off += 6 * a.C * 4;   // bias1, bias2, snake parameters of both activations
    a.tmem_off = off; off += 16;
    a.stage_off = off; off += stage_bytes;
    a.smem_total = off;
    return 0;
  }
  return fail(TB200_E_NOSMEM, "respair: no tiling of C=%d K=%d dilation=%d fits in shared memory", a.C, a.K, a.dil);
}

template <bool SNAKE>
static int launch_pair(const PairArgs& a, PairDevice* d, cudaStream_t stream) {
  auto kern = respair_kernel<SNAKE>;
  if (!d->configured[SNAKE]) {
    TB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d->max_smem));
    d->configured[SNAKE] = true;
  }
  int grid = d->sm_count < a.total_tiles ? d->sm_count : a.total_tiles;
  if (grid < 1) grid = 1;
  kern<<<grid, PairRoles<SNAKE>::kThreads, a.smem_total, stream>>>(a);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int respair_trace_read(long long* host_out, int n) {
  if (!g_pair_trace) return fail(TB200_E_BADARG, "trace: TB200_TRACE was not set");
  TB200_CUDA_CHECK(cudaMemcpy(host_out, g_pair_trace, sizeof(long long) * (n < kPairTraceLen ? n : kPairTraceLen), cudaMemcpyDeviceToHost));
  return 0;
}

}  // namespace tb200

using namespace tb200;

extern "C" int tb200_respair(const tb200_respair_params* p, void* stream_v) {
  if (!p || !p->x || !p->y || !p->w1_packed || !p->w2_packed) return fail(TB200_E_BADARG, "respair: null pointer");
  if (p->B <= 0 || p->C <= 0 || p->L_max <= 0) return fail(TB200_E_BADARG, "respair: empty shape");
  const bool snake = p->act == TB200_ACT_AA_SNAKEBETA;
  if (!snake && p->act != TB200_ACT_LEAKY_RELU) return fail(TB200_E_BADARG, "respair: activation must be LeakyReLU or anti-aliased SnakeBeta");
  if (snake && (!p->act1_alpha || !p->act1_beta || !p->act2_alpha || !p->act2_beta)) return fail(TB200_E_BADARG, "respair: snake activation needs alpha/beta");
  if (!snake && (p->act_slope < 0.f || p->act_slope > 1.f)) return fail(TB200_E_BADARG, "respair: LeakyReLU slope must be in [0, 1]");
  if (p->C % 32 || p->C > 128) return fail(TB200_E_BADARG, "respair: C must be 32, 64, 96 or 128 (got %d)", p->C);
  if (PairRoles<true>::kProd % (p->C / 32)) return fail(TB200_E_BADARG, "respair: C=%d does not divide the producer warps", p->C);
  if (p->K < 1 || p->K > kMaxTaps || !(p->K & 1) || p->dilation < 1) return fail(TB200_E_BADARG, "respair: K must be odd, dilation >= 1");
  const int esz_x = p->x_dtype == TB200_F16 ? 2 : 4, esz_y = p->y_dtype == TB200_F16 ? 2 : 4;
  const int unit = 16 / esz_x;
  if ((reinterpret_cast<uintptr_t>(p->x) & 15) || p->x_ld % unit || p->x_bs % unit || p->x_ld < (p->L_max + unit - 1) / unit * unit)
    return fail(TB200_E_BADARG, "respair: x must be 16-byte aligned with row pitch and batch stride multiples of %d elements", unit);
  const int unit_y = 16 / esz_y;
  if ((reinterpret_cast<uintptr_t>(p->y) & 15) || p->y_ld % unit_y || p->y_bs % unit_y || p->y_ld < (p->L_max + unit_y - 1) / unit_y * unit_y)
    return fail(TB200_E_BADARG, "respair: y must be 16-byte aligned with row pitch and batch stride multiples of %d elements", unit_y);
  if (p->x == p->y) return fail(TB200_E_BADARG, "respair: in-place operation is not supported (tiles read their neighbours' halos)");
  // offsets are 64-bit across utterances; one utterance (C rows of the row pitch) must stay below 2^31 elements
  const long long lim = 1LL << 31;
  if ((long long)p->C * p->x_ld >= lim || (long long)p->C * p->y_ld >= lim || p->x_bs < 0 || p->y_bs < 0)
    return fail(TB200_E_BADARG, "respair: one utterance (C x row pitch) must be below 2^31 elements, batch strides >= 0");
  PairDevice* d = nullptr;
  int rc = pair_device(d);
  if (rc) return rc;
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.x = p->x; a.y = p->y; a.len = p->len;
  a.bias1 = p->bias1; a.bias2 = p->bias2;
  a.alpha1 = p->act1_alpha; a.beta1 = p->act1_beta; a.alpha2 = p->act2_alpha; a.beta2 = p->act2_beta;
  a.w1 = p->w1_packed; a.w2 = p->w2_packed;
  a.x_bs = p->x_bs; a.y_bs = p->y_bs; a.x_ld = p->x_ld; a.y_ld = p->y_ld;
  a.B = p->B; a.C = p->C; a.L_max = p->L_max; a.K = p->K; a.dil = p->dilation;
  a.x_f16 = p->x_dtype == TB200_F16; a.y_f16 = p->y_dtype == TB200_F16;
  a.slope = p->act_slope; a.out_alpha = p->out_alpha; a.res_beta = p->res_beta; a.accumulate = p->accumulate;
  a.res = p->x; a.r_bs = p->x_bs; a.r_ld = p->x_ld; a.r_f16 = a.x_f16; a.single = 0;
  rc = plan_pair(a, snake, d->max_smem, d->sm_count);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  a.dbg_skip = g_pair_skip;
  a.trace = nullptr;
  if (g_pair_trace_on) {
    if (!g_pair_trace) TB200_CUDA_CHECK(cudaMalloc(&g_pair_trace, kPairTraceLen * sizeof(long long)));
    TB200_CUDA_CHECK(cudaMemsetAsync(g_pair_trace, 0, kPairTraceLen * sizeof(long long), stream));
    a.trace = g_pair_trace;
  }
  if (g_pair_debug)
    fprintf(stderr, "tb200 respair plan: C=%d K=%d dil=%d snake=%d L=%d x_f16=%d -> S=%d T_out=%d R1=%d tiles=%d %s ring=%d smem=%d tmem=%d\n",
            a.C, a.K, a.dil, (int)snake, a.L_max, a.x_f16, a.S, a.T_out, a.R1, a.total_tiles, a.resident ? "resident" : "streamed",
            a.ring_slots, a.smem_total, a.tmem_cols);
  return snake ? launch_pair<true>(a, d, stream) : launch_pair<false>(a, d, stream);
}

// One convolution through the same pipeline (TMA-fed X tile, smem-fed staging, double-buffered operand tiles and
// accumulators, vectorised epilogue): same parameter block as tb200_conv1d.
extern "C" int tb200_conv1d_staged(const tb200_conv1d_params* p, void* stream_v) {
  if (!p || !p->x || !p->y || !p->w_packed) return fail(TB200_E_BADARG, "conv1d_staged: null pointer");
  if (p->B <= 0 || p->C_in <= 0 || p->L_in_max <= 0) return fail(TB200_E_BADARG, "conv1d_staged: empty shape");
  const bool snake = p->act == TB200_ACT_AA_SNAKEBETA;
  if (p->precision != TB200_PREC_F16 || p->transposed_stride != 0 || p->out_act != TB200_OUT_NONE || p->C_in != p->C_out ||
      p->C_in % 32 || p->C_in > 128 || !(p->K & 1) || p->K > kMaxTaps || p->dilation < 1 || p->pad != (p->K - 1) / 2 * p->dilation ||
      (!snake && p->act != TB200_ACT_LEAKY_RELU))
    return fail(TB200_E_BADARG, "conv1d_staged: needs fp16 operands, C_in == C_out in {32,64,96,128}, an odd 'same'-padded kernel, "
                                "LeakyReLU or anti-aliased SnakeBeta in front and no output activation");
  if (snake && (!p->act_alpha || !p->act_beta)) return fail(TB200_E_BADARG, "conv1d_staged: snake activation needs alpha/beta");
  if (!snake && (p->act_slope < 0.f || p->act_slope > 1.f)) return fail(TB200_E_BADARG, "conv1d_staged: LeakyReLU slope must be in [0, 1]");
  if (PairRoles<true>::kProd % (p->C_in / 32)) return fail(TB200_E_BADARG, "conv1d_staged: C=%d does not divide the producer warps", p->C_in);
  const int ux = p->x_dtype == TB200_F16 ? 8 : 4, uy = p->y_dtype == TB200_F16 ? 8 : 4, ur = p->r_dtype == TB200_F16 ? 8 : 4;
  auto aligned = [&](const void* ptr, int64_t bs, int32_t ld, int unit) {
    return !(reinterpret_cast<uintptr_t>(ptr) & 15) && ld % unit == 0 && bs % unit == 0 && ld >= (p->L_in_max + unit - 1) / unit * unit;
  };
  if (!aligned(p->x, p->x_bs, p->x_ld, ux) || !aligned(p->y, p->y_bs, p->y_ld, uy) || (p->residual && !aligned(p->residual, p->r_bs, p->r_ld, ur)))
    return fail(TB200_E_BADARG, "conv1d_staged: x, y and residual must be 16-byte aligned with 16-byte multiples as row pitch and batch stride");
  if (p->x == p->y) return fail(TB200_E_BADARG, "conv1d_staged: in-place operation is not supported (tiles read their neighbours' halos)");
  const long long lim = 1LL << 31;
  if ((long long)p->C_in * p->x_ld >= lim || (long long)p->C_out * p->y_ld >= lim || (p->residual && (long long)p->C_out * p->r_ld >= lim) ||
      p->x_bs < 0 || p->y_bs < 0 || (p->residual && p->r_bs < 0))
    return fail(TB200_E_BADARG, "conv1d_staged: one utterance (C x row pitch) must be below 2^31 elements, batch strides >= 0");
  PairDevice* d = nullptr;
  int rc = pair_device(d);
  if (rc) return rc;
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.single = 1;
  a.x = p->x; a.y = p->y; a.len = p->len_in;
  a.bias1 = p->bias; a.alpha1 = p->act_alpha; a.beta1 = p->act_beta;
  a.w1 = p->w_packed; a.w2 = p->w_packed;
  a.x_bs = p->x_bs; a.y_bs = p->y_bs; a.x_ld = p->x_ld; a.y_ld = p->y_ld;
  a.B = p->B; a.C = p->C_in; a.L_max = p->L_in_max; a.K = p->K; a.dil = p->dilation;
  a.x_f16 = p->x_dtype == TB200_F16; a.y_f16 = p->y_dtype == TB200_F16;
  a.slope = p->act_slope; a.out_alpha = p->out_alpha; a.res_beta = p->residual ? p->res_beta : 0.f; a.accumulate = p->accumulate;
  a.res = p->residual; a.r_bs = p->r_bs; a.r_ld = p->r_ld; a.r_f16 = p->r_dtype == TB200_F16;
  rc = plan_pair(a, snake, d->max_smem, d->sm_count);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  a.dbg_skip = g_pair_skip;
  a.trace = nullptr;
  if (g_pair_trace_on) {
    if (!g_pair_trace) TB200_CUDA_CHECK(cudaMalloc(&g_pair_trace, kPairTraceLen * sizeof(long long)));
    TB200_CUDA_CHECK(cudaMemsetAsync(g_pair_trace, 0, kPairTraceLen * sizeof(long long), stream));
    a.trace = g_pair_trace;
  }
  if (g_pair_debug)
    fprintf(stderr, "tb200 staged conv plan: C=%d K=%d dil=%d snake=%d L=%d x_f16=%d -> S=%d T_out=%d R1=%d tiles=%d %s ring=%d smem=%d tmem=%d\n",
            a.C, a.K, a.dil, (int)snake, a.L_max, a.x_f16, a.S, a.T_out, a.R1, a.total_tiles, a.resident ? "resident" : "streamed",
            a.ring_slots, a.smem_total, a.tmem_cols);
  return snake ? launch_pair<true>(a, d, stream) : launch_pair<false>(a, d, stream);
}

extern "C" int tb200_respair_trace_read(int64_t* host_out, int32_t n) {
  if (!host_out || n <= 0) return fail(TB200_E_BADARG, "trace: bad argument");
  return respair_trace_read(reinterpret_cast<long long*>(host_out), n);
}
