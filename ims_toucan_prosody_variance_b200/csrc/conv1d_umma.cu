// conv1d_umma.cu -- implicit-GEMM Conv1d / ConvTranspose1d on the sm_100a tensor cores.
//
//   GEMM view (per utterance b):   D[t, n] = sum_{tap j} sum_{ci} A_j[t, ci] * W_j[n, ci]
//     M = time (128 rows per accumulator == 128 TMEM lanes), N = output channels (<= 256 per
//     accumulator), K = input channels, one pass over K per tap.
//   A operand: ONE staged tile of ACT(x) for rows [t0 - halo_l, t0 + S*128 + halo_r), written by the
//     producer warps in the canonical K-major SWIZZLE_NONE layout with 16 B per row
//     ([Cin/E][R][E] elements).  Rows are linear in that layout, so tap j of sub-tile s is the same
//     tile read through a descriptor whose start address is shifted by (s*128 + tap_off[j]) rows: no
//     im2col, no per-tap reload.  The activation (LeakyReLU, or BigVGAN's anti-aliased SnakeBeta =
//     2x kaiser-sinc up, snake, 2x down) is fused into the staging and never touches HBM.
//   B operand: weight blocks [KC/E][NT][E] pre-packed at load time, brought in by the TMA engine
//     (cp.async.bulk + mbarrier); resident for the CTA's lifetime when they fit, else streamed
//     through a ring released by tcgen05.commit.  One block serves all S sub-tiles.
//   D: fp32 in TMEM, double buffered; epilogue = tcgen05.ld -> bias / activation / scale / residual /
//     accumulate -> coalesced NCL stores (a warp writes 32 consecutive time steps of one channel).
//
// Warp roles (one persistent CTA per SM; W worker warps, then the MMA warp and the loader warp):
//   warps 0 .. nP-1   producers : stage A[buf] (double buffered), arrive a_full
//   warps nP .. W-1   epilogue  : drain accumulator buffer, arrive acc_empty  (4 or 8 warps; any 4 consecutive
//                                 warps cover the four TMEM lane quarters)
//   warp  W           MMA issuer: one elected thread issues tcgen05.mma, commits to a_empty / acc_full / w_empty
//   warp  W+1         weight loader (TMA bulk copies)
// so staging of tile i+1, the MMAs of tile i and the epilogue of tile i-1 overlap.  Two kernels per operand type:
//   pointwise prologues: W = 14 (16 warps, 128 registers); nP = 6 (+8 epilogue) when the epilogue fetches residual /
//                        accumulate rows, 10 (+4) when it only stores
//   snake prologue:      W = 22 (24 warps, 80 registers); nP = 14 (+8): the staging is FMA-issue bound
// TB200_TRACE=1 records clock64() stamps of every role per tile (tools/conv_micro.py prints the timeline).
//
// Tiles past an utterance's length are skipped, rows past it are staged as zeros (the zero padding a
// batch-1 reference call sees).
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "conv_common.cuh"
#include "snake_stream.cuh"

#ifndef TB200_ROLE_INLINE
#define TB200_ROLE_INLINE __forceinline__
#endif
#ifndef TB200_LDMODE
#define TB200_LDMODE 0   // 0: ld.global.nc (LDG.E.CONSTANT)   1: ld.global.cg (L2 only, no L1 allocation)
#endif
#ifndef TB200_MAX_S
#define TB200_MAX_S 8
#endif

namespace tb200 {

// Worker warps = producers [0, a.n_prod) + epilogue [a.n_prod, W); warp W issues the MMAs, warp W+1 loads weights.
//   pointwise staging family: W = 14 (16 warps, 128 registers per thread)
//   snake staging family:     W = 22 (24 warps,  80 registers per thread): the anti-aliased snake is a chain of
//   dependent FMAs (one instruction per ~4.5 cycles per warp), so it wants warps, not registers.
#ifndef TB200_SNAKE_EPI_BATCH
#define TB200_SNAKE_EPI_BATCH 1
#endif
#ifndef TB200_PW_WORKERS
#define TB200_PW_WORKERS 14
#endif
#ifndef TB200_SNAKE_WORKERS
#define TB200_SNAKE_WORKERS 22
#endif
template <bool SNAKE>
struct Roles {
  static constexpr int kWorkers = SNAKE ? TB200_SNAKE_WORKERS : TB200_PW_WORKERS;
  static constexpr int kMma = kWorkers;
  static constexpr int kLoad = kWorkers + 1;
  static constexpr int kThreads = (kWorkers + 2) * 32;
};
constexpr int kMaxProdWarps = 18;
constexpr int kBiasCache = 2048;            // floats of (bias * out_alpha) cached in shared memory for the plain epilogue
constexpr int kMaxRing = 64;

template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<__half> {
  static constexpr int kEpc = 8;
  static constexpr bool kTf32 = false;
};
template <>
struct ElemTraits<float> {
  static constexpr int kEpc = 4;
  static constexpr bool kTf32 = true;
};

// Optional per-tile timeline of CTA 0 (TB200_TRACE=1, debugging aid): clock64() stamps per tile.
//  0 producer: a_empty acquired   1 producer: tile staged     2 MMA: a_full acquired   3 MMA: all MMAs issued
//  4 epilogue: acc_full acquired  5 epilogue: tile drained    6 MMA: acc_empty acquired
constexpr int kTraceTiles = 96;
constexpr int kTraceCtas = 160;   // after the tile stamps: elapsed cycles and %smid of every CTA
constexpr int kTraceWarps = 32;   // then (start, end) of every producer warp of CTA 0 while staging its 6th A buffer
constexpr int kTraceLen = kTraceTiles * 8 + kTraceCtas + 2 * kTraceWarps;
__device__ __forceinline__ void trace(const ConvArgs& a, int slot, uint32_t idx) {
  if (a.trace && blockIdx.x == 0 && idx < (uint32_t)kTraceTiles) a.trace[idx * 8 + slot] = clock64();
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A-tile store policy: 16-byte group g, row r -> canonical K-major SWIZZLE_NONE position.
template <typename T>
struct UmmaStore {
  T* base;
  int R;
  __device__ __forceinline__ void operator()(int g, int r, const float (&v)[ElemTraits<T>::kEpc]) const {
    T* dst = base + ((long long)g * R + r) * ElemTraits<T>::kEpc;
    if constexpr (ElemTraits<T>::kEpc == 8) {
      uint4 u;
      u.x = f16x2_sat(v[0], v[1]);
      u.y = f16x2_sat(v[2], v[3]);
      u.z = f16x2_sat(v[4], v[5]);
      u.w = f16x2_sat(v[6], v[7]);
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      *reinterpret_cast<float4*>(dst) = make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// producer, pointwise activations: lane = time (coalesced 128-byte rows), branch-free loads with
// clamped addresses, the next task's loads in flight while the current one is converted/stored.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ TB200_ROLE_INLINE void stage_pointwise_mlp(const ConvArgs& a, int b, int t_lo, int g0, int ng, int len, T* smA,
                                                    int pw, int lane) {
  const int kProdWarps = a.n_prod;
  constexpr int E = ElemTraits<T>::kEpc;
  const int R = a.R;
  const int nrb = (R + 31) / 32;
  const int ntask = ng * nrb;
  const long long xb = (long long)b * a.x_bs;
  const UmmaStore<T> st{smA, R};
  float cur[E], nxt[E];
  auto issue = [&](int task, float (&r)[E]) {
    const int g = task / nrb, rb = task - g * nrb;
    const int t = min(max(t_lo + rb * 32 + lane, 0), len - 1);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int c = min((g0 + g) * E + e, a.Cin - 1);
      r[e] = load_x(a.x, a.x_f16, xb + (long long)c * a.x_ld + t);
    }
  };
  int task = pw;
  if (task < ntask) issue(task, cur);
  for (; task < ntask; task += kProdWarps) {
    if (task + kProdWarps < ntask) issue(task + kProdWarps, nxt);
    const int g = task / nrb, rb = task - g * nrb;
    const int r = rb * 32 + lane;
    const int t = t_lo + r;
    const bool valid = (t >= 0) && (t < len);
    float v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float av = apply_pointwise(cur[e], a.act, a.slope);
      v[e] = (valid && (g0 + g) * E + e < a.Cin) ? av : 0.f;
    }
    if (r < R) st(g, r, v);
#pragma unroll
    for (int e = 0; e < E; ++e) cur[e] = nxt[e];
  }
}

// ---------------------------------------------------------------------------------------------
// producer, pointwise activations, vectorised: a lane owns V consecutive time steps (one 16-byte
// load per channel: V = 4 fp32 / 8 fp16) of the E channels of one 16-byte operand group, so a warp
// reads 512 contiguous bytes per channel row and a thread keeps E x 16 bytes in flight (x2 with the
// next task prefetched): ~3 instructions per element instead of ~15 for the scalar path.
// Loads start at a 16-byte aligned time index t_al <= t_lo; vectors entirely outside [0, len) are
// not loaded (zeros), elements past len inside the last vector are masked.
// Preconditions (host: a.pw_vec): x base 16-byte aligned, x_ld and x_bs multiples of V.
// ---------------------------------------------------------------------------------------------
struct Vec16 {
  uint4 u;
};

__device__ __forceinline__ uint32_t act_half2(uint32_t v, int act, __half2 slope2) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  if (act == TB200_ACT_LEAKY_RELU) h = __hmax2(h, __hmul2(h, slope2));
  else if (act == TB200_ACT_RELU) h = __hmax2(h, __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}

template <typename T, bool XF16>
__device__ TB200_ROLE_INLINE void stage_pointwise_vec(const ConvArgs& a, int b, int t_lo, int g0, int ng, int len, T* smA,
                                                    int pw, int lane) {
  const int kProdWarps = a.n_prod;
  constexpr int E = ElemTraits<T>::kEpc;
  constexpr int V = XF16 ? 8 : 4;   // time steps per 16-byte load
  constexpr int CH = 4;             // channels per task: half (fp16 operands) or all (tf32) of a 16-byte operand group
  constexpr int SUB = E / CH;
  static_assert(!(XF16 && E != 8), "fp16 activations feed fp16 operands only");
  const int R = a.R;
  const int shift = ((t_lo % V) + V) % V;       // t_lo - t_al
  const int t_al = t_lo - shift;
  const int nq = (R + shift + V - 1) / V;
  const int nqb = (nq + 31) / 32;
  const int ntask = ng * SUB * nqb;
  const long long xb = (long long)b * a.x_bs;
  const bool simple_act = a.act == TB200_ACT_NONE || a.act == TB200_ACT_LEAKY_RELU || a.act == TB200_ACT_RELU;
  const __half2 slope2 = __float2half2_rn(a.slope);
  uint4 bufa[CH], bufb[CH];
  auto issue = [&](int task, uint4 (&v)[CH]) {
    const int gs = task / nqb, q = (task - gs * nqb) * 32 + lane;
    const int tq = t_al + V * q;
    const bool ok = q < nq && tq >= 0 && tq < len;
#pragma unroll
    for (int e = 0; e < CH; ++e) {
      const int c = g0 * E + gs * CH + e;
      if (ok && c < a.Cin) {
        const long long idx = xb + (long long)c * a.x_ld + tq;
        const uint4* src = XF16 ? reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.x) + idx)
                                : reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(a.x) + idx);
        v[e] = TB200_LDMODE ? __ldcg(src) : __ldg(src);
      } else {
        v[e] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  auto process = [&](int task, uint4 (&cur)[CH]) {
    const int gs = task / nqb, q = (task - gs * nqb) * 32 + lane;
    const int g = gs / SUB, sub = gs - g * SUB;
    const int tq = t_al + V * q;
    const int r0 = V * q - shift;
    if (q < nq) {
      T* dst = smA + ((long long)g * R + r0) * E + sub * CH;   // row r0 of this task's CH channels
      const bool full = tq >= 0 && tq + V <= len;
      if constexpr (XF16) {
        if (full && simple_act) {
          // 8 (time) x 4 (channel) halves: activation on time pairs, then a register transpose
#pragma unroll
          for (int e = 0; e < CH; ++e) {
            cur[e].x = act_half2(cur[e].x, a.act, slope2);
            cur[e].y = act_half2(cur[e].y, a.act, slope2);
            cur[e].z = act_half2(cur[e].z, a.act, slope2);
            cur[e].w = act_half2(cur[e].w, a.act, slope2);
          }
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const int r = r0 + i;
            const uint32_t sel = (i & 1) ? 0x7632u : 0x5410u;
            auto word = [&](int e) { return (i >> 1) == 0 ? cur[e].x : (i >> 1) == 1 ? cur[e].y : (i >> 1) == 2 ? cur[e].z : cur[e].w; };
            uint2 o;
            o.x = __byte_perm(word(0), word(1), sel);
            o.y = __byte_perm(word(2), word(3), sel);
            if (r >= 0 && r < R) *reinterpret_cast<uint2*>(dst + (long long)i * E) = o;
          }
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const int r = r0 + i, t = tq + i;
            const bool valid = t >= 0 && t < len;
            float v[CH];
#pragma unroll
            for (int e = 0; e < CH; ++e) {
              const uint32_t word = (i >> 1) == 0 ? cur[e].x : (i >> 1) == 1 ? cur[e].y : (i >> 1) == 2 ? cur[e].z : cur[e].w;
              const __half2 h2 = *reinterpret_cast<const __half2*>(&word);
              const float xv = (i & 1) ? __high2float(h2) : __low2float(h2);
              v[e] = valid ? apply_pointwise(xv, a.act, a.slope) : 0.f;
            }
            if (r >= 0 && r < R) {
              uint2 o;
              o.x = f16x2_sat(v[0], v[1]);
              o.y = f16x2_sat(v[2], v[3]);
              *reinterpret_cast<uint2*>(dst + (long long)i * E) = o;
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const int r = r0 + i, t = tq + i;
          const bool valid = full || (t >= 0 && t < len);
          float v[CH];
#pragma unroll
          for (int e = 0; e < CH; ++e) {
            const uint32_t word = i == 0 ? cur[e].x : i == 1 ? cur[e].y : i == 2 ? cur[e].z : cur[e].w;
            const float xv = valid ? __uint_as_float(word) : 0.f;   // every supported activation maps 0 -> 0
            if (a.act == TB200_ACT_LEAKY_RELU) v[e] = fmaxf(xv, xv * a.slope);   // 0 <= slope <= 1 (checked on the host)
            else if (a.act == TB200_ACT_NONE) v[e] = xv;
            else v[e] = apply_pointwise(xv, a.act, a.slope);
          }
          if (r >= 0 && r < R) {
            if constexpr (E == 8) {
              uint2 o;
              o.x = f16x2_sat(v[0], v[1]);
              o.y = f16x2_sat(v[2], v[3]);
              *reinterpret_cast<uint2*>(dst + (long long)i * E) = o;
            } else {
              *reinterpret_cast<float4*>(dst + (long long)i * E) =
                  make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
            }
          }
        }
      }
    }
  };
  // Batches of 2 tasks: the 2 x CH 16-byte loads of a batch are issued back to back, then both tasks are converted.
  // (A rotating scheme with one task of look-ahead pays one full memory latency per task -- the per-tile trace
  // showed 3.4 K cycles per task; a batch pays it once per 2 tasks with the same register footprint.  Larger
  // batches spill, and with the L1 carved out for shared memory a spill costs a DRAM round trip.)
  const int P = kProdWarps;
  for (int task0 = pw; task0 < ntask; task0 += 2 * P) {
    issue(task0, bufa);
    if (task0 + P < ntask) issue(task0 + P, bufb);
    process(task0, bufa);
    if (task0 + P < ntask) process(task0 + P, bufb);
  }
}

template <typename T>
__device__ TB200_ROLE_INLINE void stage_aa_channel(const ConvArgs& a, int b, int t_lo, int cb0, int ncb, int nseg, int len,
                                                 T* smA, int pw, int lane) {
  constexpr int E = ElemTraits<T>::kEpc;
  const int R = a.R;
  const int seg_rows = (R + nseg - 1) / nseg;
  // Interior segment boundaries sit where (t + 6) % 8 == 0 (absolute time): the first output of the segment is then
  // the first output of a steady 8-step block and its last output the last of one, so only the tile's first and last
  // segment run partial (checked) blocks.  Needs seg_rows >= 8 to keep the boundaries increasing.
  auto seg_start = [&](int sgm) {
    if (sgm <= 0) return 0;
    if (sgm >= nseg) return R;
    int r = sgm * seg_rows;
    if (seg_rows >= 8) {
      const int ph = (t_lo + r + 6) & 7;          // two's complement: correct for negative t_lo as well
      r += ph > 4 ? 8 - ph : -ph;
    }
    return min(max(r, 0), R);
  };
  const long long xb = (long long)b * a.x_bs;
  for (int task = pw; task < ncb * nseg; task += a.n_prod) {
    const int cb = task / nseg, seg = task - cb * nseg;
    const int r_beg = seg_start(seg);
    const int r_end = seg_start(seg + 1);
    if (r_beg >= r_end) continue;
    const int cl = cb * 32 + lane;          // channel inside this staged panel
    const int c = cb0 * 32 + cl;            // absolute input channel
    const int t_beg = t_lo + r_beg, t_end = t_lo + r_end;
    T* dst = smA + ((long long)(cl / E) * R) * E + (cl % E);
    if (t_beg >= len || t_end <= 0) {       // segment entirely outside the utterance: the conv's zero padding
      for (int t = t_beg; t < t_end; ++t) dst[(long long)(t - t_lo) * E] = to_operand<T>(0.f);
      continue;
    }
    const long long row = xb + (long long)c * a.x_ld;
    const bool edge = (((t_beg - 9) & ~7) < 0) || (t_end + 32 > len);   // warp-uniform
    const float ea = __expf(__ldg(a.alpha + c));
    const float ib = 1.0f / (__expf(__ldg(a.beta + c)) + 1e-9f);
    if (a.x_f16) {
      if (edge) aa_channel_task<T, true, true, false>(a.x, row, ea, ib, t_lo, t_beg, t_end, len, dst);
      else aa_channel_task<T, false, true, false>(a.x, row, ea, ib, t_lo, t_beg, t_end, len, dst);
    } else {
      if (edge) aa_channel_task<T, true, false, false>(a.x, row, ea, ib, t_lo, t_beg, t_end, len, dst);
      else aa_channel_task<T, false, false, false>(a.x, row, ea, ib, t_lo, t_beg, t_end, len, dst);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct WsLayout {  // shared-memory carve-up, computed identically on host and device
  int a_off, w_off, bar_off, tmem_off, scratch_off, bias_off, total;
};
__host__ __device__ inline WsLayout ws_layout(const ConvArgs& a) {
  WsLayout l;
  l.a_off = 0;
  l.w_off = a.a_bufs * a.a_bytes;
  l.bar_off = l.w_off + a.ring_slots * a.chunk_bytes;
  const int nbars = 2 * a.ring_slots + 2 * a.a_bufs + 2 * a.acc_bufs;
  l.tmem_off = l.bar_off + nbars * 8;
  l.scratch_off = l.tmem_off + 16;
  l.bias_off = l.scratch_off + kMaxProdWarps * 2 * kAaScratch * 4;
  l.total = l.bias_off + kBiasCache * 4;
  return l;
}

// ---------------------------------------------------------------------------------------------
// epilogue bodies.  A warp owns TMEM lanes [32q, 32q+32) = 32 consecutive output rows (time steps) and
// walks 16-column (channel) slabs: one tcgen05.ld per slab, then for each column a 128-byte coalesced
// row of 32 time steps.  All global loads of a slab (residual / accumulate / bias) are issued before
// the TMEM load is waited on, so their latency is paid once per slab.
// ---------------------------------------------------------------------------------------------
struct EpiAux {           // the one auxiliary input of the plain epilogue: residual (fp32) or the old y (accumulate)
  const void* ptr;
  uint32_t base;          // element offset of this utterance
  uint32_t ld;
  int f16;
  float beta;
};

// Plain epilogue: regular conv, N_total % 16 == 0, no output activation, NAUX auxiliary inputs (residual and/or
// the old y), and every tensor small enough for 32-bit element offsets (host: a.epi_fast).  Addresses are
// (uniform 64-bit base) + (32-bit offset), one integer add per access.
// BATCH = items whose auxiliary loads are issued together (2 with 128 registers per thread, 1 in the 80-register
// snake kernel, which has 8 epilogue warps to hide the latency instead).
template <int NAUX, int BATCH>
__device__ TB200_ROLE_INLINE void epilogue_plain(const ConvArgs& a, const float* bias_s, const EpiAux& ax0, const EpiAux& ax1, uint32_t tm0,
                                                 int nsub, int slabs, int slab0, int slab_step, int nt, int m_base,
                                                 int len_out, uint32_t ybase) {
  // work items = (slab, sub-tile) pairs, slab-major.  The auxiliary rows of the next item(s) are requested
  // before item k is finished (rotating register sets), so a warp keeps 32 row loads (4 KB) in flight.
  const int nsl = (slabs - slab0 + slab_step - 1) / slab_step;
  const int nitems = nsl * nsub;
  float* yf = reinterpret_cast<float*>(a.y);
  __half* yh = reinterpret_cast<__half*>(a.y);
  const uint32_t y_ld = (uint32_t)a.y_ld;
  const bool relu = a.out_act == TB200_OUT_RELU;
  const bool y_f16 = a.y_f16 != 0;
  const float out_alpha = a.out_alpha, beta0 = ax0.beta, beta1 = ax1.beta;
  auto fetch1 = [&](int k, const EpiAux& ax, float (&r)[16]) {
    const int sl = k / nsub, sub = k - sl * nsub;
    const uint32_t n0 = (uint32_t)(nt * a.NT + (slab0 + sl * slab_step) * 16);
    uint32_t off = ax.base + n0 * ax.ld + (uint32_t)min(m_base + sub * kTileM, len_out - 1);
    if (ax.f16) {
      const __half* p = reinterpret_cast<const __half*>(ax.ptr);
#pragma unroll
      for (int i = 0; i < 16; ++i, off += ax.ld) r[i] = __half2float(TB200_LDMODE ? __ldcg(p + off) : p[off]);
    } else {
      const float* p = reinterpret_cast<const float*>(ax.ptr);
#pragma unroll
      for (int i = 0; i < 16; ++i, off += ax.ld) r[i] = TB200_LDMODE ? __ldcg(p + off) : p[off];
    }
  };
  auto finish_item = [&](int k, const float (&r0)[16], const float (&r1)[16]) {
    const int sl = k / nsub, sub = k - sl * nsub;
    const int s = slab0 + sl * slab_step;
    const uint32_t n0 = (uint32_t)(nt * a.NT + s * 16);
    float bias[16];   // bias * out_alpha, cached in shared memory by the whole CTA at kernel start
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n0 + 4 * i);
      bias[4 * i] = b4.x; bias[4 * i + 1] = b4.y; bias[4 * i + 2] = b4.z; bias[4 * i + 3] = b4.w;
    }
    const int m = m_base + sub * kTileM;
    const bool row_ok = m < len_out;
    uint32_t v[16];
    __syncwarp();
    tmem_ld_x16(tm0 + (uint32_t)(sub * a.NT + s * 16), v);
    tmem_ld_wait();
    // all per-item switches (ReLU, output type, row inside the utterance) are warp-uniform or hoisted out of the
    // 16-column loops: the per-element work is 2 FFMA + address add + store
    float val[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) val[i] = fmaf(__uint_as_float(v[i]), out_alpha, bias[i]);
    if (relu) {   // alpha * relu(x) == relu(alpha * x) for alpha >= 0 (host-checked)
#pragma unroll
      for (int i = 0; i < 16; ++i) val[i] = fmaxf(val[i], 0.f);
    }
    if constexpr (NAUX >= 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) val[i] = fmaf(beta0, r0[i], val[i]);
    }
    if constexpr (NAUX >= 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) val[i] = fmaf(beta1, r1[i], val[i]);
    }
    if (row_ok) {
      const uint32_t yo = ybase + n0 * y_ld + (uint32_t)m;
      if (y_f16) {
        __half* yp = yh + yo;
#pragma unroll
        for (int i = 0; i < 16; ++i, yp += y_ld) *yp = f16_sat(val[i]);
      } else {
        float* yp = yf + yo;
#pragma unroll
        for (int i = 0; i < 16; ++i, yp += y_ld) *yp = val[i];
      }
    }
  };
  float ra[16], rb[16];
  if constexpr (NAUX == 1) {
    // batches of 2 items: 32 row loads requested back to back, then both items drained (one memory latency per
    // 2 items; one item of look-ahead pays it per item -- measured 3.3 K cycles per 16-column item)
    if constexpr (BATCH == 2) {
      for (int k = 0; k < nitems; k += 2) {
        fetch1(k, ax0, ra);
        if (k + 1 < nitems) fetch1(k + 1, ax0, rb);
        finish_item(k, ra, ra);
        if (k + 1 < nitems) finish_item(k + 1, rb, rb);
      }
    } else {
      for (int k = 0; k < nitems; ++k) {
        fetch1(k, ax0, ra);
        finish_item(k, ra, ra);
      }
    }
  } else if constexpr (NAUX == 2) {
    // two inputs (rare: last pair of the 2nd/3rd residual block of a stage): 32 loads per item, no look-ahead
    for (int k = 0; k < nitems; ++k) {
      fetch1(k, ax0, ra);
      fetch1(k, ax1, rb);
      finish_item(k, ra, rb);
    }
  } else {
    for (int k = 0; k < nitems; ++k) finish_item(k, ra, ra);
  }
}

// Transposed conv with nothing but bias and scale behind it (every up-sampler of the vocoders).  Column n = co*UP + phase
// of input row m lands at t = m*UP - UP/2 + phase.  UP is a template parameter, so channel and phase of each of the
// 16 columns of a slab are compile-time (16 % UP == 0) or uniform arithmetic (UP == 6); bias*alpha per column comes from
// the shared-memory cache.  When UP/2 is even, phases (p, p+1) with p even are an aligned pair of one lane: 8-byte
// (fp32) / 4-byte (fp16) stores (PAIR, host-checked alignment of y).
template <int UP, bool PAIR>
__device__ TB200_ROLE_INLINE void epilogue_up(const ConvArgs& a, const float* bias_s, uint32_t tm0, int nsub, int slabs,
                                              int slab0, int slab_step, int nt, int m_base, int len_out, uint32_t ybase) {
  float* yf = reinterpret_cast<float*>(a.y);
  __half* yh = reinterpret_cast<__half*>(a.y);
  const uint32_t y_ld = (uint32_t)a.y_ld;
  const bool y_f16 = a.y_f16 != 0;
  const float out_alpha = a.out_alpha;
  for (int sub = 0; sub < nsub; ++sub) {
    const int t_first = (m_base + sub * kTileM) * UP - UP / 2;
    for (int s = slab0; s < slabs; s += slab_step) {
      uint32_t v[16];
      __syncwarp();
      tmem_ld_x16(tm0 + (uint32_t)(sub * a.NT + s * 16), v);
      const uint32_t n0 = (uint32_t)(nt * a.NT + s * 16);
      float val[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n0 + 4 * i);
        val[4 * i] = b4.x; val[4 * i + 1] = b4.y; val[4 * i + 2] = b4.z; val[4 * i + 3] = b4.w;
      }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) val[i] = fmaf(__uint_as_float(v[i]), out_alpha, val[i]);
      if constexpr (16 % UP == 0) {
        const uint32_t co0 = n0 / UP;
#pragma unroll
        for (int g = 0; g < 16 / UP; ++g) {
          const uint32_t rowoff = ybase + (co0 + g) * y_ld;
          if constexpr (PAIR && (UP / 2) % 2 == 0) {
#pragma unroll
            for (int ph = 0; ph < UP; ph += 2) {
              const int t = t_first + ph;                  // even; len_out = len*UP is even: t < len_out covers t + 1
              if (t >= 0 && t < len_out) {
                if (y_f16) *reinterpret_cast<uint32_t*>(yh + rowoff + t) = f16x2_sat(val[g * UP + ph], val[g * UP + ph + 1]);
                else *reinterpret_cast<float2*>(yf + rowoff + t) = make_float2(val[g * UP + ph], val[g * UP + ph + 1]);
              }
            }
          } else {
#pragma unroll
            for (int ph = 0; ph < UP; ++ph) {
              const int t = t_first + ph;
              if (t >= 0 && t < len_out) {
                if (y_f16) yh[rowoff + t] = f16_sat(val[g * UP + ph]);
                else yf[rowoff + t] = val[g * UP + ph];
              }
            }
          }
        }
      } else {
        uint32_t co = n0 / UP;
        int ph = (int)(n0 - co * UP);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int t = t_first + ph;
          if (t >= 0 && t < len_out) {
            if (y_f16) yh[ybase + co * y_ld + t] = f16_sat(val[i]);
            else yf[ybase + co * y_ld + t] = val[i];
          }
          if (++ph == UP) {
            ph = 0;
            ++co;
          }
        }
      }
    }
  }
}

// every other combination (transposed conv scatter, tanh / relu outputs)
__device__ TB200_ROLE_INLINE void epilogue_generic(const ConvArgs& a, uint32_t tm0, int nsub, int slabs, int slab0,
                                                   int slab_step, int nt, int m_base, int len_out, long long ybase,
                                                   long long rbase) {
  const bool plain_out = !a.residual && !a.accumulate && a.out_act == TB200_OUT_NONE;
  for (int sub = 0; sub < nsub; ++sub) {
    const int m = m_base + sub * kTileM;  // output row (regular) / input row (transposed)
    for (int s = slab0; s < slabs; s += slab_step) {
      uint32_t v[16];
      __syncwarp();
      tmem_ld_x16(tm0 + (uint32_t)(sub * a.NT + s * 16), v);
      const int n0 = nt * a.NT + s * 16;
      tmem_ld_wait();
      if (a.up == 0) {
        if (m >= len_out) continue;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int co = n0 + i;
          if (co >= a.N_total) continue;
          const long long yi = ybase + (long long)co * a.y_ld + m;
          store_y(a, yi, finish(__uint_as_float(v[i]), a, co, rbase + (long long)co * a.r_ld + m, yi));
        }
      } else {
        // transposed conv: column n = co*u + phase lands at t = m*u + phase - u/2
        const int t_first = m * a.up - a.up_pad;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = n0 + i;
          const int co = n / a.up;
          const int t = t_first + (n - co * a.up);
          if (n >= a.N_total || t < 0 || t >= len_out) continue;
          const long long yi = ybase + (long long)co * a.y_ld + t;
          if (plain_out) {
            const float val = (__uint_as_float(v[i]) + (a.bias ? __ldg(a.bias + co) : 0.f)) * a.out_alpha;
            if (a.y_f16) reinterpret_cast<__half*>(a.y)[yi] = f16_sat(val);
            else reinterpret_cast<float*>(a.y)[yi] = val;
          } else {
            store_y(a, yi, finish(__uint_as_float(v[i]), a, co, rbase + (long long)co * a.r_ld + t, yi));
          }
        }
      }
    }
  }
}

// SNAKE selects the staging family compiled into the kernel (anti-aliased SnakeBeta vs pointwise activations):
// two smaller kernels allocate registers better than one with every path inlined.
// CTAS = resident CTAs per SM: 2 halves the per-CTA registers (64), shared memory and TMEM columns but doubles the
// independent warps (and scoreboards) that hide global-memory latency -- used for the pointwise staging family.
// Valid input length of utterance b, clamped to the padded extent the caller declared (a length above L_in_max would
// make the staging read and the epilogue write rows behind the row pitch).
__device__ __forceinline__ int tile_len(const ConvArgs& a, int b) {
  return a.len_in ? min(__ldg(a.len_in + b), a.L_in_max) : a.L_in_max;
}

template <typename T, bool SNAKE, int CTAS>
__global__ void __launch_bounds__(Roles<SNAKE>::kThreads, CTAS) conv1d_umma_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int kWorkerWarps = Roles<SNAKE>::kWorkers, kMmaWarp = Roles<SNAKE>::kMma, kLoadWarp = Roles<SNAKE>::kLoad;
  constexpr int E = ElemTraits<T>::kEpc;
  constexpr bool kTf32 = ElemTraits<T>::kTf32;
  constexpr int kStepK = 2 * E;  // K per tcgen05.mma

  extern __shared__ __align__(128) uint8_t smem[];
  const WsLayout lay = ws_layout(a);
  uint8_t* smW = smem + lay.w_off;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* w_empty = w_full + a.ring_slots;
  uint64_t* a_full = w_empty + a.ring_slots;
  uint64_t* a_empty = a_full + a.a_bufs;
  uint64_t* acc_full = a_empty + a.a_bufs;
  uint64_t* acc_empty = acc_full + a.acc_bufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + lay.tmem_off);
  float* scratch = reinterpret_cast<float*>(smem + lay.scratch_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_kernel_start = clock64();

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < a.ring_slots; ++i) {
        mbar_init(w_full + i, 1);
        mbar_init(w_empty + i, 1);
      }
      for (int i = 0; i < a.a_bufs; ++i) {
        mbar_init(a_full + i, 1);
        mbar_init(a_empty + i, 1);
      }
      for (int i = 0; i < a.acc_bufs; ++i) {
        mbar_init(acc_full + i, 1);
        mbar_init(acc_empty + i, kWorkerWarps - a.n_prod);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, a.tmem_cols);
    tmem_relinquish();
  }
  float* bias_s = reinterpret_cast<float*>(smem + lay.bias_off);
  if (a.epi_fast)
    for (int i = threadIdx.x; i < a.N_total; i += blockDim.x) bias_s[i] = a.bias ? __ldg(a.bias + i) * a.out_alpha : 0.f;
  else if (a.epi_up)
    for (int i = threadIdx.x; i < a.N_total; i += blockDim.x) bias_s[i] = a.bias ? __ldg(a.bias + i / a.up) * a.out_alpha : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int rows_tile = a.S * kTileM;
  const int extra_row = a.up > 0 ? 1 : 0;
  const bool restage_per_nt = a.n_panels > 1;   // the A buffer holds one channel panel only
  const int kc_per_panel = a.n_kchunks / a.n_panels;
  const int groups_per_panel = kc_per_panel * (a.KC / E);

  if (warp < a.n_prod) {
    // ================================ producers ================================
    uint32_t ai = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_utt;
      const int t0 = (tile - b * a.tiles_per_utt) * rows_tile;
      const int len = tile_len(a, b);
      if (t0 >= len + extra_row || len <= 0) continue;
      const int t_lo = t0 - a.halo_l;
      // pull this CTA's NEXT tile towards L2 while the current one is staged (the tile schedule is static)
      if (a.l2_prefetch && a.n_panels == 1 && tile + (int)gridDim.x < a.total_tiles) {
        const int ntile = tile + gridDim.x;
        const int nb = ntile / a.tiles_per_utt;
        const int nt0 = (ntile - nb * a.tiles_per_utt) * rows_tile;
        const int nlen = tile_len(a, nb);
        if (nt0 < nlen + extra_row && nlen > 0) {
          const int esz = a.x_f16 ? 2 : 4, per_line = 128 / esz;
          const int lo = max(nt0 - a.halo_l, 0) / per_line * per_line;
          const int hi = min(nt0 - a.halo_l + a.R, nlen);
          const int nlines = (hi - lo + per_line - 1) / per_line;
          const char* base = reinterpret_cast<const char*>(a.x) + (long long)nb * a.x_bs * esz;
          for (int idx = warp * 32 + lane; idx < a.Cin * nlines; idx += a.n_prod * 32) {
            const int c = idx / nlines, l = idx - c * nlines;
            prefetch_l2(base + ((long long)c * a.x_ld + lo + l * per_line) * esz);
          }
        }
      }
      for (int nt = 0; nt < a.n_ntiles; ++nt) {
        if (!restage_per_nt && nt > 0) break;
        for (int pn = 0; pn < a.n_panels; ++pn, ++ai) {
          const int buf = ai % a.a_bufs;
          mbar_wait_relaxed(a_empty + buf, ((ai / a.a_bufs) & 1) ^ 1);
          if (threadIdx.x == 0) trace(a, 0, ai);
          const bool trace_warp = a.trace && blockIdx.x == 0 && ai == 5 && lane == 0;
          if (trace_warp) a.trace[kTraceTiles * 8 + kTraceCtas + 2 * warp] = clock64();
          T* smA = reinterpret_cast<T*>(smem + lay.a_off + buf * a.a_bytes);
          const int g0 = pn * groups_per_panel;
          if constexpr (SNAKE) {
            // lane = channel streaming path (edge segments clamp their reads inside); the generic per-row path only
            // when its alignment / channel-count preconditions fail
            if (a.aa_fast) {
              const int ncb = groups_per_panel * E / 32;
              // row segments per 32-channel block: tasks = ncb * nseg is a multiple of the producer warps (balanced rounds)
              int gcd = ncb, r2 = a.n_prod;
              while (r2) {
                const int tmp = gcd % r2;
                gcd = r2;
                r2 = tmp;
              }
              stage_aa_channel<T>(a, b, t_lo, g0 * E / 32, ncb, a.n_prod / gcd, len, smA, warp, lane);
            } else {
              const UmmaStore<T> st{smA, a.R};
              stage_aa_snake<E, true>(a, b, t_lo, a.R, g0, groups_per_panel, len, st, scratch, warp, a.n_prod, lane);
            }
          } else {
           if (a.pw_vec) {
            if constexpr (E == 8) {
              if (a.x_f16) stage_pointwise_vec<T, true>(a, b, t_lo, g0, groups_per_panel, len, smA, warp, lane);
              else stage_pointwise_vec<T, false>(a, b, t_lo, g0, groups_per_panel, len, smA, warp, lane);
            } else {
              stage_pointwise_vec<T, false>(a, b, t_lo, g0, groups_per_panel, len, smA, warp, lane);
            }
           } else {
            stage_pointwise_mlp<T>(a, b, t_lo, g0, groups_per_panel, len, smA, warp, lane);
           }
          }
          if (trace_warp) a.trace[kTraceTiles * 8 + kTraceCtas + 2 * warp + 1] = clock64();
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, %0;" ::"r"(a.n_prod * 32) : "memory");
          if (threadIdx.x == 0) {
            trace(a, 1, ai);
            mbar_arrive(a_full + buf);
          }
        }
      }
    }
  } else if (warp == kLoadWarp) {
    // ================================ weight loader ================================
    // (whole warp convergent, one elected lane issues: see the MMA issuer)
    {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w);
      if (a.resident) {
        // A CTA whose tiles are all skipped (ragged batch) goes straight to the final barrier and exits: it must not
        // leave bulk copies (and their complete_tx) in flight into shared memory that the SM's next CTA may own.
        bool any = false;
        for (int tile = blockIdx.x; tile < a.total_tiles && !any; tile += gridDim.x) {
          const int b = tile / a.tiles_per_utt;
          const int t0 = (tile - b * a.tiles_per_utt) * rows_tile;
          const int len = tile_len(a, b);
          any = !(t0 >= len + extra_row || len <= 0);
        }
        for (int c = 0; any && c < a.n_chunks; ++c) {
          if (elect_one()) {
            mbar_arrive_expect_tx(w_full + c, a.chunk_bytes);
            bulk_copy_g2s(smW + (long long)c * a.chunk_bytes, wsrc + (long long)c * a.chunk_bytes, a.chunk_bytes, w_full + c);
          }
        }
      } else {
        uint32_t cc = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
          const int b = tile / a.tiles_per_utt;
          const int t0 = (tile - b * a.tiles_per_utt) * rows_tile;
          const int len = __shfl_sync(0xffffffffu, tile_len(a, b), 0);
          if (t0 >= len + extra_row || len <= 0) continue;
          for (int nt = 0; nt < a.n_ntiles; ++nt)
            for (int pn = 0; pn < a.n_panels; ++pn)
              for (int j = 0; j < a.ntaps; ++j)
                for (int kcl = 0; kcl < kc_per_panel; ++kcl, ++cc) {
                  const int c = (nt * a.ntaps + j) * a.n_kchunks + pn * kc_per_panel + kcl;
                  const int slot = cc % a.ring_slots;
                  mbar_wait(w_empty + slot, ((cc / a.ring_slots) & 1) ^ 1);
                  __syncwarp();
                  if (elect_one()) {
                    mbar_arrive_expect_tx(w_full + slot, a.chunk_bytes);
                    bulk_copy_g2s(smW + (long long)slot * a.chunk_bytes, wsrc + (long long)c * a.chunk_bytes, a.chunk_bytes,
                                  w_full + slot);
                  }
                }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer ================================
    // The whole warp runs this loop convergently and lane 0 alone issues tcgen05.mma / commit: every operand is then
    // warp-uniform (uniform registers), where a lane-0-only region made the compiler wrap each MMA in an
    // ELECT / R2UR.BROADCAST waterfall (~18 instructions, ~115 cycles per MMA against 48-64 of the tensor pipe).
    {
      const uint32_t idesc = make_instr_desc(a.NT, kTf32);
      const uint32_t lbo_a = a.R * 16, lbo_b = a.NT * 16;
      const uint32_t desc_hi = smem_desc_hi(128);
      const uint32_t a_kstep = (2u * lbo_a) >> 4, b_kstep = (2u * lbo_b) >> 4;   // descriptor units (16 bytes) per K step
      const uint32_t smW_u = smem_u32(smW);
      uint32_t ai = 0, ac = 0, cc = 0;
      long long w_wait = 0;   // cycles the MMA warp spent waiting for streamed weight chunks (trace only)
      bool first_tile = true;
      uint32_t a_buf_cur = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int b = tile / a.tiles_per_utt;
        const int t0 = (tile - b * a.tiles_per_utt) * rows_tile;
        const int len = __shfl_sync(0xffffffffu, tile_len(a, b), 0);
        const int rows = len + extra_row;
        if (t0 >= rows || len <= 0) continue;
        const int nsub = min(a.S, (rows - t0 + kTileM - 1) / kTileM);
        for (int nt = 0; nt < a.n_ntiles; ++nt, ++ac) {
          const int abuf = ac % a.acc_bufs;
          mbar_wait(acc_empty + abuf, ((ac / a.acc_bufs) & 1) ^ 1);
          __syncwarp();
          if (lane == 0) trace(a, 6, ac);
          const uint32_t acc_col = tmem_base + (uint32_t)(abuf * a.S * a.NT);
          for (int pn = 0; pn < a.n_panels; ++pn) {
            if (restage_per_nt || nt == 0) {
              a_buf_cur = ai % a.a_bufs;
              mbar_wait(a_full + a_buf_cur, (ai / a.a_bufs) & 1);
              __syncwarp();
              if (lane == 0) trace(a, 2, ai);
              ++ai;
            }
            tc_fence_after();
            const uint32_t smA_u = smem_u32(smem + lay.a_off + a_buf_cur * a.a_bytes);
            for (int j = 0; j < a.ntaps; ++j) {
              const uint32_t a_row = smA_u + (uint32_t)(a.tap_off[j] + a.halo_l) * 16u;
              for (int kcl = 0; kcl < kc_per_panel; ++kcl, ++cc) {
                int slot;
                if (a.resident) {
                  slot = (nt * a.ntaps + j) * a.n_kchunks + pn * kc_per_panel + kcl;
                  if (first_tile) mbar_wait(w_full + slot, 0);
                  __syncwarp();
                } else {
                  slot = cc % a.ring_slots;
                  const long long tw = a.trace ? clock64() : 0;
                  mbar_wait(w_full + slot, (cc / a.ring_slots) & 1);
                  if (a.trace) w_wait += clock64() - tw;
                  __syncwarp();
                }
                tc_fence_after();
                const uint32_t b_base = smW_u + (uint32_t)slot * (uint32_t)a.chunk_bytes;
                const uint32_t a_base = a_row + (uint32_t)(kcl * (a.KC / E)) * lbo_a;
                const uint32_t fresh = (pn == 0 && j == 0 && kcl == 0) ? 0u : 1u;
                // descriptors advance by 32-bit adds on their start-address field (the single issuing thread is
                // instruction bound: ~30 dependent integer ops per MMA measured 160-290 cycles per MMA)
                const uint32_t a_lo0 = smem_desc_lo(a_base, lbo_a), b_lo0 = smem_desc_lo(b_base, lbo_b);
                const int nks = a.KC / kStepK;
                if (elect_one()) {   // one leader issues every MMA of this weight chunk
                  for (int sub = 0; sub < nsub; ++sub) {
                    uint32_t a_lo = a_lo0 + (uint32_t)(sub * kTileM), b_lo = b_lo0;   // 16 bytes per row -> +1 per row
                    const uint32_t d_col = acc_col + (uint32_t)(sub * a.NT);
                    umma_ss_lohi<kTf32>(d_col, a_lo, desc_hi, b_lo, desc_hi, idesc, fresh);
                    for (int ks = 1; ks < nks; ++ks) {
                      a_lo += a_kstep;
                      b_lo += b_kstep;
                      umma_ss_lohi<kTf32>(d_col, a_lo, desc_hi, b_lo, desc_hi, idesc, 1u);
                    }
                  }
                  if (!a.resident) umma_commit(w_empty + slot);
                }
                __syncwarp();
              }
            }
            if ((restage_per_nt || nt == a.n_ntiles - 1) && elect_one()) umma_commit(a_empty + a_buf_cur);
          }
          if (elect_one()) umma_commit(acc_full + abuf);
          if (lane == 0) trace(a, 3, ac);
          if (lane == 0 && a.trace && blockIdx.x == 0 && ac < (uint32_t)kTraceTiles) a.trace[ac * 8 + 7] = w_wait;   // cumulative
        }
        first_tile = false;
      }
    }
  } else {
    // ================================ epilogue ================================
    const int n_epi = kWorkerWarps - a.n_prod;
    const int ew = warp - a.n_prod;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int slab_step = n_epi >> 2;             // 1 or 2 warps per quarter split the 16-column slabs
    const int slab0 = ew >> 2;
    const int mode = a.epi_fast ? ((a.residual ? 1 : 0) + (a.accumulate ? 1 : 0)) : -1;
    uint32_t ac = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_utt;
      const int t0 = (tile - b * a.tiles_per_utt) * rows_tile;
      const int len = tile_len(a, b);
      const int rows = len + extra_row;
      if (t0 >= rows || len <= 0) continue;
      const int len_out = a.up > 0 ? len * a.up : len;
      const int nsub = min(a.S, (rows - t0 + kTileM - 1) / kTileM);
      const long long ybase = (long long)b * a.y_bs, rbase = (long long)b * a.r_bs;
      // pull the auxiliary rows (residual / old y) of this CTA's NEXT tile towards L2
      if (a.l2_prefetch && mode >= 1 && a.n_ntiles == 1 && tile + (int)gridDim.x < a.total_tiles) {
        const int ntile = tile + gridDim.x;
        const int nb = ntile / a.tiles_per_utt;
        const int nt0 = (ntile - nb * a.tiles_per_utt) * rows_tile;
        const int nlen = tile_len(a, nb);
        if (nt0 < nlen) {
          const int esz = a.residual ? (a.r_f16 ? 2 : 4) : (a.y_f16 ? 2 : 4);
          const char* base = a.residual ? reinterpret_cast<const char*>(a.residual) + (long long)nb * a.r_bs * esz
                                        : reinterpret_cast<const char*>(a.y) + (long long)nb * a.y_bs * esz;
          const int ld = a.residual ? a.r_ld : a.y_ld;
          const int my_ch = ((a.NT / 16 - slab0 + slab_step - 1) / slab_step) * 16;   // channels this warp will drain
          for (int idx = lane; idx < my_ch * a.S; idx += 32) {
            const int ci = idx / a.S, sub = idx - ci * a.S;
            const int ch = (slab0 + (ci >> 4) * slab_step) * 16 + (ci & 15);
            const int row = nt0 + sub * kTileM + q * 32;
            if (row < nlen) prefetch_l2(base + ((long long)ch * ld + row) * esz);
          }
        }
      }
      for (int nt = 0; nt < a.n_ntiles; ++nt, ++ac) {
        const int abuf = ac % a.acc_bufs;
        mbar_wait_relaxed(acc_full + abuf, (ac / a.acc_bufs) & 1);
        if (ew == 0 && lane == 0) trace(a, 4, ac);
        tc_fence_after();
        const int slabs = a.NT / 16;
        const uint32_t tm0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * a.S * a.NT);
        const int m_base = t0 + q * 32 + lane;
        const EpiAux axr{a.residual, (uint32_t)rbase, (uint32_t)a.r_ld, a.r_f16, a.res_beta};
        const EpiAux axy{a.y, (uint32_t)ybase, (uint32_t)a.y_ld, a.y_f16, 1.0f};
        if (mode == 2) epilogue_plain<2, TB200_SNAKE_EPI_BATCH>(a, bias_s, axr, axy, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase);
        else if (mode == 1) epilogue_plain<1, SNAKE ? TB200_SNAKE_EPI_BATCH : 2>(a, bias_s, a.residual ? axr : axy, axy, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase);
        else if (mode == 0) epilogue_plain<0, 2>(a, bias_s, axy, axy, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase);
        else if (a.epi_up) {
          const bool pair = a.epi_up == 2;
          switch (a.up) {
            case 2: epilogue_up<2, false>(a, bias_s, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase); break;
            case 4:
              if (pair) epilogue_up<4, true>(a, bias_s, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase);
              else epilogue_up<4, false>(a, bias_s, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase);
              break;
            case 6: epilogue_up<6, false>(a, bias_s, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase); break;
            default:
              if (pair) epilogue_up<8, true>(a, bias_s, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase);
              else epilogue_up<8, false>(a, bias_s, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, (uint32_t)ybase);
              break;
          }
        } else epilogue_generic(a, tm0, nsub, slabs, slab0, slab_step, nt, m_base, len_out, ybase, rbase);
        tc_fence_before();
        __syncwarp();
        if (ew == 0 && lane == 0) trace(a, 5, ac);
        if (lane == 0) mbar_arrive(acc_empty + abuf);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (a.trace && threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    a.trace[kTraceTiles * 8 + blockIdx.x] = ((long long)smid << 48) | (clock64() - t_kernel_start);
  }
  if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
static long long* g_trace = nullptr;   // TB200_TRACE=1 only (debugging aid; the one allocation the library ever makes)

// Per-device state (a process may drive several GPUs): SM count, shared-memory limit, and which kernel instantiations
// have had their dynamic shared-memory limit raised on that device.
constexpr int kMaxDevices = 64;
struct DeviceState {
  int sm_count = 0, max_smem = 0;
  bool configured[4] = {false, false, false, false};
};
static DeviceState g_dev[kMaxDevices];
static thread_local DeviceState* t_dev = nullptr;   // device of the call in progress
#define g_sm_count (t_dev->sm_count)
#define g_max_smem (t_dev->max_smem)

// Tuning / debugging knobs: read from the environment ONCE per process (not per launch).
struct Knobs {
  bool init = false;
  bool snake_plan_old = false, trace = false, l2_prefetch = false, plan_debug = false;
  double snake_a2_ratio = 0.3;
  int max_s = 0, nprod_snake = 0, nprod_pw = 0;
};
static Knobs g_knobs;
static void read_knobs() {
  if (g_knobs.init) return;
  g_knobs.snake_plan_old = getenv("TB200_SNAKE_PLAN_OLD") != nullptr;
  if (const char* e = getenv("TB200_SNAKE_A2_RATIO")) g_knobs.snake_a2_ratio = atof(e);
  g_knobs.trace = getenv("TB200_TRACE") != nullptr;
  if (const char* e = getenv("TB200_L2_PREFETCH")) g_knobs.l2_prefetch = atoi(e) != 0;
  if (const char* e = getenv("TB200_MAX_S")) g_knobs.max_s = atoi(e);
  if (const char* e = getenv("TB200_NPROD_SNAKE")) g_knobs.nprod_snake = atoi(e);
  if (const char* e = getenv("TB200_NPROD_PW")) g_knobs.nprod_pw = atoi(e);
  g_knobs.plan_debug = getenv("TB200_PLAN_DEBUG") != nullptr;
  g_knobs.init = true;
}

static int device_props() {
  int dev = 0;
  TB200_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(TB200_E_BADARG, "conv1d: device index %d", dev);
  t_dev = &g_dev[dev];
  if (!t_dev->sm_count) {
    TB200_CUDA_CHECK(cudaDeviceGetAttribute(&t_dev->sm_count, cudaDevAttrMultiProcessorCount, dev));
    TB200_CUDA_CHECK(cudaDeviceGetAttribute(&t_dev->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  }
  read_knobs();
  return 0;
}

int fill_conv_args(const tb200_conv1d_params* p, int precision, ConvArgs& a);  // api.cu

template <typename T, bool SNAKE, int CTAS>
static int launch_t(const ConvArgs& a, int smem_bytes, cudaStream_t stream) {
  auto kern = conv1d_umma_kernel<T, SNAKE, CTAS>;
  bool& configured = t_dev->configured[(sizeof(T) == 2 ? 0 : 2) + (SNAKE ? 1 : 0)];
  if (!configured) {
    TB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem / CTAS));
    configured = true;
  }
  const int slots = g_sm_count * CTAS;
  int grid = slots < a.total_tiles ? slots : a.total_tiles;  // persistent: CTAS CTAs per SM
  if (grid < 1) grid = 1;
  kern<<<grid, Roles<SNAKE>::kThreads, smem_bytes, stream>>>(a);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Choose sub-tiles per CTA tile (S), buffer counts and the channel-panel split so that everything fits
// in shared memory / the 512 TMEM columns.
static int plan(ConvArgs& a, int elem_bytes, int rows_max, int ctas, bool tall_first) {
  const int smem_cap = ctas == 1 ? g_max_smem : (g_max_smem + 1024) / ctas - 1024 - 512;  // 1 KB reserved per resident CTA
  const int tmem_cap = 512 / ctas;
  const int span = a.R - kTileM;  // halo rows (left + right)
  const int epc = 16 / elem_bytes;
  const int fixed = (2 * kMaxRing + 8) * 8 + 16 + kMaxProdWarps * 2 * kAaScratch * 4 + kBiasCache * 4 + 256;
  const long long w_total = (long long)a.n_chunks * a.chunk_bytes;
  // pass 0 insists on weights resident in shared memory and all channels staged at once (first fit, largest tile).
  // pass 1 (streamed weights and / or channel panels) scores every candidate: most rows per weight pass first (every
  // extra sub-tile divides the weight bytes re-read from L2 per output row), then double-buffered staging (measured:
  // a double-buffered panel that is re-staged per N tile beats a single-buffered full tile on the acoustic model's
  // GEMMs), then the fewest staging passes.
  long long best_key = -1;
  ConvArgs best = a;
  for (int pass = 0; pass < 2; ++pass) {
    // Candidate order of pass 0 (first fit wins).  Snake prologue: tile sizes outermost -- a taller tile with a single
    // staging buffer beats a shorter double-buffered one (per-segment warm-up of the streaming filter, MMAs a small
    // part of a tile; measured -8 % on the C = 128 layers).  Pointwise prologues: double buffering outermost
    // (measured +2 % the other way round).
    for (int oi = 0; oi < 8; ++oi) {
      {
        const int si = tall_first ? oi / 2 : oi % 4, bi = tall_first ? oi % 2 : oi / 4;
        const int S = TB200_MAX_S >> si, a_bufs = 2 - bi;
        if (S < 1) continue;
        if (S > 1 && ((S / 2) * kTileM >= rows_max)) continue;  // tile longer than the data
        // keep the SMs busy: at least two tiles per SM with resident weights; four with streamed weights (single-
        // buffered accumulators lose the MMA / epilogue overlap, so a ragged last wave costs more: measured on the
        // acoustic model's GEMMs, where 1.3 waves of double-height tiles were 8 % slower than 2.2 waves of single ones)
        if (S > 1 && (long long)a.B * ((rows_max + S * kTileM - 1) / (S * kTileM)) < (pass == 0 ? 2 : 4) * g_sm_count * ctas) continue;
        if (S > 1 && pass == 0 && 2 * S * a.NT > tmem_cap) continue;   // resident weights: insist on two accumulator buffers
        if (S * a.NT > tmem_cap) continue;
        const int acc_bufs = 2 * S * a.NT <= tmem_cap ? 2 : 1;
        const int R = S * kTileM + span;
        for (int n_panels = 1; n_panels <= a.n_kchunks; ++n_panels) {
          if (a.n_kchunks % n_panels) continue;
          const int a_bytes = (a.Cin_pad / n_panels / epc) * R * 16;
          const long long budget = (long long)smem_cap - fixed - (long long)a_bufs * a_bytes;
          if (budget < 2LL * a.chunk_bytes) continue;
          const bool resident = w_total <= budget && a.n_chunks <= kMaxRing;
          if (pass == 0 && (!resident || n_panels > 1)) continue;
          ConvArgs c = a;
          c.S = S; c.a_bufs = a_bufs; c.acc_bufs = acc_bufs; c.n_panels = n_panels; c.R = R; c.a_bytes = a_bytes;
          if (resident) {
            c.resident = 1;
            c.ring_slots = c.n_chunks;
          } else {
            c.resident = 0;
            long long slots = budget / c.chunk_bytes;
            c.ring_slots = (int)(slots > 8 ? 8 : slots);
          }
          int cols = 32;
          while (cols < acc_bufs * S * c.NT) cols <<= 1;
          c.tmem_cols = cols;
          c.tiles_per_utt = (rows_max + S * kTileM - 1) / (S * kTileM);
          c.total_tiles = c.tiles_per_utt * c.B;
          if (pass == 0) {
            // Snake prologue, tall tile with ONE staging buffer (weights resident): staging and MMA issue alternate.
            // When the MMAs are a large part of a tile (many taps), stream the weights through a small ring instead
            // and double-buffer the staging -- estimated from the in-kernel traces (profiles/r1_trace_*.txt):
            //   staging ~ tasks per producer warp x (rows per segment + 13 warm-up rows) x 310 cycles
            //   MMA     ~ S x taps x K-steps x max(70, N) cycles (one issuing thread)
            // Measured on the k = 11, C = 64 layers: 29 K + 20 K cycles in series -> 2.20 ms; overlapped -> 1.95 ms
            // (MMA issue has since dropped to 12 K cycles; the rule fires from 30 % of the staging time).
            // (A full cost model over every (S, buffers, panels) was tried and lost on the C >= 128 layers, whose
            // weight re-streaming per tile it underestimates.)
            if (tall_first && a_bufs == 1 && S > 1 && !g_knobs.snake_plan_old) {
              const int n_prod = Roles<true>::kWorkers - 8;
              const int ncb = a.Cin_pad / 32;
              int g = ncb > 0 ? ncb : 1, r2 = n_prod;
              while (r2) { const int t = g % r2; g = r2; r2 = t; }
              const int nseg = n_prod / g;
              const double stage = (double)((ncb * nseg + n_prod - 1) / n_prod) * ((double)R / nseg + 13.0) * 310.0;
              const double mma = (double)S * a.ntaps * (a.Cin_pad / 16) * (a.NT > 70 ? a.NT : 70);
              const long long budget2 = (long long)smem_cap - fixed - 2LL * a_bytes;
              const long long slots = budget2 > 0 ? budget2 / c.chunk_bytes : 0;
              const double ratio = g_knobs.snake_a2_ratio;   // tuning knob
              if (mma >= ratio * stage && slots >= 4) {
                c.a_bufs = 2;
                c.resident = 0;
                c.ring_slots = (int)(slots > 8 ? 8 : slots);
              }
            }
            a = c;
            return 0;
          }
          const long long stagings = n_panels > 1 ? a.n_ntiles : 1;
          const long long key = (((TB200_MAX_S - S) * 4 + (2 - a_bufs)) * 64 + stagings) * 4096 + n_panels;   // smaller is better
          if (best_key < 0 || key < best_key) {
            best_key = key;
            best = c;
          }
          break;  // more panels of the same (a_bufs, S) are never better
        }
      }
    }
  }
  if (best_key >= 0) {
    a = best;
    return 0;
  }
  return fail(TB200_E_NOSMEM, "conv1d: no tiling of Cin=%d Cout=%d taps=%d fits in shared memory", a.Cin, a.Cout, a.ntaps);
}

int conv1d_umma(const tb200_conv1d_params* p, cudaStream_t stream) {
  int rc = device_props();
  if (rc) return rc;
  ConvArgs a;
  rc = fill_conv_args(p, p->precision, a);
  if (rc) return rc;
  const int elem_bytes = p->precision == TB200_PREC_F16 ? 2 : 4;
  const int rows_max = p->L_in_max + (a.up > 0 ? 1 : 0);
  // One persistent CTA per SM.  (Two resident CTAs per SM -- 64 registers, half the shared memory and TMEM each -- were
  // measured 1.6-2.7x slower: the register cap spills, and with the L1 carved out for shared memory a spill is a DRAM
  // round trip.  plan() and the kernel template keep the CTAS parameter.)
  const bool snake = a.act == TB200_ACT_AA_SNAKEBETA;
  constexpr int ctas = 1;
  rc = plan(a, elem_bytes, rows_max, ctas, snake);
  if (rc) return rc;
  // lane=channel snake staging: 16-byte loads need an aligned base / pitch and 32-channel blocks
  const int align = a.x_f16 ? 8 : 4;
  a.aa_fast = (a.act == TB200_ACT_AA_SNAKEBETA) && (a.Cin % 32 == 0) && ((a.Cin_pad / a.n_panels) % 32 == 0) &&
              (a.x_ld % align == 0) && (a.x_bs % align == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
  // vectorised pointwise staging: 16-byte loads along time
  a.pw_vec = (a.act != TB200_ACT_AA_SNAKEBETA) && (a.x_ld % align == 0) && (a.x_bs % align == 0) &&
             ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) && !(a.x_f16 && p->precision != TB200_PREC_F16) &&
             !(a.act == TB200_ACT_LEAKY_RELU && (a.slope < 0.f || a.slope > 1.f));
  // plain epilogue preconditions (see epilogue_plain)
  {
    const long long lim = 1LL << 31;
    const long long y_ext = (long long)a.B * a.y_bs, r_ext = a.residual ? (long long)a.B * a.r_bs : 0;
    a.epi_fast = a.up == 0 && (a.out_act == TB200_OUT_NONE || (a.out_act == TB200_OUT_RELU && a.out_alpha >= 0.f)) &&
                 a.N_total % 16 == 0 && a.N_total <= kBiasCache &&
                 y_ext < lim && r_ext < lim && a.y_bs >= 0 && a.r_bs >= 0;
  }
  // plain transposed epilogue preconditions (see epilogue_up)
  {
    const long long lim = 1LL << 31;
    const bool ok = (a.up == 2 || a.up == 4 || a.up == 6 || a.up == 8) && !a.residual && !a.accumulate &&
                    a.out_act == TB200_OUT_NONE && a.N_total % 16 == 0 && a.N_total <= kBiasCache &&
                    (long long)a.B * a.y_bs < lim && a.y_bs >= 0;
    const int pair_align = a.y_f16 ? 4 : 8;
    const bool pair = a.y_ld % 2 == 0 && a.y_bs % 2 == 0 && (reinterpret_cast<uintptr_t>(a.y) % pair_align) == 0;
    a.epi_up = ok ? (pair ? 2 : 1) : 0;
  }
  // warp split: the anti-aliased snake staging is the SIMT-heavy side (10 producers + 4 epilogue warps),
  // otherwise the epilogue is (6 + 8)
  // snake: 14 producers + 8 epilogue warps (18 + 4 measured slower even for store-only epilogues); pointwise: 6 + 8
  // pointwise: 6 + 8 when the epilogue fetches auxiliary rows, 10 + 4 when it only stores (staging is then the long
  // pole: measured 30 K vs 8 K cycles per tile)
  a.n_prod = snake ? Roles<true>::kWorkers - 8
                   : Roles<false>::kWorkers - ((a.residual || a.accumulate || !(a.epi_fast || a.epi_up)) ? 8 : 4);
  // The acoustic model's GEMMs (tf32 operands, or 1 x 1 convolutions over fp32 activations): 128-row tiles of 192-1536
  // input channels staged through registers -- every producer task is one global-memory latency, the tile time is
  // the staging time (traces: 33 K cycles per 128 x 128 panel with 6 producers, MMA issue 6 K, drain 11-17 K), so 10 + 4
  // wins even with a residual in the epilogue: 1536->192 1.50 -> 0.97 ms, 384->1536 1.05 -> 0.68, 192->1536 0.92 -> 0.65,
  // 192->192 0.134 -> 0.099 (128 x 866 / 433 positions; profiles/r2_nprod_pw.txt).
  if (!snake && a.up == 0 && !a.x_f16 && (p->precision == TB200_PREC_TF32 || a.ntaps == 1)) a.n_prod = Roles<false>::kWorkers - 4;
  a.trace = nullptr;
  if (g_knobs.trace) {
    if (!g_trace) TB200_CUDA_CHECK(cudaMalloc(&g_trace, kTraceLen * sizeof(long long)));
    TB200_CUDA_CHECK(cudaMemsetAsync(g_trace, 0, kTraceLen * sizeof(long long), stream));
    a.trace = g_trace;
  }
  a.l2_prefetch = 0;  // measured: next-tile L2 prefetch doubles DRAM reads (lines evicted before use) -- kept as a knob
  a.l2_prefetch = g_knobs.l2_prefetch;
  if (g_knobs.max_s >= 1) {  // tuning knob: cap the sub-tiles per CTA tile
    const int cap = g_knobs.max_s;
    if (cap >= 1 && a.S > cap) {
      ConvArgs t = a;
      // re-plan with a smaller S by shrinking rows_max-independent fields
      const int span = a.R - a.S * kTileM;
      t.S = cap; t.R = cap * kTileM + span;
      t.a_bytes = a.a_bytes / a.R * t.R;
      t.tiles_per_utt = (rows_max + cap * kTileM - 1) / (cap * kTileM);
      t.total_tiles = t.tiles_per_utt * a.B;
      t.acc_bufs = 2 * cap * a.NT <= 512 ? 2 : 1;
      int cols = 32;
      while (cols < t.acc_bufs * cap * a.NT) cols <<= 1;
      t.tmem_cols = cols;
      a = t;
    }
  }
  if (const int v = (a.act == TB200_ACT_AA_SNAKEBETA ? g_knobs.nprod_snake : g_knobs.nprod_pw)) {  // tuning knob
    const int w = a.act == TB200_ACT_AA_SNAKEBETA ? Roles<true>::kWorkers : Roles<false>::kWorkers;
    if (v >= 2 && v <= kMaxProdWarps && (w - v == 4 || w - v == 8)) a.n_prod = v;
  }
  const int smem_bytes = ws_layout(a).total;
  if (g_knobs.plan_debug)
    fprintf(stderr, "tb200 plan: Cin=%d Cout=%d taps=%d up=%d act=%d L=%d -> S=%d a_bufs=%d acc_bufs=%d panels=%d ntiles=%d %s ring=%d n_prod=%d smem=%d\n",
            a.Cin, a.Cout, a.ntaps, a.up, a.act, p->L_in_max, a.S, a.a_bufs, a.acc_bufs, a.n_panels, a.n_ntiles,
            a.resident ? "resident" : "streamed", a.ring_slots, a.n_prod, smem_bytes);
  if (smem_bytes > g_max_smem / ctas) return fail(TB200_E_NOSMEM, "conv1d: %d bytes of shared memory", smem_bytes);
  if (p->precision == TB200_PREC_F16)
    return snake ? launch_t<__half, true, ctas>(a, smem_bytes, stream) : launch_t<__half, false, ctas>(a, smem_bytes, stream);
  return snake ? launch_t<float, true, ctas>(a, smem_bytes, stream) : launch_t<float, false, ctas>(a, smem_bytes, stream);
}

int conv_trace_read(long long* host_out, int n) {
  if (!g_trace) return fail(TB200_E_BADARG, "trace: TB200_TRACE was not set");
  TB200_CUDA_CHECK(cudaMemcpy(host_out, g_trace, sizeof(long long) * (n < kTraceLen ? n : kTraceLen), cudaMemcpyDeviceToHost));
  return 0;
}

}  // namespace tb200
