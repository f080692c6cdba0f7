// attention_umma.cu -- relative-position multi-head attention on the sm_100a tensor cores (flash style).
//
// Reference: RelPositionMultiHeadedAttention.forward, Layers/Attention.py:159-198, with rel_shift (:138-157) and
// forward_attention (:66-92):
//     score(i, j) = ((q_i + u_h) . k_j + (q_i + v_h) . p_{i-j}) / sqrt(dk),   j < len;   out_i = sum_j softmax_j v_j
// p_r = linear_pos(PE(r)) comes pre-projected as a (D, pos_cols) table whose column (pos_center - r) is relative
// position r (rel_shift and the table order of RelPositionalEncoding collapse to r = i - j; see acoustic.cu).
//
// One CTA = (utterance, head, 128 query rows).  Per 128-key tile, all three contractions run as tcgen05.mma
// (kind::f16, fp32 accumulation in TMEM; operands are fp16 = tf32's 10-bit mantissa):
//     S  = (Q + u) K^T                     128 x 128, K = dk                       -> TMEM
//     G  = (Q + v) Pband^T                 128 x 256, Pband = the 256 relative positions i0-j0-127 .. i0-j0+128
//     PV = softmax tile x V                128 x dk,  K = 128 keys
// The relative-position term of score (i, j) is G[i][i - j + 127]: a row-dependent window of G.  A thread owns one
// query row (= one TMEM lane); it copies its window of G into its own row of a shared-memory scratch (lane-dependent
// store address, uniform register index) and reads it back in key order.  Online softmax in registers: the thread
// keeps its row's running max, sum and the dk output values; PV of each tile is read back from TMEM and accumulated
// with the usual exp(m_old - m_new) rescale -- no TMEM read-modify-write.
// Warps: 0-3 softmax (TMEM lane quarters), 4 MMA issuer (also requests the positional band: it is pre-packed per
// layer as fp16 operand rows [head][d/8][relative position][8], so a tile's band is one contiguous bulk copy per
// 16-byte group -- cp.async.bulk, no conversion), 5-11 loaders (global fp32 Q/K/V -> fp16 operand tiles, several
// tasks' loads in flight per thread; K/V/band double buffered).
#include <cstdio>
#include <cstdlib>

#include "conv_common.cuh"

namespace tb200 {

constexpr int kAtM = 128;          // query rows per CTA
constexpr int kAtN = 128;          // keys per tile
constexpr int kAtBand = 256;       // relative positions per tile
constexpr int kAtScrPitch = kAtN * 2 + 16;   // bytes per scratch row (fp16 window + 16: conflict-free 16-byte reads)
constexpr int kAtLoaders = 7;      // loader warps
constexpr int kAtThreads = (4 + 1 + kAtLoaders) * 32;

template <int DK>
struct AtLayout {
  static constexpr int kPlanes = DK / 8;                     // 16-byte K groups of the d dimension
  static constexpr int q_bytes = kPlanes * kAtM * 16;        // one Q tile ([d/8][128][8] halves)
  static constexpr int k_bytes = kPlanes * kAtN * 16;
  static constexpr int pb_bytes = kPlanes * kAtBand * 16;
  static constexpr int v_bytes = (kAtN / 8) * DK * 16;       // [keys/8][DK][8]
  static constexpr int p_bytes = (kAtN / 8) * kAtM * 16;     // probabilities [keys/8][128][8]
  static constexpr int scr_bytes = kAtM * kAtScrPitch;
  static constexpr int qu_off = 0, qv_off = q_bytes;
  static constexpr int k_off = 2 * q_bytes;                  // x2 buffers
  static constexpr int pb_off = k_off + 2 * k_bytes;
  static constexpr int v_off = pb_off + 2 * pb_bytes;
  static constexpr int p_off = v_off + 2 * v_bytes;
  static constexpr int scr_off = p_off + p_bytes;
  static constexpr int bar_off = scr_off + scr_bytes;
  static constexpr int total = bar_off + 16 * 8 + 16;
};

enum { AQ_FULL = 0, AKV_FULL0, AKV_FULL1, AKV_EMPTY0, AKV_EMPTY1, ASG_FULL, ASG_EMPTY, AP_FULL, APV_FULL, APV_EMPTY, ANUM };

__device__ __forceinline__ void at_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __noinline__ void at_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    if (mbar_try_wait(bar, parity)) return;
  }
  mbar_timeout();
}

template <int DK>
__global__ void __launch_bounds__(kAtThreads, 1)
relpos_attention_umma_kernel(const float* __restrict__ qkv, long long qkv_bs, int qkv_ld, const __half* __restrict__ pos16, int pos_rows,
                             int pos_center, const float* __restrict__ bias_u, const float* __restrict__ bias_v,
                             const int* __restrict__ len_ptr, int H, int L_max, float* __restrict__ out, long long out_bs, int out_ld) {
  using Lay = AtLayout<DK>;
  constexpr int kPlanes = Lay::kPlanes;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Lay::bar_off);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Lay::bar_off + 16 * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * kAtM;
  const int len = len_ptr ? min(__ldg(len_ptr + b), L_max) : L_max;
  if (i0 >= len) return;                               // whole CTA (uniform)
  const int D = H * DK;
  const int nkt = (len + kAtN - 1) / kAtN;
  const float* qb = qkv + (long long)b * qkv_bs + (long long)(h * DK) * qkv_ld;
  const float* kb = qb + (long long)D * qkv_ld;
  const float* vb = kb + (long long)D * qkv_ld;
  const __half* pb16 = pos16 + (long long)h * kPlanes * pos_rows * 8;   // [d/8][pos_rows][8]

  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bars + AQ_FULL, kAtLoaders);
      mbar_init(bars + AKV_FULL0, kAtLoaders + 1);   // + the band's bulk copies (expect_tx arrival of the MMA warp)
      mbar_init(bars + AKV_FULL1, kAtLoaders + 1);
      mbar_init(bars + AKV_EMPTY0, 1);
      mbar_init(bars + AKV_EMPTY1, 1);
      mbar_init(bars + ASG_FULL, 1);
      mbar_init(bars + ASG_EMPTY, 4);
      mbar_init(bars + AP_FULL, 4);
      mbar_init(bars + APV_FULL, 1);
      mbar_init(bars + APV_EMPTY, 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_g = tmem_base + kAtN, tm_pv = tmem_base + kAtN + kAtBand;

  if (warp < 4) {
    // ================================ softmax: thread = query row ================================
    const int i = warp * 32 + lane;                    // row of the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float sc = rsqrtf((float)DK) * 1.4426950408889634f;   // scores in log2 units
    float m_run = -1e30f, l_run = 0.f, o[DK];
#pragma unroll
    for (int d = 0; d < DK; ++d) o[d] = 0.f;
    uint8_t* scr_row = smem + Lay::scr_off + i * kAtScrPitch;
    __half* scr_h = reinterpret_cast<__half*>(scr_row);
    uint8_t* p_row = smem + Lay::p_off + i * 16;       // + plane * (128 * 16)
    for (int kt = 0; kt < nkt; ++kt) {
      const int j0 = kt * kAtN;
      at_wait(bars + ASG_FULL, kt & 1);
      tc_fence_after();
      // the window of G this row needs, in key order, into its scratch row: bd[j] = G[i][i + 127 - j]
      {
        const int c_lo = warp * 32, c_hi = warp * 32 + 31 + 127;       // columns touched by this warp's rows
#pragma unroll 1
        for (int c0 = (c_lo / 16) * 16; c0 <= c_hi; c0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(tm_g + lane_base + (uint32_t)c0, v);
          tmem_ld_wait();
          const int jb = i + 127 - c0;                 // key index of v[0]; v[k] -> key jb - k
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int j = jb - k;
            if (j >= 0 && j < kAtN) scr_h[j] = __float2half_rn(__uint_as_float(v[k]));
          }
        }
      }
      // pass A: row maximum of this tile
      float tmax = -1e30f;
#pragma unroll 1
      for (int c = 0; c < kAtN / 16; ++c) {
        uint32_t v[16];
        tmem_ld_x16(tm_s + lane_base + (uint32_t)(c * 16), v);
        const uint4 b0 = *reinterpret_cast<const uint4*>(scr_row + c * 32), b1 = *reinterpret_cast<const uint4*>(scr_row + c * 32 + 16);
        const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const __half2 hp = *reinterpret_cast<const __half2*>(&bw[k >> 1]);
          const float bd = (k & 1) ? __high2float(hp) : __low2float(hp);
          const float s = (__uint_as_float(v[k]) + bd) * sc;
          if (j0 + c * 16 + k < len) tmax = fmaxf(tmax, s);
        }
      }
      const float m_new = fmaxf(m_run, tmax);
      const float alpha = exp2f(m_run - m_new);
      // pass B: probabilities (fp16 operand tile of PV) and the row sum
      float psum = 0.f;
#pragma unroll 1
      for (int c = 0; c < kAtN / 16; ++c) {
        uint32_t v[16];
        tmem_ld_x16(tm_s + lane_base + (uint32_t)(c * 16), v);
        const uint4 b0 = *reinterpret_cast<const uint4*>(scr_row + c * 32), b1 = *reinterpret_cast<const uint4*>(scr_row + c * 32 + 16);
        const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          const __half2 hp = *reinterpret_cast<const __half2*>(&bw[k >> 1]);
          const float s0 = (__uint_as_float(v[k]) + __low2float(hp)) * sc, s1 = (__uint_as_float(v[k + 1]) + __high2float(hp)) * sc;
          const float p0 = (j0 + c * 16 + k < len) ? exp2f(s0 - m_new) : 0.f;
          const float p1 = (j0 + c * 16 + k + 1 < len) ? exp2f(s1 - m_new) : 0.f;
          const __half2 ph = __floats2half2_rn(p0, p1);
          // the sum runs over the ROUNDED probabilities, the values the tensor core multiplies with V
          psum += __low2float(ph) + __high2float(ph);
          pk[k >> 1] = *reinterpret_cast<const uint32_t*>(&ph);
        }
        *reinterpret_cast<uint4*>(p_row + (2 * c) * (kAtM * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(p_row + (2 * c + 1) * (kAtM * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      l_run = l_run * alpha + psum;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        at_arrive(bars + ASG_EMPTY);
        at_arrive(bars + AP_FULL);
      }
      // PV of this tile: o = o * alpha + P V
      at_wait(bars + APV_FULL, kt & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DK / 16; ++c) {
        uint32_t v[16];
        tmem_ld_x16(tm_pv + lane_base + (uint32_t)(c * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) o[c * 16 + k] = fmaf(o[c * 16 + k], alpha, __uint_as_float(v[k]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) at_arrive(bars + APV_EMPTY);
    }
    const int qi = i0 + i;
    if (qi < len) {
      const float inv = 1.0f / l_run;
      float* op = out + (long long)b * out_bs + (long long)(h * DK) * out_ld + qi;
#pragma unroll
      for (int d = 0; d < DK; ++d) op[(long long)d * out_ld] = o[d] * inv;
    }
  } else if (warp == 4) {
    // ================================ MMA issuer ================================
    const uint32_t desc_hi = smem_desc_hi(128);
    const uint32_t idesc_s = make_instr_desc(kAtN, false), idesc_g = make_instr_desc(kAtBand, false), idesc_pv = make_instr_desc(DK, false);
    const uint32_t qu = smem_u32(smem + Lay::qu_off), qv = smem_u32(smem + Lay::qv_off), pt = smem_u32(smem + Lay::p_off);
    constexpr uint32_t lbo_q = kAtM * 16, lbo_k = kAtN * 16, lbo_pb = kAtBand * 16, lbo_p = kAtM * 16, lbo_v = DK * 16;
    // positional band of tile kt: relative positions (i0 - j0 - 127) + n, n < 256 -> rows of the packed table
    auto request_band = [&](int kt) {
      const int buf = kt & 1;
      at_wait(bars + AKV_EMPTY0 + buf, ((kt >> 1) & 1) ^ 1);
      __syncwarp();
      if (elect_one()) {
        const int row0 = pos_center + (i0 - kt * kAtN - 127);
        mbar_arrive_expect_tx(bars + AKV_FULL0 + buf, Lay::pb_bytes);
#pragma unroll
        for (int g = 0; g < kPlanes; ++g)
          bulk_copy_g2s(smem + Lay::pb_off + buf * Lay::pb_bytes + g * (kAtBand * 16), pb16 + ((long long)g * pos_rows + row0) * 8,
                        kAtBand * 16, bars + AKV_FULL0 + buf);
      }
      __syncwarp();
    };
    request_band(0);
    if (nkt > 1) request_band(1);
    at_wait(bars + AQ_FULL, 0);
    for (int kt = 0; kt < nkt; ++kt) {
      const int buf = kt & 1;
      at_wait(bars + AKV_FULL0 + buf, (kt >> 1) & 1);
      at_wait(bars + ASG_EMPTY, (kt & 1) ^ 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t ktile = smem_u32(smem + Lay::k_off + buf * Lay::k_bytes), pband = smem_u32(smem + Lay::pb_off + buf * Lay::pb_bytes);
      const uint32_t vtile = smem_u32(smem + Lay::v_off + buf * Lay::v_bytes);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks)
          umma_ss_lohi<false>(tm_s, smem_desc_lo(qu + ks * 2 * lbo_q, lbo_q), desc_hi, smem_desc_lo(ktile + ks * 2 * lbo_k, lbo_k), desc_hi,
                              idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < DK / 16; ++ks)
          umma_ss_lohi<false>(tm_g, smem_desc_lo(qv + ks * 2 * lbo_q, lbo_q), desc_hi, smem_desc_lo(pband + ks * 2 * lbo_pb, lbo_pb), desc_hi,
                              idesc_g, ks > 0 ? 1u : 0u);
        umma_commit(bars + ASG_FULL);
      }
      __syncwarp();
      at_wait(bars + AP_FULL, kt & 1);
      at_wait(bars + APV_EMPTY, (kt & 1) ^ 1);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kAtN / 16; ++ks)
          umma_ss_lohi<false>(tm_pv, smem_desc_lo(pt + ks * 2 * lbo_p, lbo_p), desc_hi, smem_desc_lo(vtile + ks * 2 * lbo_v, lbo_v), desc_hi,
                              idesc_pv, ks > 0 ? 1u : 0u);
        umma_commit(bars + APV_FULL);
        umma_commit(bars + AKV_EMPTY0 + buf);
      }
      __syncwarp();
      if (kt + 2 < nkt) request_band(kt + 2);   // the buffer is free once the commit above lands
    }
  } else {
    // ================================ loaders: global fp32 -> fp16 operand tiles ================================
    const int lt = (warp - 5) * 32 + lane, nlt = kAtLoaders * 32;
    auto pack8 = [](const float (&f)[8]) {
      return make_uint4(f16x2_sat(f[0], f[1]), f16x2_sat(f[2], f[3]), f16x2_sat(f[4], f[5]), f16x2_sat(f[6], f[7]));
    };
    // Q + u and Q + v: [d/8][row][8]
    for (int task = lt; task < kPlanes * kAtM; task += nlt) {
      const int g = task / kAtM, r = task - g * kAtM;
      const int qi = i0 + r;
      float fu[8], fv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float q = qi < len ? __ldg(qb + (long long)(g * 8 + e) * qkv_ld + qi) : 0.f;
        fu[e] = qi < len ? q + __ldg(bias_u + h * DK + g * 8 + e) : 0.f;
        fv[e] = qi < len ? q + __ldg(bias_v + h * DK + g * 8 + e) : 0.f;
      }
      *reinterpret_cast<uint4*>(smem + Lay::qu_off + (g * kAtM + r) * 16) = pack8(fu);
      *reinterpret_cast<uint4*>(smem + Lay::qv_off + (g * kAtM + r) * 16) = pack8(fv);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) at_arrive(bars + AQ_FULL);
    for (int kt = 0; kt < nkt; ++kt) {
      const int buf = kt & 1, j0 = kt * kAtN;
      at_wait(bars + AKV_EMPTY0 + buf, ((kt >> 1) & 1) ^ 1);
      uint8_t* ktile = smem + Lay::k_off + buf * Lay::k_bytes;
      uint8_t* vtile = smem + Lay::v_off + buf * Lay::v_bytes;
      // K: [d/8][key][8] (zero rows behind the utterance: their scores are masked, but 0 x garbage must stay finite)
      // V: [key/8][d][8].  Task t < nK: K group (g, key); else V group (key/8, d).  kU tasks' loads are in flight at once.
      constexpr int nK = kPlanes * kAtN, nV = (kAtN / 8) * DK, kU = 4;
      for (int t0 = lt; t0 < nK + nV; t0 += kU * nlt) {
        float f[kU][8];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int task = t0 + u * nlt;
          if (task < nK) {
            const int g = task / kAtN, r = task - g * kAtN;
            const int kj = j0 + r;
#pragma unroll
            for (int e = 0; e < 8; ++e) f[u][e] = kj < len ? __ldg(kb + (long long)(g * 8 + e) * qkv_ld + kj) : 0.f;
          } else if (task < nK + nV) {
            const int tv = task - nK;
            const int kp = tv / DK, d = tv - kp * DK;
            const int kj = j0 + kp * 8;
            if (kj + 8 <= len) {       // rows are 16-byte aligned (qkv_ld % 4 == 0, kj % 8 == 0)
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(vb + (long long)d * qkv_ld + kj));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(vb + (long long)d * qkv_ld + kj + 4));
              f[u][0] = a0.x; f[u][1] = a0.y; f[u][2] = a0.z; f[u][3] = a0.w; f[u][4] = a1.x; f[u][5] = a1.y; f[u][6] = a1.z; f[u][7] = a1.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[u][e] = kj + e < len ? __ldg(vb + (long long)d * qkv_ld + kj + e) : 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int task = t0 + u * nlt;
          if (task < nK) {
            *reinterpret_cast<uint4*>(ktile + task * 16) = pack8(f[u]);           // (g * 128 + r) == task
          } else if (task < nK + nV) {
            *reinterpret_cast<uint4*>(vtile + (task - nK) * 16) = pack8(f[u]);    // (kp * DK + d) == task - nK
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) at_arrive(bars + AKV_FULL0 + buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int DK>
static int launch_attention_umma(const float* qkv, int64_t qkv_bs, int32_t qkv_ld, const __half* pos16, int32_t pos_rows,
                                 int32_t pos_center, const float* bias_u, const float* bias_v, const int32_t* len,
                                 int32_t B, int32_t H, int32_t L_max, float* out, int64_t out_bs, int32_t out_ld, cudaStream_t s) {
  auto kern = relpos_attention_umma_kernel<DK>;
  static bool configured_per_dev[kMaxDeviceSlots] = {};
  const int slot = current_device_slot();
  if (slot < 0) return fail(TB200_E_NODEVICE, "relpos_attention_tc: no current CUDA device");
  if (!configured_per_dev[slot]) {
    TB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AtLayout<DK>::total));
    configured_per_dev[slot] = true;
  }
  dim3 grid((L_max + kAtM - 1) / kAtM, H, B);
  kern<<<grid, kAtThreads, AtLayout<DK>::total, s>>>(qkv, qkv_bs, qkv_ld, pos16, pos_rows, pos_center, bias_u, bias_v, len, H, L_max,
                                                     out, out_bs, out_ld);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace tb200

using namespace tb200;

extern "C" int tb200_relpos_attention_tc(const float* qkv, int64_t qkv_bs, int32_t qkv_ld, const void* pos16, int32_t pos_rows,
                                         int32_t pos_center, const float* bias_u, const float* bias_v, const int32_t* len,
                                         int32_t B, int32_t H, int32_t dk, int32_t L_max, float* out, int64_t out_bs,
                                         int32_t out_ld, void* stream) {
  if (!qkv || !pos16 || !bias_u || !bias_v || !out) return fail(TB200_E_BADARG, "relpos_attention_tc: null pointer");
  if (B <= 0 || H <= 0 || L_max <= 0 || B > 65535 || H > 65535) return fail(TB200_E_BADARG, "relpos_attention_tc: bad shape");
  // every tile's 256-row band must lie inside the packed table: relative positions -(L_max+254) .. L_max+127
  if (pos_center - (L_max + 254) < 0 || pos_center + (L_max + 127) >= pos_rows)
    return fail(TB200_E_BADARG, "relpos_attention_tc: packed positional table (%d rows, centre %d) too short for L=%d", pos_rows,
                pos_center, L_max);
  if ((reinterpret_cast<uintptr_t>(pos16) & 15) || qkv_ld % 4 || qkv_bs % 4 || (reinterpret_cast<uintptr_t>(qkv) & 15))
    return fail(TB200_E_BADARG, "relpos_attention_tc: qkv rows and the packed table must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __half* p16 = reinterpret_cast<const __half*>(pos16);
  switch (dk) {
    case 32: return launch_attention_umma<32>(qkv, qkv_bs, qkv_ld, p16, pos_rows, pos_center, bias_u, bias_v, len, B, H, L_max, out, out_bs, out_ld, s);
    case 48: return launch_attention_umma<48>(qkv, qkv_bs, qkv_ld, p16, pos_rows, pos_center, bias_u, bias_v, len, B, H, L_max, out, out_bs, out_ld, s);
    case 64: return launch_attention_umma<64>(qkv, qkv_bs, qkv_ld, p16, pos_rows, pos_center, bias_u, bias_v, len, B, H, L_max, out, out_bs, out_ld, s);
    default: return fail(TB200_E_BADARG, "relpos_attention_tc: head size %d not in {32,48,64}", dk);
  }
}
