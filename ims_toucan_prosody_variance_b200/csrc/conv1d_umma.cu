// conv1d_umma.cu -- implicit-GEMM Conv1d / ConvTranspose1d on the sm_100a tensor cores.
//
//   GEMM view (per utterance b):   D[t, n] = sum_{tap j} sum_{ci} A_j[t, ci] * W_j[n, ci]
//     M = time (128 output rows per tile == 128 TMEM lanes), N = output channels (<= 256 per
//     accumulator), K = input channels, one pass over K per tap.
//   A operand: ONE staged tile of ACT(x) for rows [t0 - halo_l, t0 + 128 + halo_r), written by the
//     producer warps in the canonical K-major SWIZZLE_NONE layout with 16 B per row
//     ([Cin/E][R][E] elements).  Because rows are linear in that layout, tap j is the same tile
//     read through a descriptor whose start address is shifted by tap_off[j] rows: no im2col, no
//     per-tap reload.  The activation (LeakyReLU, or BigVGAN's anti-aliased SnakeBeta = 2x
//     kaiser-sinc up, snake, 2x down) is fused into the staging, so it never touches HBM.
//   B operand: weight blocks [KC/E][NT][E] pre-packed at load time, brought in by the TMA engine
//     (cp.async.bulk + mbarrier), resident for the CTA's lifetime when they fit, else streamed
//     through a ring that tcgen05.commit releases.
//   D: fp32 in TMEM; epilogue = tcgen05.ld -> bias / activation / scale / residual / accumulate ->
//     coalesced NCL stores (a warp writes 32 consecutive time steps of one channel).
//
// One persistent CTA loops over (utterance, time-tile) pairs; tiles past an utterance's length
// are skipped, rows past it are staged as zeros (the reference's zero padding at batch 1).
#include <cstdio>
#include <cstdlib>

#include "conv_common.cuh"

namespace tb200 {

constexpr int kComputeWarps = 8;
constexpr int kThreads = (kComputeWarps + 1) * 32;  // + 1 weight-loader warp

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

// A-tile store policy: 16-byte group g, row r -> canonical K-major SWIZZLE_NONE position.
template <typename T>
struct UmmaStore {
  T* base;
  int R;
  __device__ __forceinline__ void operator()(int g, int r, const float (&v)[ElemTraits<T>::kEpc]) const {
    T* dst = base + ((long long)g * R + r) * ElemTraits<T>::kEpc;
    if constexpr (ElemTraits<T>::kEpc == 8) {
      __half2 h0 = __floats2half2_rn(clamp_f16(v[0]), clamp_f16(v[1]));
      __half2 h1 = __floats2half2_rn(clamp_f16(v[2]), clamp_f16(v[3]));
      __half2 h2 = __floats2half2_rn(clamp_f16(v[4]), clamp_f16(v[5]));
      __half2 h3 = __floats2half2_rn(clamp_f16(v[6]), clamp_f16(v[7]));
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2);
      u.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst) = u;
    } else {
      // the tensor core truncates fp32 operands to tf32: round to nearest here instead
      *reinterpret_cast<float4*>(dst) = make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
    }
  }
};

// ---------------------------------------------------------------------------------------------
// epilogue for one 16-column slab held in registers
// ---------------------------------------------------------------------------------------------


// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <typename T, bool kFast>
__global__ void __launch_bounds__(kThreads, 1) conv1d_umma_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int E = ElemTraits<T>::kEpc;
  constexpr bool kTf32 = ElemTraits<T>::kTf32;
  constexpr int kStepK = 2 * E;  // K per tcgen05.mma

  extern __shared__ __align__(128) uint8_t smem[];
  T* smA = reinterpret_cast<T*>(smem);
  uint8_t* smW = smem + a.a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smW + (long long)a.ring_slots * a.chunk_bytes);
  uint64_t* empty_bar = full_bar + a.ring_slots;
  uint64_t* acc_bar = empty_bar + a.ring_slots;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  float* scratch = reinterpret_cast<float*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < a.ring_slots; ++i) {
        mbar_init(full_bar + i, 1);
        mbar_init(empty_bar + i, 1);
      }
      mbar_init(acc_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kComputeWarps) {
    // ======================= weight loader warp (TMA bulk copies) =======================
    if (lane == 0) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.w);
      if (a.resident) {
        for (int c = 0; c < a.n_chunks; ++c) {
          mbar_arrive_expect_tx(full_bar + c, a.chunk_bytes);
          bulk_copy_g2s(smW + (long long)c * a.chunk_bytes, wsrc + (long long)c * a.chunk_bytes, a.chunk_bytes, full_bar + c);
        }
      } else {
        uint32_t cc = 0;
        for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
          const int b = tile / a.tiles_per_utt;
          const int t0 = (tile - b * a.tiles_per_utt) * kTileM;
          const int len = a.len_in ? __ldg(a.len_in + b) : a.L_in_max;
          const int rows = len + (a.up > 0 ? 1 : 0);
          if (t0 >= rows || len <= 0) continue;
          for (int c = 0; c < a.n_chunks; ++c, ++cc) {
            const int slot = cc % a.ring_slots;
            const uint32_t ph = (cc / a.ring_slots) & 1;
            mbar_wait(empty_bar + slot, ph ^ 1);
            mbar_arrive_expect_tx(full_bar + slot, a.chunk_bytes);
            bulk_copy_g2s(smW + (long long)slot * a.chunk_bytes, wsrc + (long long)c * a.chunk_bytes, a.chunk_bytes,
                          full_bar + slot);
          }
        }
      }
    }
  } else {
    // ======================= compute warps: stage A, issue MMA, epilogue =======================
    const uint32_t idesc = make_instr_desc(a.NT, kTf32);
    const uint32_t lbo_a = a.R * 16, lbo_b = a.NT * 16;
    const uint32_t smA_u = smem_u32(smA), smW_u = smem_u32(smW);
    uint32_t cc = 0;        // running weight-block counter (ring position; issuer thread only)
    uint32_t acc_cnt = 0;   // accumulators completed so far (acc_bar phase)
    bool first_tile = true;
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half_id = warp >> 2;     // column half handled by this warp
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_utt;
      const int t0 = (tile - b * a.tiles_per_utt) * kTileM;
      const int len = a.len_in ? __ldg(a.len_in + b) : a.L_in_max;
      const int rows = len + (a.up > 0 ? 1 : 0);
      if (t0 >= rows || len <= 0) continue;
      const int len_out = a.up > 0 ? len * a.up : len;

      // ---- stage the activated input tile ----
      {
        UmmaStore<T> st{smA, a.R};
        if (a.act == TB200_ACT_AA_SNAKEBETA)
          stage_aa_snake<E, kFast>(a, b, t0 - a.halo_l, a.R, 0, a.Cin_pad / E, len, st, scratch, warp, kComputeWarps, lane);
        else
          stage_pointwise<E>(a, b, t0 - a.halo_l, a.R, 0, a.Cin_pad / E, len, st, warp, kComputeWarps, lane);
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(kComputeWarps * 32) : "memory");

      for (int nt = 0; nt < a.n_ntiles; ++nt) {
        // ---- MMA issue: one elected thread ----
        if (warp == 0) {
          if (lane == 0) {
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int j = 0; j < a.ntaps; ++j) {
              const uint32_t a_row = smA_u + (uint32_t)(a.tap_off[j] + a.halo_l) * 16u;
              for (int kc = 0; kc < a.n_kchunks; ++kc) {
                const int c = (nt * a.ntaps + j) * a.n_kchunks + kc;
                int slot;
                if (a.resident) {
                  slot = c;
                  if (first_tile) mbar_wait(full_bar + slot, 0);
                } else {
                  slot = cc % a.ring_slots;
                  mbar_wait(full_bar + slot, (cc / a.ring_slots) & 1);
                }
                tc_fence_after();
                const uint32_t b_base = smW_u + (uint32_t)slot * (uint32_t)a.chunk_bytes;
                const uint32_t a_base = a_row + (uint32_t)(kc * (a.KC / E)) * lbo_a;
                for (int ks = 0; ks < a.KC / kStepK; ++ks) {
                  const uint64_t da = make_smem_desc(a_base + (uint32_t)(2 * ks) * lbo_a, lbo_a, 128);
                  const uint64_t db = make_smem_desc(b_base + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128);
                  umma_ss<kTf32>(tmem_base, da, db, idesc, accumulate);
                  accumulate = 1;
                }
                if (!a.resident) umma_commit(empty_bar + slot);
                ++cc;
              }
            }
            umma_commit(acc_bar);
          }
          __syncwarp();
        }

        // ---- epilogue: TMEM -> registers -> global ----
        mbar_wait(acc_bar, acc_cnt & 1);
        ++acc_cnt;
        tc_fence_after();
        const int r = q * 32 + lane;       // accumulator row == TMEM lane
        const int m = t0 + r;              // output row (regular) / input row (transposed)
        const int slabs = a.NT / 16;
        for (int s = half_id; s < slabs; s += 2) {
          uint32_t v[16];
          __syncwarp();
          tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 16), v);
          tmem_ld_wait();
          const int n0 = nt * a.NT + s * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int n = n0 + i;
            if (n >= a.N_total) break;
            int co, t;
            if (a.up > 0) {
              co = n / a.up;
              t = m * a.up + (n - co * a.up) - a.up_pad;
            } else {
              co = n;
              t = m;
            }
            if (t < 0 || t >= len_out) continue;
            const long long yidx = (long long)b * a.y_bs + (long long)co * a.y_ld + t;
            const long long ridx = (long long)b * a.r_bs + (long long)co * a.r_ld + t;
            store_y(a, yidx, finish(__uint_as_float(v[i]), a, co, ridx, yidx));
          }
        }
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(kComputeWarps * 32) : "memory");
      }
      first_tile = false;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int g_max_smem = 0;

static int device_props() {
  if (g_sm_count) return 0;
  int dev = 0;
  TB200_CUDA_CHECK(cudaGetDevice(&dev));
  TB200_CUDA_CHECK(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
  TB200_CUDA_CHECK(cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return 0;
}

int fill_conv_args(const tb200_conv1d_params* p, int precision, ConvArgs& a);  // api.cu

template <typename T, bool kFast>
static int launch_t(const ConvArgs& a, int smem_bytes, int grid, cudaStream_t stream) {
  auto kern = conv1d_umma_kernel<T, kFast>;
  TB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  kern<<<grid, kThreads, smem_bytes, stream>>>(a);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int conv1d_umma(const tb200_conv1d_params* p, cudaStream_t stream) {
  int rc = device_props();
  if (rc) return rc;
  ConvArgs a;
  rc = fill_conv_args(p, p->precision, a);
  if (rc) return rc;

  const int bar_bytes = (2 * 256 + 2) * 8 + 16;
  const int scratch_bytes = kComputeWarps * 2 * kAaScratch * 4;
  const int budget = g_max_smem - a.a_bytes - bar_bytes - scratch_bytes - 256;
  if (budget < 2 * a.chunk_bytes) return fail(TB200_E_NOSMEM, "conv1d: input tile of %d bytes leaves no room for weight blocks", a.a_bytes);
  if ((long long)a.n_chunks * a.chunk_bytes <= budget && a.n_chunks <= 256) {
    a.resident = 1;
    a.ring_slots = a.n_chunks;
  } else {
    a.resident = 0;
    a.ring_slots = budget / a.chunk_bytes;
    if (a.ring_slots > 6) a.ring_slots = 6;
  }
  const int smem_bytes = a.a_bytes + a.ring_slots * a.chunk_bytes + bar_bytes + scratch_bytes;
  int grid = a.total_tiles < g_sm_count ? a.total_tiles : g_sm_count;
  // small footprints: let two or three CTAs share an SM so one CTA's staging overlaps another's MMA
  int per_sm = 1;
  if (smem_bytes * 2 + 2048 <= g_max_smem && a.tmem_cols * 2 <= 512) per_sm = 2;
  if (smem_bytes * 3 + 3072 <= g_max_smem && a.tmem_cols * 3 <= 512) per_sm = 3;
  if (a.total_tiles > g_sm_count) grid = a.total_tiles < g_sm_count * per_sm ? a.total_tiles : g_sm_count * per_sm;
  if (grid < 1) grid = 1;
  const bool fast = true;
  if (p->precision == TB200_PREC_F16) return fast ? launch_t<__half, true>(a, smem_bytes, grid, stream) : launch_t<__half, false>(a, smem_bytes, grid, stream);
  return launch_t<float, false>(a, smem_bytes, grid, stream);
}

}  // namespace tb200
