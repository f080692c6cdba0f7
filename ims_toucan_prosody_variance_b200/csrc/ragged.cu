// ragged.cu -- K7 / K12: duration rounding + prosody edits + prefix sum (one warp per utterance),
// pitch/energy edits + variance scaling, and the length-regulator gather/expand fused with the
// pitch/energy embedding add.  Integer results are bit-exact with the reference given the same
// fp32 log-durations; the expand is a pure HBM-bound gather: 4*C*(T+F) + 8*T bytes per utterance.
#include "common.cuh"

namespace tb200 {

constexpr int kFeatDim = 62;          // articulatory vector width
constexpr int kFeatPhoneme = 15;      // Preprocessing/articulatory_features.py:817-901
constexpr int kFeatSilence = 16;
constexpr int kFeatWordBoundary = 21;
constexpr int kFeatVoiced = 61;

// DurationPredictor.py:79 + InferenceToucanTTS.py:219-225 + LengthRegulator.py:52-53.
__global__ void duration_finalize_kernel(const float* __restrict__ log_dur, const long long* __restrict__ gold,
                                         const float* __restrict__ text, const int* __restrict__ text_len, int B,
                                         int T_max, int T_ld, float pause_scale, float dur_scale,
                                         long long* __restrict__ dur_out, int* __restrict__ cum_out,
                                         int* __restrict__ frames_out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int T = min(text_len[b], T_max);
  const float* tx = text + (long long)b * T_max * kFeatDim;
  long long* dout = dur_out + (long long)b * T_ld;
  int* cout = cum_out + (long long)b * T_ld;

  // pass 1: round, edit, store, total
  long long total = 0;
  for (int i = lane; i < T; i += 32) {
    long long d;
    if (gold) {
      d = gold[(long long)b * T_ld + i];
    } else {
      // clamp(round(exp(x) - 1), min=0).long(); exp evaluated in double and rounded once to fp32
      const float e = static_cast<float>(exp(static_cast<double>(log_dur[(long long)b * T_ld + i])));
      const float r = fmaxf(rintf(e - 1.0f), 0.0f);
      d = static_cast<long long>(r);
    }
    const float* row = tx + (long long)i * kFeatDim;
    if (row[kFeatWordBoundary] == 1.0f) d = 0;
    if (row[kFeatSilence] == 1.0f && pause_scale != 1.0f) d = static_cast<long long>(rintf(static_cast<float>(d) * pause_scale));
    if (dur_scale != 1.0f) d = static_cast<long long>(rintf(static_cast<float>(d) * dur_scale));
    dout[i] = d;
    total += d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  const bool rescue = (total == 0);  // every phoneme gets one frame
  __syncwarp();

  // pass 2: inclusive prefix sum, 32 phonemes per step with a running carry
  int carry = 0;
  for (int base = 0; base < T; base += 32) {
    const int i = base + lane;
    int d = 0;
    if (i < T) {
      if (rescue) {
        d = 1;
        dout[i] = 1;
      } else {
        d = static_cast<int>(dout[i]);
      }
    }
    int s = d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    if (i < T) cout[i] = carry + s;
    carry += __shfl_sync(0xffffffffu, s, 31);
  }
  for (int i = T + lane; i < T_ld; i += 32) {  // keep the padding deterministic
    dout[i] = 0;
    cout[i] = carry;
  }
  if (lane == 0) frames_out[b] = carry;
}

// InferenceToucanTTS.py:214-218 (zeroing) and :333-343 (_scale_variance).  One warp per utterance.
__global__ void variance_edit_kernel(float* __restrict__ curve, const float* __restrict__ text,
                                     const int* __restrict__ text_len, int B, int T_max, int T_ld, int feat,
                                     float scale) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int T = min(text_len[b], T_max);
  float* cv = curve + (long long)b * T_ld;
  const float* tx = text + (long long)b * T_max * kFeatDim;
  float sum = 0.f, cnt = 0.f;
  for (int i = lane; i < T; i += 32) {
    float v = cv[i];
    if (tx[(long long)i * kFeatDim + feat] == 0.0f) v = 0.0f;
    cv[i] = v;
    if (v != 0.0f) {
      sum += v;
      cnt += 1.f;
    }
  }
  if (scale == 1.0f) return;
  sum = warp_sum(sum);
  cnt = warp_sum(cnt);
  const float avg = sum / cnt;  // empty selection -> NaN, as torch's mean of an empty tensor
  __syncwarp();
  for (int i = lane; i < T; i += 32) {
    float v = (cv[i] - avg) * scale + avg;
    if (v < 0.0f) v = 0.0f;
    cv[i] = v;
  }
}

// LengthRegulator.py:57-61 (repeat_interleave) + utils.py:475-494 (pad) + InferenceToucanTTS.py:230-232.
// grid (frame tiles of 512, B); a thread resolves the phoneme of 4 consecutive frames (one binary search over the
// inclusive prefix sums, then a short forward scan), and streams the channels: 16-byte coalesced stores, L1-resident
// gathers (consecutive frames mostly share a phoneme).
constexpr int kLrPer = 4;
__global__ void __launch_bounds__(128) length_regulate_kernel(
    const float* __restrict__ enc, long long enc_bs, int enc_ld, const float* __restrict__ pitch,
    const float* __restrict__ energy, int pe_ld, const float* __restrict__ wp, const float* __restrict__ bp,
    const float* __restrict__ we, const float* __restrict__ be, const int* __restrict__ cum, int cum_ld,
    const int* __restrict__ text_len, const int* __restrict__ frames, int C, float* __restrict__ out, long long out_bs,
    int out_ld, int* __restrict__ f2p, int f2p_ld) {
  const int b = blockIdx.y;
  const int f0 = (blockIdx.x * 128 + threadIdx.x) * kLrPer;
  const int F = frames[b];
  if (f0 >= F) return;
  const int T = text_len[b];
  const int* cm = cum + (long long)b * cum_ld;
  int idx[kLrPer];
  float pv[kLrPer], ev[kLrPer];
  {
    int lo = 0, hi = T;  // first i with cum[i] > f0
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(cm + mid) <= f0) lo = mid + 1;
      else hi = mid;
    }
    idx[0] = lo;
#pragma unroll
    for (int k = 1; k < kLrPer; ++k) {
      int i = idx[k - 1];
      while (i < T - 1 && __ldg(cm + i) <= f0 + k) ++i;   // zero-duration phonemes are skipped
      idx[k] = i;
    }
  }
#pragma unroll
  for (int k = 0; k < kLrPer; ++k) {
    const bool ok = f0 + k < F;
    if (!ok) idx[k] = idx[0];
    if (f2p && ok) f2p[(long long)b * f2p_ld + f0 + k] = idx[k];
    pv[k] = pitch ? pitch[(long long)b * pe_ld + idx[k]] : 0.f;
    ev[k] = energy ? energy[(long long)b * pe_ld + idx[k]] : 0.f;
  }
  const float* eb = enc + (long long)b * enc_bs;
  float* ob = out + (long long)b * out_bs + f0;
  const bool vec = f0 + kLrPer <= F && (out_ld & 3) == 0 && (out_bs & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 8
  for (int c = 0; c < C; ++c) {   // 8 channels x 4 gathers in flight per thread
    const float* er = eb + (long long)c * enc_ld;
    float v[kLrPer];
#pragma unroll
    for (int k = 0; k < kLrPer; ++k) {
      v[k] = __ldg(er + idx[k]);
      if (pitch) v[k] += fmaf(pv[k], __ldg(wp + c), __ldg(bp + c));
      if (energy) v[k] += fmaf(ev[k], __ldg(we + c), __ldg(be + c));
    }
    float* orow = ob + (long long)c * out_ld;
    if (vec) {
      *reinterpret_cast<float4*>(orow) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < kLrPer; ++k)
        if (f0 + k < F) orow[k] = v[k];
    }
  }
}

}  // namespace tb200

using namespace tb200;

extern "C" {

int tb200_duration_finalize(const float* log_dur, const int64_t* gold_dur, const float* text, const int32_t* text_len,
                            int32_t B, int32_t T_max, int32_t T_ld, float pause_scale, float duration_scale,
                            int64_t* dur_out, int32_t* cum_out, int32_t* frames_out, void* stream) {
  if ((!log_dur && !gold_dur) || !text || !text_len || !dur_out || !cum_out || !frames_out)
    return fail(TB200_E_BADARG, "duration_finalize: null pointer");
  if (B <= 0 || T_max <= 0 || T_ld < T_max) return fail(TB200_E_BADARG, "duration_finalize: bad shape");
  const int warps = 4;
  duration_finalize_kernel<<<(B + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      log_dur, reinterpret_cast<const long long*>(gold_dur), text, text_len, B, T_max, T_ld, pause_scale, duration_scale,
      reinterpret_cast<long long*>(dur_out), cum_out, frames_out);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_variance_edit(float* curve, const float* text, const int32_t* text_len, int32_t B, int32_t T_max, int32_t T_ld,
                        int32_t which, float variance_scale, void* stream) {
  if (!curve || !text || !text_len) return fail(TB200_E_BADARG, "variance_edit: null pointer");
  if (B <= 0 || T_max <= 0 || T_ld < T_max || which < 0 || which > 1) return fail(TB200_E_BADARG, "variance_edit: bad argument");
  const int warps = 4;
  variance_edit_kernel<<<(B + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      curve, text, text_len, B, T_max, T_ld, which == 0 ? kFeatVoiced : kFeatPhoneme, variance_scale);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_length_regulate(const float* enc, int64_t enc_bs, int32_t enc_ld, const float* pitch, const float* energy,
                          int32_t pe_ld, const float* wp, const float* bp, const float* we, const float* be,
                          const int32_t* cum, int32_t cum_ld, const int32_t* text_len, const int32_t* frames, int32_t B,
                          int32_t C, int32_t F_max, float* out, int64_t out_bs, int32_t out_ld, int32_t* frame_to_phone,
                          int32_t f2p_ld, void* stream) {
  if (!enc || !cum || !text_len || !frames || !out) return fail(TB200_E_BADARG, "length_regulate: null pointer");
  if ((pitch && (!wp || !bp)) || (energy && (!we || !be))) return fail(TB200_E_BADARG, "length_regulate: missing embed weights");
  if (B <= 0 || C <= 0 || F_max <= 0) return fail(TB200_E_BADARG, "length_regulate: bad shape");
  dim3 grid((F_max + 128 * kLrPer - 1) / (128 * kLrPer), B);
  length_regulate_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      enc, enc_bs, enc_ld, pitch, energy, pe_ld, wp, bp, we, be, cum, cum_ld, text_len, frames, C, out, out_bs, out_ld,
      frame_to_phone, f2p_ld);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // extern "C"
