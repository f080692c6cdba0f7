/*
 * toucan_b200.h -- C ABI of libtoucan_b200.so, the B200 (sm_100a) engine for the
 * IMS-Toucan hot path: ToucanTTS.inference -> vocoder Generator.
 *
 * The reference (PaulMayer123/IMS-Toucan-Prosody-Variance) is pure Python/PyTorch:
 * the "FFI" it would bind is ctypes from the nn.Module forwards.  Every entry point
 * below names the reference op group it replaces (paths relative to the reference
 * root; K-numbers are SURVEY.md section 2.2).
 *
 * Conventions
 *   - plain C: pointers, ints, floats.  No torch types.  All pointers are DEVICE
 *     pointers owned by the caller (tensor.data_ptr()); the library never allocates,
 *     frees or retains them.  Work is enqueued on `stream` (a cudaStream_t passed as
 *     void*); no call synchronises the device, so CUDA-graph capture works.
 *   - return value: 0 ok; <0 library error (TB200_E_*); >0 a cudaError_t.
 *     tb200_last_error() returns a thread-local human-readable message.
 *   - activations are "NCL": x[b][c][l], l contiguous, row pitch `ld` elements,
 *     batch stride `bs` elements.  Ragged batches carry a device int32 array of
 *     per-utterance valid lengths; positions >= length behave exactly like the
 *     zero padding a batch-1 reference call would see (no cross-utterance leakage),
 *     and are never written.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef TOUCAN_B200_H
#define TOUCAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB200_VERSION 104

enum {
  TB200_OK = 0,
  TB200_E_BADARG = -1,      /* invalid shapes / unsupported configuration            */
  TB200_E_NOSMEM = -2,      /* configuration does not fit in shared / tensor memory  */
  TB200_E_NODEVICE = -3     /* no sm_100 device                                      */
};

/* element types of activation tensors */
enum { TB200_F32 = 0, TB200_F16 = 1 };

/* tensor-core operand precision of tb200_conv1d */
enum {
  TB200_PREC_FP32_SIMT = 0, /* fp32 CUDA-core FMA (exact-parity mode)                          */
  TB200_PREC_F16 = 1,       /* tcgen05 kind::f16, fp16 operands (10-bit mantissa), fp32 accum  */
  TB200_PREC_TF32 = 2       /* tcgen05 kind::tf32, fp32 accum                                  */
};

/* prologue activation applied to the input before the convolution */
enum {
  TB200_ACT_NONE = 0,
  TB200_ACT_LEAKY_RELU = 1, /* HiFiGAN: ResidualBlock.py:83-98, InferenceAvocodo.py:36-43,53   */
  TB200_ACT_AA_SNAKEBETA = 2, /* BigVGAN: alias_free_torch.Activation1d(SnakeBeta), AMP.py:45-57,
                                 Snake.py:56-69 -- 2x kaiser-sinc up, x+1/(e^b+1e-9) sin^2(x e^a),
                                 2x down; replicate padding at the utterance's own ends        */
  TB200_ACT_RELU = 3,
  TB200_ACT_SWISH = 4,
  TB200_ACT_TANH = 5
};

/* epilogue activation applied to (acc + bias) */
enum { TB200_OUT_NONE = 0, TB200_OUT_TANH = 1, TB200_OUT_RELU = 2 };

int tb200_version(void);
const char* tb200_last_error(void);
/* number of SMs of the current device, or <0 */
int tb200_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * tb200_conv1d -- implicit-GEMM Conv1d / ConvTranspose1d / Linear (k=1) with fused prologue
 * activation and fused bias / activation / scale / residual / accumulate epilogue.
 * Replaces (K3,K5,K6,K8,K9,K10,K11): torch.nn.Conv1d / ConvTranspose1d / Linear call sites in
 *   InferenceAvocodo.py:69-80, Layers/ResidualBlock.py:83-98          (HiFiGAN generator)
 *   InferenceBigVGAN.py:72-95, BigVGAN/AMP.py:51-60                   (BigVGAN generator)
 *   Layers/MultiLayeredConv1d.py:40-51, Convolution.py:43-55, Attention.py:42-64,92
 *   Layers/DurationPredictor.py:66-77, VariancePredictor.py:67-77, PostNet.py:62-74
 *   ToucanTTS/Glow.py:248-269,342-365, wavenet.py:89-122
 *
 *   y[b][co][t] = out_alpha * OUT( bias[co] + sum_{ci,j} W[co][ci][j] * ACT(x)[b][ci][t - pad + j*dilation] )
 *               + res_beta * residual[b][co][t]  (+ y_old[b][co][t] if accumulate)
 * for 0 <= t < len_out[b];  ACT(x) is zero outside [0, len_in[b]).
 * transposed_stride u > 0 selects ConvTranspose1d(kernel 2u, stride u, padding u/2):
 *   len_out = u*len_in, weight given as (Cin, Cout, 2u) (torch layout).
 * ---------------------------------------------------------------------------------------- */
typedef struct tb200_conv1d_params {
  /* input */
  const void* x;            /* (B, C_in, L) NCL                                         */
  int32_t x_dtype;          /* TB200_F32 | TB200_F16                                    */
  int64_t x_bs;             /* batch stride, elements                                   */
  int32_t x_ld;             /* row pitch, elements                                      */
  const int32_t* len_in;    /* (B) valid input lengths, device; NULL = L_in_max for all */
  int32_t B, C_in, L_in_max;
  /* convolution */
  int32_t C_out, K, dilation, pad;
  int32_t transposed_stride;
  const void* w_packed;     /* from tb200_pack_conv_weight with the same geometry       */
  const float* bias;        /* (C_out) or NULL                                          */
  int32_t precision;        /* TB200_PREC_*                                             */
  /* prologue */
  int32_t act;              /* TB200_ACT_*                                              */
  float act_slope;          /* leaky relu negative slope                                */
  const float* act_alpha;   /* (C_in) log-scale alpha for AA_SNAKEBETA                  */
  const float* act_beta;    /* (C_in) log-scale beta                                    */
  /* epilogue */
  int32_t out_act;          /* TB200_OUT_*                                              */
  float out_alpha;
  const void* residual;     /* (B, C_out, L_out) or NULL                                */
  int32_t r_dtype;          /* TB200_F32 | TB200_F16                                    */
  int64_t r_bs; int32_t r_ld;
  float res_beta;
  int32_t accumulate;       /* y += ...                                                 */
  void* y;                  /* (B, C_out, L_out)                                        */
  int32_t y_dtype; int64_t y_bs; int32_t y_ld;
} tb200_conv1d_params;

int tb200_conv1d(const tb200_conv1d_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * tb200_respair -- one residual pair of the vocoder generators in ONE launch; the tensor between the
 * two convolutions never leaves the SM.  Replaces the loop body
 *     xt = a1(x); xt = c1(xt); xt = a2(xt); xt = c2(xt); x = xt + x
 *   BigVGAN AMPBlock1.forward   TrainingInterfaces/Spectrogram_to_Wave/BigVGAN/AMP.py:51-60
 *   HiFiGAN HiFiGANResidualBlock.forward   Layers/ResidualBlock.py:93-97
 * and, through out_alpha / res_beta / accumulate, the multi-receptive-field mean of
 * InferenceAvocodo.py:75-78 / InferenceBigVGAN.py:82-89:
 *   y[b][c][t] = out_alpha * (b2[c] + conv2(ACT2(b1 + conv1(ACT1(x))))[b][c][t]) + res_beta * x[b][c][t]
 *              (+ y_old[b][c][t] if accumulate)                        for 0 <= t < len[b]
 *   conv1: Conv1d(C, C, K, dilation, padding (K-1)/2*dilation); conv2: Conv1d(C, C, K, padding (K-1)/2).
 *   ACT: TB200_ACT_LEAKY_RELU (act_slope in [0,1]) or TB200_ACT_AA_SNAKEBETA (per-channel alpha/beta of
 *   each of the two activations).  ACT(x) and the conv inputs are zero outside [0, len[b]); the
 *   anti-aliasing filters replicate-pad at the utterance's own ends (batch-1 semantics).
 * Operands are fp16 (tcgen05 kind::f16, fp32 accumulate): w1/w2 from tb200_pack_conv_weight(C, C, K, 0,
 * TB200_PREC_F16).  x: 16-byte aligned, row pitch / batch stride multiples of 8 (fp16) or 4 (fp32)
 * elements, row pitch >= L_max rounded up to that unit.  y must not alias x.  C in {32, 64, 128}.
 * ---------------------------------------------------------------------------------------- */
typedef struct tb200_respair_params {
  const void* x;            /* (B, C, L) NCL residual stream in                         */
  int32_t x_dtype;          /* TB200_F32 | TB200_F16                                    */
  int64_t x_bs; int32_t x_ld;
  const int32_t* len;       /* (B) valid lengths, device; NULL = L_max for all          */
  int32_t B, C, L_max;
  int32_t K, dilation;      /* conv1 dilation; conv2 has dilation 1                     */
  const void* w1_packed; const float* bias1;
  const void* w2_packed; const float* bias2;
  int32_t act;              /* TB200_ACT_LEAKY_RELU | TB200_ACT_AA_SNAKEBETA            */
  float act_slope;
  const float *act1_alpha, *act1_beta;   /* (C) log-scale snake parameters before conv1 */
  const float *act2_alpha, *act2_beta;   /* (C) ... before conv2                        */
  float out_alpha, res_beta;
  int32_t accumulate;
  void* y;                  /* (B, C, L) residual stream out                            */
  int32_t y_dtype; int64_t y_bs; int32_t y_ld;
} tb200_respair_params;

int tb200_respair(const tb200_respair_params* p, void* stream);

/* tb200_conv1d_staged -- ONE 'same'-padded C -> C convolution of the generators' residual stacks through the
 * pipeline of tb200_respair: the raw input tile is brought into shared memory by the TMA engine (cp.async.bulk), the
 * prologue activation (LeakyReLU / anti-aliased SnakeBeta) is staged from shared memory into double-buffered
 * tensor-core operand tiles, the epilogue (bias, out_alpha, res_beta * residual, accumulate) moves 16 bytes per
 * access.  Same parameter block and result as tb200_conv1d for the cases it accepts: precision TB200_PREC_F16,
 * C_in == C_out in {32, 64, 128}, odd K with pad == (K-1)/2*dilation, transposed_stride 0, act LEAKY_RELU or
 * AA_SNAKEBETA, out_act NONE; x, y, residual 16-byte aligned rows (pitch / batch stride multiples of 16 bytes);
 * y must not alias x.  Returns TB200_E_BADARG otherwise (callers fall back to tb200_conv1d).               */
int tb200_conv1d_staged(const tb200_conv1d_params* p, void* stream);

/* Debugging aid: like tb200_debug_trace_read, for tb200_respair (16 stamps per tile).            */
int tb200_respair_trace_read(int64_t* host_out, int32_t n);

/* Bytes of the packed weight blob for a conv geometry and precision. */
int64_t tb200_packed_weight_bytes(int32_t C_in, int32_t C_out, int32_t K, int32_t transposed_stride, int32_t precision);

/* Pack torch-layout fp32 weights (device) into the operand image tb200_conv1d streams into
 * shared memory.  w: (C_out, C_in, K) for Conv1d/Linear, (C_in, C_out, 2u) for ConvTranspose1d.
 * Load-time only (the fold of weight_g*v/||v|| -- remove_weight_norm,
 * InferenceAvocodo.py:82-89, InferenceBigVGAN.py:97-105, InferenceToucanTTS.py:321-330 --
 * happens on the host side before packing). */
int tb200_pack_conv_weight(const float* w, void* w_packed, int32_t C_in, int32_t C_out, int32_t K,
                           int32_t transposed_stride, int32_t precision, void* stream);

/* ------------------------------------------------------------------------------------------
 * K7 / K12 -- ragged duration rounding, prosody edits, prefix sum, frame expansion.
 * Integer results are bit-exact with the reference given the same fp32 inputs.
 * ---------------------------------------------------------------------------------------- */

/* DurationPredictor.py:79  d = clamp(round(exp(x) - 1), 0) (round-half-even), then the edit
 * loop of InferenceToucanTTS.py:214-225 (word boundary -> 0, silence * pause scale, global
 * scale; each round(float(d) * s)), then the per-utterance LengthRegulator rescue
 * (LengthRegulator.py:52-53, batch-1 semantics: an utterance whose durations sum to 0 gets all
 * ones) and the inclusive prefix sum.  One warp per utterance.
 *   log_dur   (B, T_ld) fp32 predicted log durations, or NULL when gold_dur is given
 *   gold_dur  (B, T_ld) int64 external durations or NULL  (not modified; reference mutates)
 *   text      (B, T_max, 62) fp32 articulatory features (row pitch 62)
 *   text_len  (B) int32
 *   dur_out   (B, T_ld) int64 final durations
 *   cum_out   (B, T_ld) int32 inclusive prefix sums
 *   frames_out(B) int32 total frames per utterance                                          */
int tb200_duration_finalize(const float* log_dur, const int64_t* gold_dur, const float* text,
                            const int32_t* text_len, int32_t B, int32_t T_max, int32_t T_ld,
                            float pause_scale, float duration_scale,
                            int64_t* dur_out, int32_t* cum_out, int32_t* frames_out, void* stream);

/* Pitch / energy edits + _scale_variance (InferenceToucanTTS.py:214-218, 226-227, 333-343):
 * unvoiced (feature 61 == 0) -> pitch 0; non-phoneme (feature 15 == 0) -> energy 0; then
 * (x - mean_nonzero) * s + mean_nonzero, negatives -> 0, skipped when s == 1.
 * curve (B, T_ld) fp32 in/out; which: 0 pitch, 1 energy.                                    */
int tb200_variance_edit(float* curve, const float* text, const int32_t* text_len, int32_t B,
                        int32_t T_max, int32_t T_ld, int32_t which, float variance_scale, void* stream);

/* LengthRegulator.forward (LengthRegulator.py:37-61, utils.py:475-494) fused with the
 * pitch/energy embedding add (InferenceToucanTTS.py:230-232):
 *   out[b][c][f] = enc[b][c][i] + pitch[b][i]*wp[c] + bp[c] + energy[b][i]*we[c] + be[c],
 *   i = #(cum[b][:] <= f)  (searchsorted right), f < frames[b].
 * enc NCL (B,C,T) fp32; out NCL (B,C,F) fp32; frame_to_phone (B, F_ld) int32 optional.      */
int tb200_length_regulate(const float* enc, int64_t enc_bs, int32_t enc_ld,
                          const float* pitch, const float* energy, int32_t pe_ld,
                          const float* wp, const float* bp, const float* we, const float* be,
                          const int32_t* cum, int32_t cum_ld, const int32_t* text_len,
                          const int32_t* frames, int32_t B, int32_t C, int32_t F_max,
                          float* out, int64_t out_bs, int32_t out_ld,
                          int32_t* frame_to_phone, int32_t f2p_ld, void* stream);

/* Debugging aid (not part of the reference-facing surface): with the environment variable TB200_TRACE set,
 * tb200_conv1d records clock64() stamps of its pipeline roles for the first tiles of CTA 0; this copies up to
 * n int64 values (8 per tile) to host memory after a device synchronisation by the caller.               */
int tb200_debug_trace_read(int64_t* host_out, int32_t n);

/* ------------------------------------------------------------------------------------------
 * Acoustic model (ToucanTTS) -- the non-GEMM kernels.  All fp32, NCL tensors, ragged by `len`
 * (device int32 (B) or NULL = L_max): positions >= len[b] are neither read as data nor written.
 * ---------------------------------------------------------------------------------------- */

/* LayerNorm over the channel axis (Layers/LayerNorm.py:17, eps 1e-12, biased variance) --
 * mode 0: y = (x - mean) * rsqrt(var + eps) * gamma[c] + beta[c]
 * or ConditionalLayerNorm (Layers/ConditionalLayerNorm.py:52-67) --
 * mode 1: y = gamma[b][c] * (x - mean) / var + beta[b][c]   (VARIANCE, no epsilon: reference quirk).
 * gamma/beta are (C) with gb_bs = 0, or (B, C) with batch stride gb_bs.  C <= 256.           */
int tb200_channel_norm(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                       const int32_t* len, int32_t B, int32_t C, int32_t L_max,
                       const float* gamma, const float* beta, int64_t gb_bs, int32_t mode, float eps, void* stream);

/* GroupNorm over (channels of a group) x (the utterance's own frames) + optional tanh + optional
 * residual add: y = residual + OUT(gn(x) * gamma + beta)   (Layers/PostNet.py:45-59,62-74 and the
 * `mel + postnet(mel)` of InferenceToucanTTS.py:241).  out_act: TB200_OUT_NONE | TB200_OUT_TANH.   */
int tb200_group_norm(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                     const float* residual, int64_t r_bs, int32_t r_ld, const int32_t* len, int32_t B, int32_t C,
                     int32_t L_max, int32_t groups, const float* gamma, const float* beta, float eps, int32_t out_act,
                     void* stream);

/* Middle of the Conformer convolution module (Layers/Convolution.py:43-52): GLU over channels of
 * x (B, 2C, L), depthwise Conv1d (C, K) zero padded at the utterance's own ends + bias, eval-mode
 * BatchNorm1d ((h - mean) * rsqrt(var + eps) * gamma + beta), Swish.  y (B, C, L).  K odd, <= 63. */
int tb200_glu_dwconv(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                     const int32_t* len, int32_t B, int32_t C, int32_t L_max, const float* w, const float* bias, int32_t K,
                     const float* bn_mean, const float* bn_var, const float* bn_gamma, const float* bn_beta, float bn_eps,
                     void* stream);

/* Relative-position multi-head attention core (Layers/Attention.py:159-198 with rel_shift :138-157
 * and forward_attention :66-92), flash style (no (T, 2T-1) score tensor):
 *   score(i,j) = ((q_i + u_h) . k_j + (q_i + v_h) . p_{i-j}) / sqrt(dk)  for j < len[b];
 *   out_i = sum_j softmax_j(score(i, .)) v_j.
 * qkv (B, 3*H*dk, L): rows [0,D) q, [D,2D) k, [2D,3D) v (D = H*dk, head h = rows h*dk..);
 * pos (D, pos_cols): column (pos_center - r) holds linear_pos(PE(r)) for relative position r;
 * bias_u / bias_v (H, dk); out (B, D, L).  dk in {32, 48, 64}.                                  */
int tb200_relpos_attention(const float* qkv, int64_t qkv_bs, int32_t qkv_ld, const float* pos, int32_t pos_ld,
                           int32_t pos_center, int32_t pos_cols, const float* bias_u, const float* bias_v,
                           const int32_t* len, int32_t B, int32_t H, int32_t dk, int32_t L_max,
                           float* out, int64_t out_bs, int32_t out_ld, void* stream);

/* The same contraction on the tensor cores (tcgen05.mma kind::f16, fp32 accumulation in TMEM): QK^T, the
 * relative-position band (Q + v) P^T and softmax x V are 128 x 128 / 128 x 256 / 128 x dk MMA tiles per (utterance,
 * head, 128 queries); the skewed read of the band (rel_shift, Attention.py:138-157) goes through a per-row window in
 * shared memory.  Operands are rounded to fp16 (tf32's mantissa): used by the tf32 / f16 precision modes, the fp32
 * mode keeps tb200_relpos_attention.
 * pos16: the projected positional table packed once per layer as fp16 operand rows [H][dk/8][pos_rows][8]
 * (16 bytes per (head, d-group, relative position)); row (pos_center + r) is relative position r, rows outside the
 * table's own range are zero; pos_rows must cover r in [-(L_max+254), L_max+127].  qkv rows 16-byte aligned.     */
int tb200_relpos_attention_tc(const float* qkv, int64_t qkv_bs, int32_t qkv_ld, const void* pos16, int32_t pos_rows,
                              int32_t pos_center, const float* bias_u, const float* bias_v,
                              const int32_t* len, int32_t B, int32_t H, int32_t dk, int32_t L_max,
                              float* out, int64_t out_bs, int32_t out_ld, void* stream);

/* y[b][c][t] = (x[b][c][t] + vec[b][c]) * scale; x or vec may be NULL (treated as 0).  Covers the
 * language-embedding add and the sqrt(adim) scale of RelPositionalEncoding (Conformer.py:108-118,
 * PositionalEncoding.py:119-130) and the broadcast of the utterance embedding that
 * Conformer.py:131-134 concatenates to the encoder output.                                      */
int tb200_rowvec_affine(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                        const int32_t* len, int32_t B, int32_t C, int32_t L_max, const float* vec, int64_t vec_bs,
                        float scale, void* stream);

/* (B, L, C) row-major <-> NCL (B, C, L).  to_ncl != 0: in is (B,L,C) with row pitch in_ld, out NCL;
 * else in is NCL and out (B,L,C) (the reference's module-boundary layout).                       */
int tb200_transpose(const float* in, int64_t in_bs, int32_t in_ld, float* out, int64_t out_bs, int32_t out_ld,
                    const int32_t* len, int32_t B, int32_t C, int32_t L_max, int32_t to_ncl, void* stream);

/* Glow squeeze / unsqueeze with n_sqz = 2 (ToucanTTS/glow_utils.py:28-53).
 * inverse == 0: x (B,C,L) -> y (B,2C,L/2), y[s*C+c][tau] = x[c][2tau+s]; len = unsqueezed lengths
 * (an odd last frame is dropped).  inverse != 0: x (B,2C,L2) -> y (B,C,2*L2); len = squeezed.     */
int tb200_squeeze2(const float* x, int64_t x_bs, int32_t x_ld, float* y, int64_t y_bs, int32_t y_ld,
                   const int32_t* len, int32_t B, int32_t C, int32_t L_max, int32_t inverse, void* stream);

/* WaveNet gate (ToucanTTS/wavenet.py:29-35,102-111): y[c] = tanh(a[c]) * sigmoid(a[c + hidden]);
 * a (B, 2*hidden, L), y (B, hidden, L).                                                          */
int tb200_wn_gate(const float* a, int64_t a_bs, int32_t a_ld, float* y, int64_t y_bs, int32_t y_ld,
                  const int32_t* len, int32_t B, int32_t hidden, int32_t L_max, void* stream);

/* Closes one reversed flow block in place on x (B, C, L) (ToucanTTS/Glow.py:260-263, 116-128, 30-32):
 * coupling^-1  x[C/2:] = (x[C/2:] - ml[:C/2]) * exp(-ml[C/2:]);  invconv^-1 with the cached 4x4
 * inverse w_inv (row-major) over the channel groups of InvConvNear; actnorm^-1
 * (x - an_bias[c]) * exp(-an_logs[c]).                                                           */
int tb200_flow_close(float* x, int64_t x_bs, int32_t x_ld, const float* ml, int64_t ml_bs, int32_t ml_ld,
                     const int32_t* len, int32_t B, int32_t C, int32_t L_max, const float* w_inv,
                     const float* an_bias, const float* an_logs, void* stream);

/* torch.nn.functional.normalize on (B, C) rows (InferenceToucanTTS.py:202, Conformer.py:132).     */
int tb200_l2_normalize(const float* x, float* y, int32_t B, int32_t C, void* stream);

/* The N stacked conditioning MLPs of ConditionalLayerNorm (ConditionalLayerNorm.py:27-50):
 * out[n][b] = W4_n tanh(W2_n tanh(W0_n e_b + b0_n) + b2_n) + b4_n;  e (B,E); w0 (N,E,E), w2 (N,Cc,E),
 * w4 (N,Cc,Cc) in torch Linear layout (out, in); out (N, B, Cc).                                  */
int tb200_cln_mlp(const float* e, int32_t B, int32_t E, int32_t Cc, int32_t N, const float* w0, const float* b0,
                  const float* w2, const float* b2, const float* w4, const float* b4, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOUCAN_B200_H */
