// api.cu -- extern "C" surface of libtoucan_b200.so: argument validation, geometry derivation,
// weight packing, error reporting.  See include/toucan_b200.h for the contract.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "conv_common.cuh"

namespace tb200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// kaiser_sinc_filter1d(cutoff 0.25, half_width 0.3, kernel 12) of alias_free_torch, in double.
static double bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  for (int k = 1; k < 64; ++k) {
    term *= (x / (2.0 * k)) * (x / (2.0 * k));
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}

static void aa_filter_taps(float out[12]) {
  const int ks = 12, half = 6;
  const double cutoff = 0.25, half_width = 0.3, pi = 3.14159265358979323846;
  const double delta_f = 4 * half_width;
  const double att = 2.285 * (half - 1) * pi * delta_f + 7.95;
  double beta = 0.0;
  if (att > 50.0) beta = 0.1102 * (att - 8.7);
  else if (att >= 21.0) beta = 0.5842 * std::pow(att - 21.0, 0.4) + 0.07886 * (att - 21.0);
  double f[12], sum = 0.0;
  for (int i = 0; i < ks; ++i) {
    const double r = 2.0 * i / (ks - 1) - 1.0;  // torch.kaiser_window(periodic=False)
    const double win = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / bessel_i0(beta);
    const double tm = (i - half) + 0.5;
    const double xx = 2 * cutoff * tm;
    const double sinc = xx == 0.0 ? 1.0 : std::sin(pi * xx) / (pi * xx);
    f[i] = 2 * cutoff * win * sinc;
    sum += f[i];
  }
  for (int i = 0; i < ks; ++i) out[i] = static_cast<float>(f[i] / sum);
}

static int ensure_constants() {
  static bool done[kMaxDeviceSlots] = {};   // __constant__ memory is per device
  const int slot = current_device_slot();
  if (slot < 0) return fail(TB200_E_NODEVICE, "no current CUDA device");
  if (done[slot]) return 0;
  float taps[12];
  aa_filter_taps(taps);
  TB200_CUDA_CHECK(cudaMemcpyToSymbol(c_aa_filter, taps, sizeof(taps)));
  done[slot] = true;
  return 0;
}

static int next_pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}

int fill_conv_args(const tb200_conv1d_params* p, int precision, ConvArgs& a) {
  if (!p || !p->x || !p->y || !p->w_packed) return fail(TB200_E_BADARG, "conv1d: null pointer");
  if (p->B <= 0 || p->C_in <= 0 || p->C_out <= 0 || p->L_in_max <= 0) return fail(TB200_E_BADARG, "conv1d: empty shape");
  const int up = p->transposed_stride;
  if (up < 0 || (up > 0 && (p->K != 2 * up || (up & 1)))) return fail(TB200_E_BADARG, "conv1d: transposed conv needs even stride u and K == 2u");
  if (up == 0 && (p->K < 1 || p->K > kMaxTaps || p->dilation < 1)) return fail(TB200_E_BADARG, "conv1d: K must be in [1,%d]", kMaxTaps);
  if (p->act == TB200_ACT_AA_SNAKEBETA && (!p->act_alpha || !p->act_beta)) return fail(TB200_E_BADARG, "conv1d: snake activation needs alpha/beta");
  int rc = ensure_constants();
  if (rc) return rc;
  memset(&a, 0, sizeof(a));
  const ConvGeom g = conv_geom(p->C_in, p->C_out, p->K, up, precision == TB200_PREC_FP32_SIMT ? TB200_PREC_TF32 : precision);
  a.x = p->x; a.len_in = p->len_in; a.y = p->y; a.residual = p->residual; a.bias = p->bias;
  a.alpha = p->act_alpha; a.beta = p->act_beta; a.w = p->w_packed;
  a.x_bs = p->x_bs; a.y_bs = p->y_bs; a.r_bs = p->r_bs; a.x_ld = p->x_ld; a.y_ld = p->y_ld; a.r_ld = p->r_ld;
  a.B = p->B; a.Cin = p->C_in; a.Cin_pad = g.Cin_pad; a.Cout = p->C_out; a.L_in_max = p->L_in_max;
  a.ntaps = g.ntaps;
  int mn = 0, mx = 0;
  for (int j = 0; j < a.ntaps; ++j) {
    a.tap_off[j] = up > 0 ? -j : j * p->dilation - p->pad;
    mn = a.tap_off[j] < mn ? a.tap_off[j] : mn;
    mx = a.tap_off[j] > mx ? a.tap_off[j] : mx;
  }
  a.halo_l = -mn;
  a.R = kTileM + mx - mn;
  a.up = up; a.up_pad = up / 2;
  a.N_total = g.N_total; a.NT = g.NT; a.n_ntiles = g.n_ntiles;
  a.KC = g.KC; a.n_kchunks = g.n_kchunks; a.n_chunks = g.n_chunks;
  a.chunk_bytes = static_cast<int>(g.chunk_elems * g.elem_bytes);
  a.act = p->act; a.slope = p->act_slope;
  a.x_f16 = p->x_dtype == TB200_F16; a.y_f16 = p->y_dtype == TB200_F16; a.r_f16 = p->r_dtype == TB200_F16;
  a.out_act = p->out_act; a.out_alpha = p->out_alpha; a.res_beta = p->res_beta; a.accumulate = p->accumulate;
  const int rows_max = p->L_in_max + (up > 0 ? 1 : 0);
  a.tiles_per_utt = (rows_max + kTileM - 1) / kTileM;
  a.total_tiles = a.tiles_per_utt * p->B;
  a.tmem_cols = next_pow2_cols(a.NT);
  a.a_bytes = (a.Cin_pad / g.epc) * a.R * 16;
  a.S = 1; a.a_bufs = 1; a.acc_bufs = 1; a.n_panels = 1; a.aa_fast = 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// weight packing: torch fp32 -> [chunk = (ntile, tap, kchunk)][KC/E][NT][E] operand image
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Cin, int Cout, int K, int up,
                                   ConvGeom g) {
  const long long total = g.chunk_elems * g.n_chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = i % g.epc;
    const int nn = (i / g.epc) % g.NT;
    const int kg = (i / ((long long)g.epc * g.NT)) % (g.KC / g.epc);
    const long long chunk = i / g.chunk_elems;
    const int kc = chunk % g.n_kchunks;
    const int j = (chunk / g.n_kchunks) % g.ntaps;
    const int nt = chunk / ((long long)g.n_kchunks * g.ntaps);
    const int n = nt * g.NT + nn;
    const int ci = kc * g.KC + kg * g.epc + e;
    float v = 0.f;
    if (n < g.N_total && ci < Cin) {
      if (up > 0) {
        const int co = n / up, ph = n - co * up;
        v = w[((long long)ci * Cout + co) * (2 * up) + ph + j * up];
      } else {
        v = w[((long long)n * Cin + ci) * K + j];
      }
    }
    if constexpr (sizeof(T) == 2) out[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    else out[i] = round_tf32(v);
  }
}

int conv1d_umma(const tb200_conv1d_params* p, cudaStream_t stream);
int conv_trace_read(long long* host_out, int n);
int conv1d_simt(const tb200_conv1d_params* p, cudaStream_t stream);

}  // namespace tb200

using namespace tb200;

extern "C" {

int tb200_version(void) { return TB200_VERSION; }

const char* tb200_last_error(void) { return g_err; }

int tb200_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return TB200_E_NODEVICE;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return TB200_E_NODEVICE;
  return n;
}

int64_t tb200_packed_weight_bytes(int32_t C_in, int32_t C_out, int32_t K, int32_t up, int32_t precision) {
  if (precision == TB200_PREC_FP32_SIMT) return (int64_t)C_in * C_out * K * 4;
  return conv_geom(C_in, C_out, K, up, precision).packed_bytes;
}

int tb200_pack_conv_weight(const float* w, void* w_packed, int32_t C_in, int32_t C_out, int32_t K, int32_t up,
                           int32_t precision, void* stream) {
  if (!w || !w_packed) return fail(TB200_E_BADARG, "pack: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precision == TB200_PREC_FP32_SIMT) {
    TB200_CUDA_CHECK(cudaMemcpyAsync(w_packed, w, (size_t)C_in * C_out * K * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const ConvGeom g = conv_geom(C_in, C_out, K, up, precision);
  const long long total = g.chunk_elems * g.n_chunks;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  if (precision == TB200_PREC_F16)
    pack_weight_kernel<__half><<<blocks, 256, 0, s>>>(w, reinterpret_cast<__half*>(w_packed), C_in, C_out, K, up, g);
  else
    pack_weight_kernel<float><<<blocks, 256, 0, s>>>(w, reinterpret_cast<float*>(w_packed), C_in, C_out, K, up, g);
  TB200_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int tb200_debug_trace_read(int64_t* host_out, int32_t n) {
  if (!host_out || n <= 0) return fail(TB200_E_BADARG, "trace: bad argument");
  return conv_trace_read(reinterpret_cast<long long*>(host_out), n);
}

int tb200_conv1d(const tb200_conv1d_params* p, void* stream) {
  if (!p) return fail(TB200_E_BADARG, "conv1d: null params");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->precision) {
    case TB200_PREC_FP32_SIMT: return conv1d_simt(p, s);
    case TB200_PREC_F16:
    case TB200_PREC_TF32: return conv1d_umma(p, s);
    default: return fail(TB200_E_BADARG, "conv1d: unknown precision %d", p->precision);
  }
}

}  // extern "C"
