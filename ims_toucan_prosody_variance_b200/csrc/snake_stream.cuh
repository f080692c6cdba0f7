// snake_stream.cuh -- BigVGAN's anti-aliased SnakeBeta (alias_free_torch.Activation1d(SnakeBeta): AMP.py:45-57,
// Snake.py:56-69 of the reference) as a streaming filter, lane = channel, sequential in time.  Shared by the per-layer
// conv kernel (source: global memory) and the fused residual-pair kernel (source: shared-memory tiles).
#pragma once
#include <type_traits>

#include "conv_common.cuh"

namespace tb200 {

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <typename T>
__device__ __forceinline__ T to_operand(float v);
template <>
__device__ __forceinline__ __half to_operand<__half>(float v) {
  return f16_sat(v);
}
template <>
__device__ __forceinline__ float to_operand<float>(float v) {
  return round_tf32(v);  // the tensor core would truncate fp32 -> tf32; round to nearest instead
}

// ---------------------------------------------------------------------------------------------
// producer, anti-aliased SnakeBeta, interior tiles: lane = channel, sequential in time with register
// sliding windows (8 inputs, 8 (s_even, s_odd) pairs) -> one 16-byte load per 4 (fp32) / 8 (fp16)
// steps per lane, 24 filter FMAs + 2 sin per output, no scratch, no intra-warp exchange.
//   stream step n (absolute time): ingest x[n]; pair(n-3) = snake(up-filter around n-3);
//   out(n-6) = down-filter over pairs n-9 .. n-3.
// A warp task = (32-channel block, row segment).  Requires every touched x index inside [0, len).
// ---------------------------------------------------------------------------------------------
// 8 consecutive inputs of one channel as loaded (fp16: 4 words, fp32: 8 words); converted where they are used, so the
// loads can run far ahead of their first consumer and an fp16 buffer costs 4 registers
template <bool XF16>
struct XRaw {
  uint32_t w[XF16 ? 4 : 8];
};
template <bool XF16>
__device__ __forceinline__ float xget(const XRaw<XF16>& b, int i) {
  if constexpr (XF16) {
    const __half2 h = *reinterpret_cast<const __half2*>(&b.w[i >> 1]);
    return (i & 1) ? __high2float(h) : __low2float(h);
  } else {
    return __uint_as_float(b.w[i]);
  }
}
// SMEM: the source is a shared-memory tile (plain loads) instead of global memory (read-only path)
template <bool SMEM>
__device__ __forceinline__ uint4 ld16(const uint4* p) {
  if constexpr (SMEM) return *p;
  else return __ldg(p);
}
template <bool SMEM>
__device__ __forceinline__ float ld_x1(const void* x, bool f16, long long idx) {
  if constexpr (SMEM) return f16 ? __half2float(reinterpret_cast<const __half*>(x)[idx]) : reinterpret_cast<const float*>(x)[idx];
  else return load_x(x, f16, idx);
}
template <bool XF16, bool SMEM>
__device__ __forceinline__ void load8(const void* x, long long idx, XRaw<XF16>& r) {
  if constexpr (XF16) {
    const uint4 u = ld16<SMEM>(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(x) + idx));
    r.w[0] = u.x; r.w[1] = u.y; r.w[2] = u.z; r.w[3] = u.w;
  } else {
    const uint4 p0 = ld16<SMEM>(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(x) + idx));
    const uint4 p1 = ld16<SMEM>(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(x) + idx + 4));
    r.w[0] = p0.x; r.w[1] = p0.y; r.w[2] = p0.z; r.w[3] = p0.w;
    r.w[4] = p1.x; r.w[5] = p1.y; r.w[6] = p1.z; r.w[7] = p1.w;
  }
}

// kaiser_sinc_filter1d(cutoff 0.25, half-width 0.3, 12 taps) of alias_free_torch as compile-time immediates for the
// streaming path (FFMA with an immediate operand: no constant-bank reloads in the inner loop).  Same values as the
// host-computed c_aa_filter: float(double closed form).
__device__ __forceinline__ constexpr float aa_tap(int k) {
  constexpr float t[12] = {2.028966555e-03f, 9.389463812e-03f, -2.554346435e-02f, -5.765737593e-02f, 1.285726130e-01f,
                           4.432097971e-01f, 4.432097971e-01f, 1.285726130e-01f, -5.765737593e-02f, -2.554346435e-02f,
                           9.389463812e-03f, 2.028966555e-03f};
  return t[k];
}


// s = u + ib * sin^2(u * ea).  TB200_SNAKE_SIN: 0 = sin.approx (MUFU.SIN, quarter-rate XU pipe), 2 = range reduction +
// polynomial on the FMA pipe (sin^2(pi r) = r^2 P(r^2), |r| <= 1/2, max abs error 8.5e-7), 1 = no sine (timing only).
#ifndef TB200_SNAKE_SIN
#define TB200_SNAKE_SIN 0
#endif
__device__ __forceinline__ float snake_value(float u, float ea, float ib) {
#if TB200_SNAKE_SIN == 0
  const float z = __sinf(u * ea);
  return fmaf(ib * z, z, u);
#elif TB200_SNAKE_SIN == 1
  const float z = u * ea;
  return fmaf(ib * z, z, u);
#else
  const float t = u * (ea * 0.318309886183790672f);           // theta / pi  (sin^2 has period pi)
  const float k = (t + 12582912.f) - 12582912.f;             // rint(t) for |t| < 2^22
  const float r = t - k;
  const float x = r * r;
  float p = fmaf(x, 10.57306957244873f, -29.421052932739258f);
  p = fmaf(x, p, 42.64226531982422f);
  p = fmaf(x, p, -32.46503448486328f);
  p = fmaf(x, p, 9.869522094726562f);
  return fmaf(ib * x, p, u);
#endif
}

// One (32-channel block, row segment) task.  EDGE segments (near the utterance's ends) read x with clamped indices
// (replicate padding of the 2x up-sampler), clamp the snake output index to [0, 2 len) (replicate padding of the
// down-sampler) and emit zeros outside [0, len); interior segments are compiled without any of these branches.
#ifndef TB200_SNAKE_PFL1
#define TB200_SNAKE_PFL1 0   // steady blocks: L1 prefetch distance in time steps (0 = off)
#endif
#ifndef TB200_SNAKE_PACKED
#define TB200_SNAKE_PACKED 1   // steady snake blocks on FFMA2 pairs (0: scalar FFMA blocks)
#endif
// xsrc[row + t] is the input of this lane's channel at absolute time t (global memory, or a shared-memory tile when
// SMEM); ea = e^alpha, ib = 1/(e^beta + 1e-9) of the channel; dst = operand row of time t_lo (E elements per row).
template <typename T, bool EDGE, bool XF16, bool SMEM>
__device__ __forceinline__ void aa_channel_task(const void* xsrc, long long row, float ea, float ib, int t_lo, int t_beg,
                                                int t_end, int len, T* dst) {
  constexpr int E = 16 / sizeof(T);
  const int ts = (t_beg - 9) & ~7;        // first ingested step, 16-byte aligned
  float xw[8], sv[16];
  XRaw<XF16> cur, n1, n2;
  auto loadx = [&](int base, XRaw<XF16>& r) {
    if (!EDGE || (base >= 0 && base + 8 <= len)) {
      load8<XF16, SMEM>(xsrc, row + base, r);
    } else {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ld_x1<SMEM>(xsrc, XF16, row + min(max(base + i, 0), len - 1));
      if constexpr (XF16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);   // exact: the values are fp16
          r.w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.w[i] = __float_as_uint(v[i]);
      }
    }
  };
  float s_first = 0.f, s_last = 0.f;
  if (EDGE && t_beg < 3) {                // s[0]: what the down-sampler sees left of the utterance
    float ue = 0.f, uo = 0.f;
#pragma unroll
    for (int q = 0; q < 6; q += 2) {
      ue = fmaf(ld_x1<SMEM>(xsrc, XF16, row + min(max(q - 3, 0), len - 1)), 2.f * aa_tap(11 - 2 * q), ue);
      uo = fmaf(ld_x1<SMEM>(xsrc, XF16, row + min(max(q - 2, 0), len - 1)), 2.f * aa_tap(9 - 2 * q), uo);
    }
    const float u = ue + uo;
    s_first = snake_value(u, ea, ib);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) xw[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) sv[i] = 0.f;
  loadx(ts, cur);
  loadx(ts + 8, n1);
  // one 8-step block; CHECK = false once every step both produces a pair and emits an output row
  auto block8 = [&](auto check_tag, int base) {
    constexpr bool CHECK = decltype(check_tag)::value;
    T* drow = dst + (long long)(base - 6 - t_lo) * E;   // row of step j = 0 (t = base - 6); +E per step
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = base + j;
      xw[j] = xget(cur, j);
      if (!CHECK || n >= t_beg) {  // pair(n-3) is first needed by out(t_beg)
        // even and odd taps in separate chains (the order the packed steady block needs; see block8_packed)
        float u0e = 0.f, u0o = 0.f, u1e = 0.f, u1o = 0.f;
#pragma unroll
        for (int q = 0; q < 6; q += 2) {
          u0e = fmaf(xw[(j + 2 + q) & 7], 2.f * aa_tap(11 - 2 * q), u0e);   // x[n-6+q]; the x2 of the up-sampler is exact
          u0o = fmaf(xw[(j + 3 + q) & 7], 2.f * aa_tap(9 - 2 * q), u0o);    // x[n-6+(q+1)]
          u1e = fmaf(xw[(j + 3 + q) & 7], 2.f * aa_tap(10 - 2 * q), u1e);   // x[n-5+q]
          u1o = fmaf(xw[(j + 4 + q) & 7], 2.f * aa_tap(8 - 2 * q), u1o);    // x[n-5+(q+1)]
        }
        const float u0 = u0e + u0o, u1 = u1o + u1e;
        float s0 = snake_value(u0, ea, ib), s1 = snake_value(u1, ea, ib);
        if constexpr (EDGE) {
          const int pi = n - 3;
          if (pi < 0) s0 = s1 = s_first;
          else if (pi >= len) s0 = s1 = s_last;
          else s_last = s1;
        }
        sv[2 * ((j + 5) & 7)] = s0;
        sv[2 * ((j + 5) & 7) + 1] = s1;
      }
      const int t = n - 6;
      if (!CHECK || (t >= t_beg && t < t_end)) {
        // out[t] = sum_i tap[2i] s1[t-3+i] + tap[2i+1] s0[t-2+i]; pair P lives in sv[2 (P & 7)] (s0), +1 (s1).
        // chain A: s1 terms i = 0,2,4 then s0 terms i = 1,3,5; chain B: s1 terms i = 1,3,5 then s0 terms i = 0,2,4
        float oa = 0.f, ob = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
          oa = fmaf(aa_tap(2 * i), sv[2 * ((j + 7 + i) & 7) + 1], oa);
          ob = fmaf(aa_tap(2 * i + 2), sv[2 * ((j + 8 + i) & 7) + 1], ob);
        }
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
          oa = fmaf(aa_tap(2 * i + 3), sv[2 * ((j + 9 + i) & 7)], oa);
          ob = fmaf(aa_tap(2 * i + 1), sv[2 * ((j + 8 + i) & 7)], ob);
        }
        float o = oa + ob;
        if (EDGE && (t < 0 || t >= len)) o = 0.f;
        drow[j * E] = to_operand<T>(o);
      }
    }
  };
  // Steady 8-step block on packed fp32 pairs (FFMA2: two fp32 lanes per issued instruction, taps as immediates).
  // A pair holds two ADJACENT time steps, XP(m) = (x[m], x[m+1]) with m even -- the layout the 16-byte loads deliver --
  // and every FIR is split by tap parity so that each product reads an aligned pair:
  //   u0[p] = sum_q A0[q] x[p-3+q],  u1[p] = sum_q A1[q] x[p-2+q]          (pair p of the 2x up-sampled signal)
  //   for odd p:  E0(p) = sum_{q even} A0[q] XP(p-3+q)  -> lanes (u0[p], u0[p+1]) partial
  //               O0(p) = sum_{q odd}  A0[q] XP(p-2+q)  -> lanes (u0[p+1], u0[p+2]) partial
  //               (u0[p], u0[p+1]) = (lo E0(p) + hi O0(p-2), hi E0(p) + lo O0(p));  u1 likewise with the parities swapped
  //   out[t] = sum_i tap[2i] s1[t-3+i] + tap[2i+1] s0[t-2+i];  for even t:
  //               A(t) = s1 terms i even + s0 terms i odd -> lanes (out[t], out[t+1]) partial
  //               B(t) = s1 terms i odd + s0 terms i even -> lanes (out[t+1], out[t+2]) partial
  // Every element sees exactly the scalar block's operations in the same order (bit-identical results).
  // State between packed blocks (base % 8 == 0):
  //   XH[m] = XP(base-6+2m) m<3;  OH0/OH1 = O0/O1(base-5);  S1H[m] = S1P(base-9+2m), S0H[m] = S0P(base-9+2m) m<3;  BH = B(base-8)
  uint64_t XH[3], OH0, OH1, S1H[3], S0H[3], BH;
  uint64_t S1Q[7], S0Q[7];                        // S?Q[m] = S?P(base-9+2m) of the block in flight
  auto block8_up = [&](const XRaw<XF16>& xb) {
    uint64_t XQ[7];                               // XQ[m] = XP(base-6+2m)
#pragma unroll
    for (int m = 0; m < 3; ++m) XQ[m] = XH[m];
#pragma unroll
    for (int k = 0; k < 4; ++k) XQ[3 + k] = pk2(xget(xb, 2 * k), xget(xb, 2 * k + 1));
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      S1Q[m] = S1H[m];
      S0Q[m] = S0H[m];
    }
#define TAP2(v) pk2((v), (v))
#pragma unroll
    for (int k = 0; k < 4; ++k) {                 // pairs p = base-3+2k, p+1
      uint64_t e0 = 0ull, o0 = 0ull, e1 = 0ull, o1 = 0ull;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        e0 = ffma2(XQ[k + r], TAP2(2.f * aa_tap(11 - 4 * r)), e0);          // A0[2r]   XP(p-3+2r)
        o0 = ffma2(XQ[k + 1 + r], TAP2(2.f * aa_tap(9 - 4 * r)), o0);       // A0[2r+1] XP(p-1+2r)
        e1 = ffma2(XQ[k + 1 + r], TAP2(2.f * aa_tap(8 - 4 * r)), e1);       // A1[2r+1] XP(p-1+2r)
        o1 = ffma2(XQ[k + 1 + r], TAP2(2.f * aa_tap(10 - 4 * r)), o1);      // A1[2r]   XP(p-1+2r)
      }
      const uint64_t u0 = pk2(lo2(e0) + hi2(OH0), hi2(e0) + lo2(o0));
      const uint64_t u1 = pk2(hi2(OH1) + lo2(e1), lo2(o1) + hi2(e1));
      OH0 = o0;
      OH1 = o1;
#if TB200_SNAKE_SIN == 0
      float a0l, a0h, a1l, a1h;
      upk2(fmul2(u0, pk2(ea, ea)), a0l, a0h);
      upk2(fmul2(u1, pk2(ea, ea)), a1l, a1h);
      const uint64_t z0 = pk2(__sinf(a0l), __sinf(a0h)), z1 = pk2(__sinf(a1l), __sinf(a1h));
      S0Q[3 + k] = ffma2(fmul2(pk2(ib, ib), z0), z0, u0);
      S1Q[3 + k] = ffma2(fmul2(pk2(ib, ib), z1), z1, u1);
#else
      S0Q[3 + k] = pk2(snake_value(lo2(u0), ea, ib), snake_value(hi2(u0), ea, ib));
      S1Q[3 + k] = pk2(snake_value(lo2(u1), ea, ib), snake_value(hi2(u1), ea, ib));
#endif
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) XH[m] = XQ[4 + m];
  };
  T* drow_run = dst;                              // row of out[base - 6] of the packed block in flight
  auto block8_down = [&]() {
    T* drow = drow_run;
    drow_run += 8 * E;
#pragma unroll
    for (int k = 0; k < 4; ++k) {                 // outputs t = base-6+2k, t+1
      uint64_t oa = 0ull, ob = 0ull;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        oa = ffma2(TAP2(aa_tap(4 * r)), S1Q[k + r], oa);                     // tap[2i] s1, i = 2r:   S1P(t-3+2r)
        ob = ffma2(TAP2(aa_tap(4 * r + 2)), S1Q[k + 1 + r], ob);             // tap[2i] s1, i = 2r+1: S1P(t-1+2r)
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        oa = ffma2(TAP2(aa_tap(4 * r + 3)), S0Q[k + 1 + r], oa);             // tap[2i+1] s0, i = 2r+1: S0P(t-1+2r)
        ob = ffma2(TAP2(aa_tap(4 * r + 1)), S0Q[k + 1 + r], ob);             // tap[2i+1] s0, i = 2r:   S0P(t-1+2r)
      }
      const float ol = lo2(oa) + hi2(BH), oh = hi2(oa) + lo2(ob);
      BH = ob;
      if constexpr (sizeof(T) == 2) {
        const uint32_t h = f16x2_sat(ol, oh);
        *reinterpret_cast<unsigned short*>(drow + (2 * k) * E) = (unsigned short)(h & 0xffffu);
        *reinterpret_cast<unsigned short*>(drow + (2 * k + 1) * E) = (unsigned short)(h >> 16);
      } else {
        drow[(2 * k) * E] = to_operand<T>(ol);
        drow[(2 * k + 1) * E] = to_operand<T>(oh);
      }
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      S1H[m] = S1Q[4 + m];
      S0H[m] = S0Q[4 + m];
    }
  };
#undef TAP2
  auto rotate = [&]() {
    cur = n1;
    n1 = n2;
  };
  int base = ts;
  while (base - 6 < t_end) {
    bool steady = false;
    if constexpr (!EDGE) steady = TB200_SNAKE_PACKED && base - 6 >= t_beg && base + 1 < t_end;
    if (steady) {
      // scalar windows -> packed state (base % 8 == 0: xw[i] = x[base-8+i]; pair P lives in sv[2 (P & 7)], +1)
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        XH[m] = pk2(xw[2 + 2 * m], xw[3 + 2 * m]);
        S0H[m] = pk2(sv[2 * ((7 + 2 * m) & 7)], sv[2 * ((2 * m) & 7)]);
        S1H[m] = pk2(sv[2 * ((7 + 2 * m) & 7) + 1], sv[2 * ((2 * m) & 7) + 1]);
      }
      {
        float h0 = 0.f, h1 = 0.f, hb = 0.f;     // the halves of O0/O1(base-5) and B(base-8) that reach into this block
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          h0 = fmaf(xw[3 + 2 * r], 2.f * aa_tap(9 - 4 * r), h0);            // u0[base-3]: odd taps, x[base-5+2r]
          h1 = fmaf(xw[3 + 2 * r], 2.f * aa_tap(10 - 4 * r), h1);           // u1[base-3]: even taps, x[base-5+2r]
          hb = fmaf(aa_tap(4 * r + 2), sv[2 * ((2 * r) & 7) + 1], hb);      // out[base-6]: s1[base-8+2r]
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) hb = fmaf(aa_tap(4 * r + 1), sv[2 * ((2 * r) & 7)], hb);   // s0[base-8+2r]
        OH0 = pk2(0.f, h0);
        OH1 = pk2(0.f, h1);
        BH = pk2(0.f, hb);
      }
      // Two blocks per trip on alternating load buffers: no register rotation, the packed state of one block is
      // produced in place for the next; a buffer is reloaded (two blocks ahead) as soon as the up-sampler has consumed it.
      drow_run = dst + (long long)(base - 6 - t_lo) * E;
      const char* xrow = reinterpret_cast<const char*>(xsrc) + row * (XF16 ? 2 : 4);
      for (;;) {
        if (TB200_SNAKE_PFL1 > 0 && !SMEM) {
          // register look-ahead is short (ptxas sinks the loads to free registers): pull the sectors of the next
          // blocks into L1 instead, which costs no registers; clamped to the utterance
          const int tp = min(base + TB200_SNAKE_PFL1, len - 16);   // the 16 steps of one trip (two blocks)
          prefetch_l1(xrow + (long long)tp * (XF16 ? 2 : 4));
          if (!XF16) prefetch_l1(xrow + (long long)(tp + 8) * 4);
          prefetch_l1(xrow + (long long)(tp + 16) * (XF16 ? 2 : 4) - 1);
        }
        block8_up(cur);
        if (base + 9 >= t_end) {            // last steady block: leave (cur, n1) = (x[base+8..], x[base+16..])
          cur = n1;
          loadx(base + 16, n1);
          block8_down();
          base += 8;
          break;
        }
        loadx(base + 16, cur);
        block8_down();
        base += 8;
        block8_up(n1);
        loadx(base + 16, n1);
        block8_down();
        base += 8;
        if (base + 1 >= t_end) break;
      }
      // packed state -> scalar windows (x[base-6 .. base-1], pairs base-9 .. base-4; older entries are never read again)
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        upk2(XH[m], xw[2 + 2 * m], xw[3 + 2 * m]);
        upk2(S0H[m], sv[2 * ((7 + 2 * m) & 7)], sv[2 * ((2 * m) & 7)]);
        upk2(S1H[m], sv[2 * ((7 + 2 * m) & 7) + 1], sv[2 * ((2 * m) & 7) + 1]);
      }
    } else {
      loadx(base + 16, n2);
      if (!TB200_SNAKE_PACKED && !EDGE && base - 6 >= t_beg && base + 1 < t_end) block8(std::false_type{}, base);
      else block8(std::true_type{}, base);
      rotate();
      base += 8;
    }
  }
}


}  // namespace tb200
