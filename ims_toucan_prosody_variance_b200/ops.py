"""Host-side wrappers over the C ABI: torch tensors in, raw pointers out.

PyTorch is plumbing here (device memory, streams); all arithmetic on the hot path happens in
libtoucan_b200.so.  Activations are NCL (B, C, L) with per-utterance int32 lengths on device.
"""
import ctypes

import torch

from . import _lib
from ._lib import (ACT_AA_SNAKEBETA, ACT_LEAKY_RELU, ACT_NONE, ACT_RELU, ACT_SWISH, ACT_TANH, F16, F32, OUT_NONE,
                   OUT_RELU, OUT_TANH, PREC_F16, PREC_FP32_SIMT, PREC_TF32)

PRECISIONS = {"fp32": PREC_FP32_SIMT, "f16": PREC_F16, "tf32": PREC_TF32}

LAUNCHES = 0  # kernels of ours enqueued since import (bench.py reports the per-step delta)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise _lib.EngineError(f"unsupported activation dtype {t.dtype}")


def _on_device(fn):
    """Run an ABI wrapper on the device its tensors live on: the kernels launch on the CURRENT device and stream, so a
    model on cuda:1 called while cuda:0 is current would otherwise launch on GPU 0 with GPU-1 pointers.  All tensor
    arguments must share one device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            for t in (a if isinstance(a, (tuple, list)) else (a,)):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    if dev is None:
                        dev = t.device
                    elif t.device != dev:
                        raise _lib.EngineError(f"toucan_b200: tensors on different devices ({dev} and {t.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.EngineError("toucan_b200 has no CPU path: tensors must live on a CUDA device")


class ConvLayer:
    """One Conv1d / ConvTranspose1d / Linear with its weights packed for tb200_conv1d.

    weight: torch layout, fp32 -- (C_out, C_in, K), or (C_in, C_out, 2u) when transposed_stride=u.
    """

    @_on_device
    def __init__(self, weight, bias=None, dilation=1, padding=0, transposed_stride=0, precision="f16"):
        lib = _lib.load()
        _require_cuda(weight, bias)
        weight = weight.detach().to(torch.float32).contiguous()
        if weight.dim() == 2:
            weight = weight.unsqueeze(-1)
        self.up = int(transposed_stride)
        if self.up:
            self.c_in, self.c_out, self.k = weight.shape
        else:
            self.c_out, self.c_in, self.k = weight.shape
        self.dilation, self.pad = int(dilation), int(padding)
        self.precision = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        nbytes = lib.tb200_packed_weight_bytes(self.c_in, self.c_out, self.k, self.up, self.precision)
        self.packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
        _lib.check(lib.tb200_pack_conv_weight(_ptr(weight), _ptr(self.packed), self.c_in, self.c_out, self.k, self.up,
                                              self.precision, _lib.stream_ptr()), "tb200_pack_conv_weight")
        self.bias = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        self._p = _lib.Conv1dParams()

    def out_len(self, length):
        return length * self.up if self.up else length

    @_on_device
    def staged_ok(self, x, out, residual, act, out_act):
        """Can this call run through tb200_conv1d_staged (TMA-fed tiles; see include/toucan_b200.h)?"""
        def al(t):
            unit = 16 // t.element_size()
            return t.data_ptr() % 16 == 0 and t.stride(1) % unit == 0 and t.stride(0) % unit == 0
        return (self.precision == PREC_F16 and self.up == 0 and self.c_in == self.c_out and self.c_in in (32, 64, 128)
                and self.k % 2 == 1 and self.pad == (self.k - 1) // 2 * self.dilation and out_act == OUT_NONE
                and act in (ACT_LEAKY_RELU, ACT_AA_SNAKEBETA) and al(x) and al(out) and (residual is None or al(residual))
                and x.data_ptr() != out.data_ptr())

    def __call__(self, x, lengths, out, l_in_max=None, act=ACT_NONE, slope=0.0, alpha=None, beta=None, out_act=OUT_NONE,
                 out_alpha=1.0, residual=None, res_beta=1.0, accumulate=False, staged=False):
        """x (B,C_in,L) -> out (B,C_out,L_out), both NCL with contiguous rows.  lengths: int32 (B) or None.
        staged=True: use tb200_conv1d_staged when the call qualifies (same result, TMA-fed pipeline)."""
        global LAUNCHES
        _require_cuda(x, out, residual, lengths)
        if x.stride(2) != 1 or out.stride(2) != 1 or x.shape[1] != self.c_in or out.shape[1] != self.c_out:
            raise _lib.EngineError(f"conv1d: bad tensor layout x={tuple(x.shape)} out={tuple(out.shape)} "
                                   f"expected C_in={self.c_in} C_out={self.c_out}")
        p = self._p
        p.x, p.x_dtype, p.x_bs, p.x_ld = x.data_ptr(), _dtype_code(x), x.stride(0), x.stride(1)
        p.len_in = lengths.data_ptr() if lengths is not None else None
        p.B, p.C_in, p.L_in_max = x.shape[0], self.c_in, int(l_in_max if l_in_max is not None else x.shape[2])
        p.C_out, p.K, p.dilation, p.pad, p.transposed_stride = self.c_out, self.k, self.dilation, self.pad, self.up
        p.w_packed, p.bias, p.precision = self.packed.data_ptr(), (self.bias.data_ptr() if self.bias is not None else None), self.precision
        p.act, p.act_slope = act, slope
        p.act_alpha = alpha.data_ptr() if alpha is not None else None
        p.act_beta = beta.data_ptr() if beta is not None else None
        p.out_act, p.out_alpha = out_act, out_alpha
        if residual is not None:
            if residual.stride(2) != 1:
                raise _lib.EngineError("conv1d: residual must be NCL with contiguous rows")
            p.residual, p.r_dtype, p.r_bs, p.r_ld = residual.data_ptr(), _dtype_code(residual), residual.stride(0), residual.stride(1)
        else:
            p.residual, p.r_dtype, p.r_bs, p.r_ld = None, F32, 0, 0
        p.res_beta, p.accumulate = res_beta, int(accumulate)
        p.y, p.y_dtype, p.y_bs, p.y_ld = out.data_ptr(), _dtype_code(out), out.stride(0), out.stride(1)
        if self.out_len(p.L_in_max) > out.shape[2]:
            raise _lib.EngineError("conv1d: output buffer too short")
        if staged and self.staged_ok(x, out, residual, act, out_act) and p.L_in_max <= min(x.stride(1), out.stride(1)):
            _lib.check(_lib.load().tb200_conv1d_staged(ctypes.byref(p), _lib.stream_ptr()), "tb200_conv1d_staged")
        else:
            _lib.check(_lib.load().tb200_conv1d(ctypes.byref(p), _lib.stream_ptr()), "tb200_conv1d")
        LAUNCHES += 1
        return out


class ResPair:
    """One residual pair  x + conv2(ACT2(conv1(ACT1(x))))  as a single tb200_respair launch
    (AMP.py:51-60, ResidualBlock.py:93-97).  conv1 / conv2: ConvLayer objects packed with precision 'f16'
    (same channel count and kernel size; conv2 has dilation 1).  act1 / act2: (alpha, beta) tensors for the
    anti-aliased SnakeBeta, or None for LeakyReLU."""

    def __init__(self, conv1, conv2, act1=None, act2=None):
        if conv1.precision != PREC_F16 or conv2.precision != PREC_F16:
            raise _lib.EngineError("ResPair needs fp16-operand ConvLayers")
        if not (conv1.c_in == conv1.c_out == conv2.c_in == conv2.c_out and conv1.k == conv2.k and conv2.dilation == 1
                and conv1.up == 0 and conv2.up == 0 and conv1.pad == (conv1.k - 1) // 2 * conv1.dilation
                and conv2.pad == (conv2.k - 1) // 2):
            raise _lib.EngineError("ResPair: the two convolutions must be 'same'-padded C->C with one kernel size")
        self.c1, self.c2, self.act1, self.act2 = conv1, conv2, act1, act2
        self._p = _lib.RespairParams()

    @staticmethod
    def supported(channels, precision):
        return precision in ("f16", PREC_F16) and channels in (32, 64, 128)

    @_on_device
    def __call__(self, x, lengths, out, l_max=None, slope=0.1, out_alpha=1.0, res_beta=1.0, accumulate=False):
        """x (B,C,L) -> out (B,C,L), NCL with contiguous rows, fp32 or fp16; out must not alias x."""
        global LAUNCHES
        _require_cuda(x, out, lengths)
        c = self.c1.c_in
        if x.stride(2) != 1 or out.stride(2) != 1 or x.shape[1] != c or out.shape[1] != c:
            raise _lib.EngineError(f"respair: bad tensor layout x={tuple(x.shape)} out={tuple(out.shape)} C={c}")
        p = self._p
        p.x, p.x_dtype, p.x_bs, p.x_ld = x.data_ptr(), _dtype_code(x), x.stride(0), x.stride(1)
        p.len = lengths.data_ptr() if lengths is not None else None
        p.B, p.C, p.L_max = x.shape[0], c, int(l_max if l_max is not None else x.shape[2])
        p.K, p.dilation = self.c1.k, self.c1.dilation
        p.w1_packed, p.bias1 = self.c1.packed.data_ptr(), (self.c1.bias.data_ptr() if self.c1.bias is not None else None)
        p.w2_packed, p.bias2 = self.c2.packed.data_ptr(), (self.c2.bias.data_ptr() if self.c2.bias is not None else None)
        if self.act1 is not None:
            p.act, p.act_slope = ACT_AA_SNAKEBETA, 0.0
            p.act1_alpha, p.act1_beta = self.act1[0].data_ptr(), self.act1[1].data_ptr()
            p.act2_alpha, p.act2_beta = self.act2[0].data_ptr(), self.act2[1].data_ptr()
        else:
            p.act, p.act_slope = ACT_LEAKY_RELU, slope
            p.act1_alpha = p.act1_beta = p.act2_alpha = p.act2_beta = None
        p.out_alpha, p.res_beta, p.accumulate = out_alpha, res_beta, int(accumulate)
        p.y, p.y_dtype, p.y_bs, p.y_ld = out.data_ptr(), _dtype_code(out), out.stride(0), out.stride(1)
        if p.L_max > out.shape[2] or p.L_max > x.shape[2]:
            raise _lib.EngineError("respair: buffers shorter than L_max")
        _lib.check(_lib.load().tb200_respair(ctypes.byref(p), _lib.stream_ptr()), "tb200_respair")
        LAUNCHES += 1
        return out


@_on_device
def duration_finalize(text, text_len, log_dur=None, gold_dur=None, pause_scale=1.0, duration_scale=1.0, t_ld=None):
    """text (B,T,62) fp32, text_len (B) int32, log_dur (B,T_ld) fp32 or gold_dur (B,T_ld) int64 (T_ld >= T, row pitch).
    Returns durations (B,T_ld) int64, inclusive prefix sums (B,T_ld) int32, frames (B) int32."""
    global LAUNCHES
    _require_cuda(text, text_len, log_dur, gold_dur)
    b, t, _ = text.shape
    src = log_dur if log_dur is not None else gold_dur
    src = src.contiguous()
    t_ld = src.shape[1]
    text = text.contiguous()
    dur = torch.empty((b, t_ld), dtype=torch.int64, device=text.device)
    cum = torch.empty((b, t_ld), dtype=torch.int32, device=text.device)
    frames = torch.empty((b,), dtype=torch.int32, device=text.device)
    _lib.check(_lib.load().tb200_duration_finalize(
        _ptr(src) if log_dur is not None else None, _ptr(src) if log_dur is None else None, _ptr(text), _ptr(text_len), b, t, t_ld,
        float(pause_scale), float(duration_scale), _ptr(dur), _ptr(cum), _ptr(frames), _lib.stream_ptr()),
        "tb200_duration_finalize")
    LAUNCHES += 1
    return dur, cum, frames


@_on_device
def variance_edit(curve, text, text_len, which, variance_scale=1.0, t_ld=None):
    """In-place pitch (which=0) / energy (which=1) edits + variance scaling on curve (B,T_ld) fp32."""
    global LAUNCHES
    _require_cuda(curve, text, text_len)
    b, t, _ = text.shape
    if not curve.is_contiguous() or curve.shape[0] != b or curve.shape[1] < t:
        raise _lib.EngineError("variance_edit: curve must be contiguous (B,T_ld)")
    _lib.check(_lib.load().tb200_variance_edit(_ptr(curve), _ptr(text.contiguous()), _ptr(text_len), b, t, curve.shape[1],
                                               int(which), float(variance_scale), _lib.stream_ptr()), "tb200_variance_edit")
    LAUNCHES += 1
    return curve


@_on_device
def length_regulate(enc, cum, text_len, frames, f_max, pitch=None, energy=None, wp=None, bp=None, we=None, be=None,
                    out=None, want_index=False):
    """enc (B,C,T) NCL fp32 -> (B,C,F_max) NCL fp32 (only f < frames[b] written)."""
    global LAUNCHES
    _require_cuda(enc, cum, text_len, frames, pitch, energy)
    b, c, _ = enc.shape
    f_ld = (f_max + 3) // 4 * 4
    if out is None:
        out = torch.zeros((b, c, f_ld), dtype=torch.float32, device=enc.device)
    f2p = torch.full((b, f_ld), -1, dtype=torch.int32, device=enc.device) if want_index else None
    _lib.check(_lib.load().tb200_length_regulate(
        _ptr(enc), enc.stride(0), enc.stride(1), _ptr(pitch), _ptr(energy), pitch.stride(0) if pitch is not None else 0,
        _ptr(wp), _ptr(bp), _ptr(we), _ptr(be), _ptr(cum), cum.stride(0), _ptr(text_len), _ptr(frames), b, c, int(f_max),
        _ptr(out), out.stride(0), out.stride(1), _ptr(f2p), f_ld, _lib.stream_ptr()), "tb200_length_regulate")
    LAUNCHES += 1
    return (out, f2p) if want_index else out


# ---------------------------------------------------------------------------------------------
# acoustic-model kernels (csrc/acoustic.cu).  All tensors fp32 NCL (B,C,L) with unit time stride.
# ---------------------------------------------------------------------------------------------

def _ncl(t):
    if t.dtype != torch.float32 or t.dim() != 3 or t.stride(2) != 1:
        raise _lib.EngineError(f"expected an fp32 NCL tensor with contiguous rows, got {t.dtype} {tuple(t.shape)} {t.stride()}")
    return ctypes.c_void_p(t.data_ptr()), t.stride(0), t.stride(1)


def _call(name, *args):
    global LAUNCHES
    _lib.check(getattr(_lib.load(), name)(*args, _lib.stream_ptr()), name)
    LAUNCHES += 1


@_on_device
def channel_norm(x, lengths, out, gamma, beta, l_max, conditional=False, eps=1e-12):
    """LayerNorm over channels (conditional=False) or ConditionalLayerNorm with per-utterance (B,C) gamma/beta."""
    _require_cuda(x, out, gamma, beta, lengths)
    b, c, _ = x.shape
    gb_bs = gamma.stride(0) if conditional else 0
    _call("tb200_channel_norm", *_ncl(x), *_ncl(out), _ptr(lengths), b, c, int(l_max), _ptr(gamma), _ptr(beta), gb_bs,
          1 if conditional else 0, float(eps))
    return out


@_on_device
def group_norm(x, lengths, out, gamma, beta, groups, l_max, residual=None, tanh=False, eps=1e-5):
    _require_cuda(x, out, gamma, beta, lengths, residual)
    b, c, _ = x.shape
    rp = _ncl(residual) if residual is not None else (None, 0, 0)
    _call("tb200_group_norm", *_ncl(x), *_ncl(out), *rp, _ptr(lengths), b, c, int(l_max), int(groups), _ptr(gamma), _ptr(beta),
          float(eps), OUT_TANH if tanh else OUT_NONE)
    return out


@_on_device
def glu_dwconv(x, lengths, out, w, bias, bn_mean, bn_var, bn_gamma, bn_beta, l_max, bn_eps=1e-5):
    _require_cuda(x, out, w, bias, lengths)
    b, c, _ = out.shape
    _call("tb200_glu_dwconv", *_ncl(x), *_ncl(out), _ptr(lengths), b, c, int(l_max), _ptr(w), _ptr(bias), int(w.shape[-1]),
          _ptr(bn_mean), _ptr(bn_var), _ptr(bn_gamma), _ptr(bn_beta), float(bn_eps))
    return out


@_on_device
def relpos_attention(qkv, lengths, out, pos, pos_center, bias_u, bias_v, heads, l_max):
    """fp32 CUDA-core kernel.  pos (D, cols): column (pos_center - r) holds relative position r."""
    _require_cuda(qkv, out, pos, bias_u, bias_v, lengths)
    b, c3, _ = qkv.shape
    dk = c3 // 3 // heads
    _call("tb200_relpos_attention", *_ncl(qkv), _ptr(pos), pos.stride(0), int(pos_center), pos.shape[1], _ptr(bias_u),
          _ptr(bias_v), _ptr(lengths), b, int(heads), dk, int(l_max), *_ncl(out))
    return out


ATTN_BAND_PAD = 256   # zero rows on both sides of the packed positional table (a tile reads a 256-row band)


def pack_relpos_table(pos, pos_center, heads):
    """(D, cols) fp32 table with column (pos_center - r) = relative position r  ->  the fp16 operand rows of
    tb200_relpos_attention_tc: (H, dk/8, rows, 8), row (center + r) = relative position r, ATTN_BAND_PAD zero rows on
    both sides.  Returns (packed, center).  Done once per layer and table size (plain tensor ops, load time)."""
    d, cols = pos.shape
    dk = d // heads
    r_lo, r_hi = pos_center - (cols - 1), pos_center            # relative positions covered: r_lo .. r_hi
    inc = torch.flip(pos[:, :cols], dims=[1])                   # column i <-> r = r_lo + i
    packed = torch.zeros((heads, dk // 8, cols + 2 * ATTN_BAND_PAD, 8), dtype=torch.float16, device=pos.device)
    packed[:, :, ATTN_BAND_PAD:ATTN_BAND_PAD + cols] = inc.reshape(heads, dk // 8, 8, cols).permute(0, 1, 3, 2).to(torch.float16)
    return packed.contiguous(), ATTN_BAND_PAD - r_lo


@_on_device
def relpos_attention_tc(qkv, lengths, out, pos16, pos_center, bias_u, bias_v, heads, l_max):
    """tcgen05 flash attention with fp16 operands (tf32 / f16 precision modes).  pos16, pos_center from pack_relpos_table."""
    _require_cuda(qkv, out, pos16, bias_u, bias_v, lengths)
    b, c3, _ = qkv.shape
    dk = c3 // 3 // heads
    _call("tb200_relpos_attention_tc", *_ncl(qkv), _ptr(pos16), pos16.shape[2], int(pos_center), _ptr(bias_u), _ptr(bias_v),
          _ptr(lengths), b, int(heads), dk, int(l_max), *_ncl(out))
    return out


@_on_device
def rowvec_affine(x, lengths, out, l_max, vec=None, scale=1.0):
    _require_cuda(x, out, vec, lengths)
    b, c, _ = out.shape
    xp = _ncl(x) if x is not None else (None, 0, 0)
    _call("tb200_rowvec_affine", *xp, *_ncl(out), _ptr(lengths), b, c, int(l_max), _ptr(vec),
          vec.stride(0) if vec is not None else 0, float(scale))
    return out


@_on_device
def to_ncl(x_blc, lengths, out, l_max):
    """(B,L,C) -> NCL (B,C,L)."""
    _require_cuda(x_blc, out, lengths)
    b, _, c = x_blc.shape
    _call("tb200_transpose", _ptr(x_blc), x_blc.stride(0), x_blc.stride(1), *_ncl(out), _ptr(lengths), b, c, int(l_max), 1)
    return out


@_on_device
def from_ncl(x, lengths, out_blc, l_max):
    """NCL (B,C,L) -> (B,L,C)."""
    _require_cuda(x, out_blc, lengths)
    b, c, _ = x.shape
    _call("tb200_transpose", *_ncl(x), _ptr(out_blc), out_blc.stride(0), out_blc.stride(1), _ptr(lengths), b, c, int(l_max), 0)
    return out_blc


@_on_device
def squeeze2(x, lengths, out, l_max, inverse=False):
    """Glow squeeze (inverse=False: lengths/l_max unsqueezed) or unsqueeze (inverse=True: lengths/l_max squeezed).
    The channel count passed to the kernel is always the UNSQUEEZED one."""
    _require_cuda(x, out, lengths)
    b = x.shape[0]
    c = out.shape[1] if inverse else x.shape[1]
    _call("tb200_squeeze2", *_ncl(x), *_ncl(out), _ptr(lengths), b, c, int(l_max), 1 if inverse else 0)
    return out


@_on_device
def wn_gate(a, lengths, out, l_max):
    _require_cuda(a, out, lengths)
    b, h, _ = out.shape
    _call("tb200_wn_gate", *_ncl(a), *_ncl(out), _ptr(lengths), b, h, int(l_max))
    return out


@_on_device
def flow_close(x, ml, lengths, l_max, w_inv, an_bias, an_logs):
    _require_cuda(x, ml, lengths, w_inv, an_bias, an_logs)
    b, c, _ = x.shape
    _call("tb200_flow_close", *_ncl(x), *_ncl(ml), _ptr(lengths), b, c, int(l_max), _ptr(w_inv), _ptr(an_bias), _ptr(an_logs))
    return x


@_on_device
def l2_normalize(x):
    _require_cuda(x)
    x = x.contiguous().float()
    y = torch.empty_like(x)
    _call("tb200_l2_normalize", _ptr(x), _ptr(y), x.shape[0], x.shape[1])
    return y


@_on_device
def cln_mlp(e, w0, b0, w2, b2, w4, b4):
    """e (B,E); stacked MLP weights (N,...) -> (N,B,Cc)."""
    _require_cuda(e, w0, b0, w2, b2, w4, b4)
    n, cc = w4.shape[0], w4.shape[1]
    out = torch.empty((n, e.shape[0], cc), dtype=torch.float32, device=e.device)
    _call("tb200_cln_mlp", _ptr(e), e.shape[0], e.shape[1], cc, n, _ptr(w0), _ptr(b0), _ptr(w2), _ptr(b2), _ptr(w4), _ptr(b4),
          _ptr(out))
    return out
