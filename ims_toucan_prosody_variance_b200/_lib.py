"""ctypes binding of libtoucan_b200.so (the C ABI declared in include/toucan_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TB200_LIB", os.path.join(_HERE, "libtoucan_b200.so"))  # TB200_LIB: tuning builds only

F32, F16 = 0, 1
PREC_FP32_SIMT, PREC_F16, PREC_TF32 = 0, 1, 2
ACT_NONE, ACT_LEAKY_RELU, ACT_AA_SNAKEBETA, ACT_RELU, ACT_SWISH, ACT_TANH = 0, 1, 2, 3, 4, 5
OUT_NONE, OUT_TANH, OUT_RELU = 0, 1, 2

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class Conv1dParams(ctypes.Structure):
    """struct tb200_conv1d_params (include/toucan_b200.h)."""
    _fields_ = [
        ("x", c_void_p), ("x_dtype", c_int), ("x_bs", c_int64), ("x_ld", c_int), ("len_in", c_void_p),
        ("B", c_int), ("C_in", c_int), ("L_in_max", c_int),
        ("C_out", c_int), ("K", c_int), ("dilation", c_int), ("pad", c_int), ("transposed_stride", c_int),
        ("w_packed", c_void_p), ("bias", c_void_p), ("precision", c_int),
        ("act", c_int), ("act_slope", c_float), ("act_alpha", c_void_p), ("act_beta", c_void_p),
        ("out_act", c_int), ("out_alpha", c_float), ("residual", c_void_p), ("r_dtype", c_int), ("r_bs", c_int64), ("r_ld", c_int),
        ("res_beta", c_float), ("accumulate", c_int),
        ("y", c_void_p), ("y_dtype", c_int), ("y_bs", c_int64), ("y_ld", c_int),
    ]


class RespairParams(ctypes.Structure):
    """struct tb200_respair_params (include/toucan_b200.h)."""
    _fields_ = [
        ("x", c_void_p), ("x_dtype", c_int), ("x_bs", c_int64), ("x_ld", c_int), ("len", c_void_p),
        ("B", c_int), ("C", c_int), ("L_max", c_int), ("K", c_int), ("dilation", c_int),
        ("w1_packed", c_void_p), ("bias1", c_void_p), ("w2_packed", c_void_p), ("bias2", c_void_p),
        ("act", c_int), ("act_slope", c_float),
        ("act1_alpha", c_void_p), ("act1_beta", c_void_p), ("act2_alpha", c_void_p), ("act2_beta", c_void_p),
        ("out_alpha", c_float), ("res_beta", c_float), ("accumulate", c_int),
        ("y", c_void_p), ("y_dtype", c_int), ("y_bs", c_int64), ("y_ld", c_int),
    ]


# name -> (restype, argtypes); must list every symbol include/toucan_b200.h declares
SIGNATURES = {
    "tb200_version": (c_int, []),
    "tb200_last_error": (ctypes.c_char_p, []),
    "tb200_sm_count": (c_int, []),
    "tb200_conv1d": (c_int, [ctypes.POINTER(Conv1dParams), c_void_p]),
    "tb200_conv1d_staged": (c_int, [ctypes.POINTER(Conv1dParams), c_void_p]),
    "tb200_respair": (c_int, [ctypes.POINTER(RespairParams), c_void_p]),
    "tb200_respair_trace_read": (c_int, [c_void_p, c_int]),
    "tb200_packed_weight_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "tb200_pack_conv_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tb200_duration_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_variance_edit": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "tb200_length_regulate": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                      c_int64, c_int, c_void_p, c_int, c_void_p]),
    "tb200_debug_trace_read": (c_int, [c_void_p, c_int]),
    "tb200_channel_norm": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "tb200_group_norm": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p,
                                 c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p]),
    "tb200_glu_dwconv": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "tb200_relpos_attention": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "tb200_relpos_attention_tc": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                          c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "tb200_rowvec_affine": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_int64, c_float, c_void_p]),
    "tb200_transpose": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                c_void_p]),
    "tb200_squeeze2": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_int,
                               c_void_p]),
    "tb200_wn_gate": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tb200_flow_close": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "tb200_l2_normalize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "tb200_cln_mlp": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load():
    """Load the library (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                              f"(ims_toucan_prosody_variance_b200/csrc/build.sh); there is no fallback path")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().tb200_last_error().decode(errors="replace")
        raise EngineError(f"{what} failed with code {rc}: {msg}")


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
