"""CPU simulation: what would the anti-aliasing FIRs of BigVGAN's SnakeBeta cost in SNR on the tensor cores?

Not a test (pytest does not collect it): `python tests/sim_fir_precision.py [frames]` prints the waveform SNR against the
fp32 oracle (oracle/restate.py::bigvgan_forward) of variants of the oracle in which the conv operands are rounded to
fp16 (what the engine's tcgen05 convs do) and the 12-tap kaiser-sinc up/down FIRs of alias_free_torch.Activation1d
additionally run with the arithmetic a tensor-core formulation would have:

  base        conv operands fp16 (activations after the snake, weights), FIRs in fp32          = the shipped engine
  sig16       + the FIR *inputs* rounded to fp16 (x before the up-FIR, snake output before the down-FIR)
  taps16      + the FIR taps rounded to fp16 as well (one-MMA Toeplitz formulation)
  taps_hilo   + taps split into fp16 hi + fp16 lo (two MMAs per FIR)
  up16 / dn16   taps16 on the up-FIR only / the down-FIR only

DESIGN.md section 5.1 quotes the output.  Uses only oracle code and torch CPU ops.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import factory, restate  # noqa: E402


def h(t):
    return t.half().float()


def make_aa(sig16, up_taps, dn_taps):
    """up_taps / dn_taps: 'f32' | 'f16' | 'hilo'."""
    def taps(kind):
        f = restate.aa_filter()
        if kind == "f16":
            return [h(f)]
        if kind == "hilo":
            return [h(f), h(f - h(f))]
        return [f]

    def aa(x, alpha, beta):
        c = x.shape[1]
        xin = h(x) if sig16 else x
        y = F.pad(xin, (5, 5), mode="replicate")
        y = sum(2 * F.conv_transpose1d(y, f.expand(c, -1, -1), stride=2, groups=c)[..., 15:-15] for f in taps(up_taps))
        y = restate.snake_beta(y, alpha, beta)
        if sig16:
            y = h(y)
        y = F.pad(y, (5, 6), mode="replicate")
        return sum(F.conv1d(y, f.expand(c, -1, -1), stride=2, groups=c) for f in taps(dn_taps))
    return aa


def run(sd, mel, aa):
    """bigvgan_forward with fp16-rounded conv operands and the given activation."""
    keep_aa, keep_c, keep_t = restate.aa_snake, F.conv1d, F.conv_transpose1d
    in_aa = [False]

    def conv(x, w, b=None, **kw):
        if in_aa[0] or kw.get("groups", 1) != 1:
            return keep_c(x, w, b, **kw)
        return keep_c(h(x), h(w), b, **kw)

    def convt(x, w, b=None, **kw):
        if in_aa[0] or kw.get("groups", 1) != 1:
            return keep_t(x, w, b, **kw)
        return keep_t(h(x), h(w), b, **kw)

    def aa_wrapped(x, a, b):
        in_aa[0] = True
        try:
            return aa(x, a, b)
        finally:
            in_aa[0] = False
    restate.aa_snake, F.conv1d, F.conv_transpose1d = aa_wrapped, conv, convt
    try:
        return restate.bigvgan_forward(sd, mel)
    finally:
        restate.aa_snake, F.conv1d, F.conv_transpose1d = keep_aa, keep_c, keep_t


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    torch.manual_seed(0)
    sd = restate.fold_weight_norm(factory.make_state_dict("bigvgan"))
    mel = factory.make_mel(1, frames, seed=7)
    with torch.no_grad():
        ref = restate.bigvgan_forward(sd, mel)
        variants = [("base", (False, "f32", "f32")), ("sig16", (True, "f32", "f32")), ("taps16", (True, "f16", "f16")),
                    ("taps_hilo", (True, "hilo", "hilo")), ("up16", (True, "f16", "f32")), ("dn16", (True, "f32", "f16"))]
        for name, cfg in variants:
            out = run(sd, mel, make_aa(*cfg))
            print(f"{name:10s} SNR {restate.snr_db(out, ref):6.2f} dB", flush=True)


if __name__ == "__main__":
    main()
