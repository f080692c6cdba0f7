"""Per-launch CUDA-event timing of one generator forward: where does the step go?
    python tools/profile_vocoder.py [bigvgan|hifigan] [batch] [frames] [precision] [activations f32|f16] [fuse 1|0]
Prints one row per launch (tb200_conv1d, or tb200_respair = "pair"): shape, ms, achieved TFLOP/s and algorithmic GB/s."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import build_generator  # noqa: E402
from ims_toucan_prosody_variance_b200 import ops  # noqa: E402
from oracle import factory  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "bigvgan"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 500
prec = sys.argv[4] if len(sys.argv) > 4 else "f16"
acts = sys.argv[5] if len(sys.argv) > 5 else "f32"
fuse = (sys.argv[6] != "0") if len(sys.argv) > 6 else True
dev = torch.device("cuda:0")
model, _ = build_generator(kind, prec, dev, activation_dtype=acts)
if not fuse:
    model.fuse_pairs = False
    model.remove_weight_norm()
mel = factory.make_mel(batch, frames, seed=1).to(dev)
lengths = torch.full((batch,), frames, dtype=torch.int32)
for _ in range(2):
    model.forward_batch(mel, lengths)
torch.cuda.synchronize()

records = []
orig = ops.ConvLayer.__call__


def timed(self, x, lengths, out, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig(self, x, lengths, out, **kw)
    e1.record()
    records.append((self, x, out, kw, e0, e1))
    return r


orig_pair = ops.ResPair.__call__


def timed_pair(self, x, lengths, out, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = orig_pair(self, x, lengths, out, **kw)
    e1.record()
    records.append((self, x, out, kw, e0, e1))
    return r


ops.ConvLayer.__call__ = timed
ops.ResPair.__call__ = timed_pair
model.forward_batch(mel, lengths)
torch.cuda.synchronize()
ops.ConvLayer.__call__ = orig
ops.ResPair.__call__ = orig_pair
total = 0.0
print(f"{'Cin':>4} {'Cout':>4} {'K':>3} {'dil':>3} {'up':>2} {'L':>7} {'act':>3} {'ms':>8} {'TFLOP/s':>8} {'GB/s':>7}")
agg = {}
for layer, x, out, kw, e0, e1 in records:
    ms = e0.elapsed_time(e1)
    total += ms
    if isinstance(layer, ops.ResPair):
        c1 = layer.c1
        L = kw.get("l_max", x.shape[2])
        flops = 4.0 * x.shape[0] * L * c1.c_in * c1.c_out * c1.k
        byts = x.shape[0] * L * c1.c_in * (x.element_size() + out.element_size() * (2 if kw.get("accumulate") else 1))
        print(f"{c1.c_in:4d} {c1.c_out:4d} {c1.k:3d} {c1.dilation:3d} pr {L:7d} {'sn' if layer.act1 is not None else 'lk':>3} "
              f"{ms:8.3f} {flops / ms / 1e9:8.1f} {byts / ms / 1e6:7.0f}")
        a = agg.setdefault((c1.c_in, c1.c_out, "pair"), [0.0, 0.0, 0.0])
        a[0] += ms; a[1] += flops; a[2] += byts
        continue
    L = kw.get("l_in_max", x.shape[2])
    Lout = L * layer.up if layer.up else L
    taps = 2 if layer.up else layer.k
    flops = 2.0 * x.shape[0] * L * layer.c_in * (layer.c_out * (layer.up or 1)) * taps
    byts = x.shape[0] * (L * layer.c_in * x.element_size() + Lout * layer.c_out * out.element_size())
    if kw.get("residual") is not None:
        byts += x.shape[0] * Lout * layer.c_out * 4
    if kw.get("accumulate"):
        byts += x.shape[0] * Lout * layer.c_out * 4
    print(f"{layer.c_in:4d} {layer.c_out:4d} {layer.k:3d} {layer.dilation:3d} {layer.up:2d} {L:7d} {kw.get('act', 0):3d} "
          f"{ms:8.3f} {flops / ms / 1e9:8.1f} {byts / ms / 1e6:7.0f}")
    key = (layer.c_in, layer.c_out, layer.up)
    a = agg.setdefault(key, [0.0, 0.0, 0.0])
    a[0] += ms; a[1] += flops; a[2] += byts
print(f"total {total:.3f} ms over {len(records)} launches")
for key, (ms, fl, by) in agg.items():
    print(f"  Cin={key[0]} Cout={key[1]} up={key[2]}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  {by / ms / 1e6:7.0f} GB/s")
