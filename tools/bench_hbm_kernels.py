"""Achieved HBM bandwidth of the memory-bound kernels (K7 expand and the element-wise / normalisation kernels of the
acoustic model) at sizes well above the 126 MB L2: python tools/bench_hbm_kernels.py
Algorithmic bytes per DESIGN.md section 3.3; peak = MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ims_toucan_prosody_variance_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
peak = 6547.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print(f"{name:34s} {nbytes / 1e6:9.1f} MB  {ms:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of measured {peak:.0f} GB/s")


g = torch.Generator().manual_seed(0)
# ---- K7: duration finalize + length regulate (expand) : B=512 utterances, 200 phonemes, ~5 frames each
B, T, C = 512, 200, 192
text = (torch.rand(B, T, 62, generator=g) < 0.3).float().to(dev)
text[:, :, 21] = 0.0                                                  # no word boundaries: every phoneme keeps its duration
tlen = torch.full((B,), T, dtype=torch.int32, device=dev)
logd = (torch.rand(B, T, generator=g) * 0.4 + 1.6).to(dev)            # durations 4..6
enc = torch.randn(B, C, T, generator=g).to(dev)
dur, cum, frames = ops.duration_finalize(text, tlen, log_dur=logd)
fmax = int(frames.max())
ftot = int(frames.sum())
pitch = torch.rand(B, T, generator=g).to(dev)
wp, bp = torch.randn(C, device=dev), torch.randn(C, device=dev)
out = torch.zeros(B, C, (fmax + 3) // 4 * 4, device=dev)
ms = timeit(lambda: ops.length_regulate(enc, cum, tlen, frames, fmax, pitch=pitch, energy=pitch, wp=wp, bp=bp, we=wp, be=bp, out=out))
report(f"length_regulate (F={ftot})", 4 * C * (B * T + ftot) + 8 * B * T, ms)
ms = timeit(lambda: ops.duration_finalize(text, tlen, log_dur=logd))
report("duration_finalize", B * T * (62 * 4 + 4 + 8 + 4), ms)

# ---- element-wise / normalisation kernels on a (B, C, L) activation of ~400 MB
B, L = 256, 2000
lens = torch.full((B,), L, dtype=torch.int32, device=dev)
x = torch.randn(B, 192, L, device=dev)
y = torch.empty_like(x)
gam, bet = torch.ones(192, device=dev), torch.zeros(192, device=dev)
ms = timeit(lambda: ops.channel_norm(x, lens, y, gam, bet, L))
report("channel_norm (LayerNorm, C=192)", 2 * x.numel() * 4, ms)
a = torch.randn(B, 384, L, device=dev)
ms = timeit(lambda: ops.wn_gate(a, lens, y, L))
report("wn_gate (2x192 -> 192)", (a.numel() + y.numel()) * 4, ms)
w, b = torch.randn(192, 31, device=dev) * 0.1, torch.zeros(192, device=dev)
ms = timeit(lambda: ops.glu_dwconv(a, lens, y, w, b, b, torch.ones(192, device=dev), gam, bet, L))
report("glu_dwconv (k=31)", (a.numel() + y.numel()) * 4, ms)
x160 = torch.randn(B, 160, L, device=dev)
ml = torch.randn(B, 160, L, device=dev) * 0.1
ms = timeit(lambda: ops.flow_close(x160, ml, lens, L, torch.eye(4, device=dev), torch.zeros(160, device=dev), torch.zeros(160, device=dev)))
report("flow_close (C=160)", 3 * x160.numel() * 4, ms)
x256 = torch.randn(B, 256, L, device=dev)
y256 = torch.empty_like(x256)
ms = timeit(lambda: ops.group_norm(x256, lens, y256, torch.ones(256, device=dev), torch.zeros(256, device=dev), 32, L, tanh=True))
report("group_norm (32 groups, 3 passes)", 4 * x256.numel() * 4, ms)
z = torch.randn(B, 80, L, device=dev)
zs = torch.empty(B, 160, L // 2, device=dev)
ms = timeit(lambda: ops.squeeze2(z, lens, zs, L))
report("squeeze2", 2 * z.numel() * 4, ms)
