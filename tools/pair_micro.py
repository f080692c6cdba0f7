"""Single residual-pair microbenchmark: python tools/pair_micro.py C K dil L B snake(0|1) dtype(f16|f32) [reps] [unfused]

Times tb200_respair (or, with `unfused`, the two tb200_conv1d launches it replaces) with CUDA events, and with
TB200_TRACE=1 prints the per-tile timeline of CTA 0 (clock cycles per pipeline stage)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ims_toucan_prosody_variance_b200 import ops  # noqa: E402

C, K, dil, L, B, snake = (int(v) for v in sys.argv[1:7])
dt = torch.float16 if (len(sys.argv) > 7 and sys.argv[7] == "f16") else torch.float32
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
mode = sys.argv[9] if len(sys.argv) > 9 else "fused"     # fused | unfused (two tb200_conv1d) | staged (two tb200_conv1d_staged)
unfused = mode in ("unfused", "staged")
staged = mode == "staged"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w1 = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
w2 = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
c1 = ops.ConvLayer(w1.to(dev), torch.zeros(C, device=dev), dilation=dil, padding=(K - 1) // 2 * dil, precision="f16")
c2 = ops.ConvLayer(w2.to(dev), torch.zeros(C, device=dev), dilation=1, padding=(K - 1) // 2, precision="f16")
act = (torch.zeros(C, device=dev), torch.zeros(C, device=dev))
pair = ops.ResPair(c1, c2, act if snake else None, act if snake else None)
Lp = (L + 7) // 8 * 8
x = torch.randn(B, C, Lp, device=dev).to(dt)
y = torch.zeros(B, C, Lp, device=dev, dtype=dt)
t = torch.zeros(B, C, Lp, device=dev, dtype=torch.float16)
lens = torch.full((B,), L, dtype=torch.int32, device=dev)


def step():
    if unfused:
        a = 2 if snake else 1
        c1(x, lens, t, l_in_max=L, act=a, slope=0.1, alpha=act[0], beta=act[1], staged=staged)
        c2(t, lens, y, l_in_max=L, act=a, slope=0.1, alpha=act[0], beta=act[1], residual=None if os.environ.get("PM_NORES") else x,
           staged=staged)
    else:
        pair(x, lens, y, l_max=L, slope=0.1)


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 2.0 * B * L * C * C * K * 2
esz = 2 if dt == torch.float16 else 4
byts = B * L * C * esz * 2
print(f"C={C} K={K} dil={dil} L={L} B={B} snake={snake} {sys.argv[7] if len(sys.argv) > 7 else 'f32'} {mode}: "
      f"{ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {byts / ms / 1e6:.0f} GB/s (algorithmic: x in + y out)  "
      f"{ms * 1e-3 * 1.9e9 * 148 / (B * L * C):.3f} SM-cycles/element @1.9GHz")

if os.environ.get("TB200_TRACE") and mode != "unfused":
    import ctypes

    from ims_toucan_prosody_variance_b200 import _lib
    n = 64 * 16 + 160
    buf = (ctypes.c_int64 * n)()
    _lib.check(_lib.load().tb200_respair_trace_read(ctypes.cast(buf, ctypes.c_void_p), n), "trace")
    tr = torch.tensor(list(buf)[:64 * 16], dtype=torch.int64).reshape(64, 16)
    per_cta = sorted(v for v in list(buf)[64 * 16:] if v)
    print(f"per-CTA elapsed cycles over {len(per_cta)} CTAs: min {per_cta[0]} median {per_cta[len(per_cta) // 2]} max {per_cta[-1]}")
    base = int(tr[:, :13][tr[:, :13] > 0].min())
    print("tile |   P1 beg    P1 end (dur) |   M1 beg  issued |   E1 beg    E1 end (dur) |   P2 beg    P2 end (dur) |   M2 beg  issued |   E2 beg    E2 end (dur) | X landed")
    for i in range(20):
        r = [int(v) - base if int(v) > 0 else -1 for v in tr[i, :13]]
        if r[0] < 0:
            break
        print(f"{i:4d} | {r[0]:8d} {r[1]:8d} ({r[1]-r[0]:6d}) | {r[8]:8d} {r[9]:8d} | {r[4]:8d} {r[5]:8d} ({r[5]-r[4]:6d}) | "
              f"{r[2]:8d} {r[3]:8d} ({r[3]-r[2]:6d}) | {r[10]:8d} {r[11]:8d} | {r[6]:8d} {r[7]:8d} ({r[7]-r[6]:6d}) | {r[12]:8d}")
