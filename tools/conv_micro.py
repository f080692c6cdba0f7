"""Single-conv microbenchmark (for ncu): python tools/conv_micro.py Cin Cout K dil up L B act prec [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ims_toucan_prosody_variance_b200 import ops  # noqa: E402

cin, cout, k, dil, up, L, B, act = (int(v) for v in sys.argv[1:9])
prec = sys.argv[9] if len(sys.argv) > 9 else "f16"
reps = int(sys.argv[10]) if len(sys.argv) > 10 else 5
use_res = not (len(sys.argv) > 11 and sys.argv[11] == "nores")
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn((cin, cout, 2 * up) if up else (cout, cin, k), generator=g) * 0.05
layer = ops.ConvLayer(w.to(dev), torch.zeros(cout, device=dev), dilation=dil, padding=(k - 1) // 2 * dil,
                      transposed_stride=up, precision=prec)
x = torch.randn(B, cin, L, device=dev)
res = torch.randn(B, cout, L * (up or 1), device=dev) if use_res else None
y = torch.zeros(B, cout, L * (up or 1), device=dev)
alpha = torch.zeros(cin, device=dev)
beta = torch.zeros(cin, device=dev)
lens = torch.full((B,), L, dtype=torch.int32, device=dev)
for _ in range(2):
    layer(x, lens, y, act=act, slope=0.1, alpha=alpha, beta=beta, residual=res)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    layer(x, lens, y, act=act, slope=0.1, alpha=alpha, beta=beta, residual=res)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 2.0 * B * L * cin * cout * (up or 1) * (2 if up else k)
byts = B * L * (cin * 4 + cout * (up or 1) * 8)
print(f"Cin={cin} Cout={cout} K={k} dil={dil} up={up} L={L} B={B} act={act} {prec}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {byts / ms / 1e6:.0f} GB/s")

if os.environ.get("TB200_TRACE"):
    import ctypes

    from ims_toucan_prosody_variance_b200 import _lib
    buf = (ctypes.c_int64 * (96 * 8 + 160 + 64))()
    _lib.check(_lib.load().tb200_debug_trace_read(ctypes.cast(buf, ctypes.c_void_p), 96 * 8 + 160 + 64), "trace")
    t = torch.tensor(list(buf)[:96 * 8], dtype=torch.int64).reshape(96, 8)
    per_cta = [(v & ((1 << 48) - 1), v >> 48) for v in list(buf)[96 * 8:96 * 8 + 160] if v]
    pw = list(buf)[96 * 8 + 160:]
    w0 = min(v for v in pw[0::2] if v)
    print("producer warps, 6th buffer (start, end, busy) cycles:", [(pw[2 * i] - w0, pw[2 * i + 1] - w0, pw[2 * i + 1] - pw[2 * i]) for i in range(32) if pw[2 * i]])
    cyc = sorted(c for c, _ in per_cta)
    print(f"per-CTA elapsed cycles over {len(cyc)} CTAs: min {cyc[0]} median {cyc[len(cyc) // 2]} max {cyc[-1]}")
    print("slowest CTAs (cycles, smid):", sorted(per_cta, reverse=True)[:6], " fastest:", sorted(per_cta)[:4])
    base = int(t[:, :7][t[:, :7] > 0].min())
    print("tile  a_empty  staged | a_full  mma_issued  acc_empty(mma) | acc_full  drained   (clock cycles since first stamp)")
    for i in range(24):
        r = [int(v) - base if int(v) > 0 else -1 for v in t[i, :7]]
        ww = int(t[i, 7]) - (int(t[i - 1, 7]) if i else 0)
        print(f"{i:4d} {r[0]:8d} {r[1]:8d} | {r[2]:8d} {r[3]:8d} {r[6]:8d} | {r[4]:8d} {r[5]:8d}   stage {r[1]-r[0]:6d}  mma-issue {r[3]-r[2]:6d}  (weight wait {ww:6d})  mma->accfull {r[4]-r[3]:6d}  drain {r[5]-r[4]:6d}")
