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
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn((cin, cout, 2 * up) if up else (cout, cin, k), generator=g) * 0.05
layer = ops.ConvLayer(w.to(dev), torch.zeros(cout, device=dev), dilation=dil, padding=(k - 1) // 2 * dil,
                      transposed_stride=up, precision=prec)
x = torch.randn(B, cin, L, device=dev)
res = torch.randn(B, cout, L * (up or 1), device=dev)
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
