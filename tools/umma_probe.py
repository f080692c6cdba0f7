"""Bring-up probe for the tcgen05 conv path: tiny GEMM-shaped cases, printed error summaries.
Run on the GPU box:  python tools/umma_probe.py   (each variant in its own process: a trap kills the context)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASE = r'''
import sys, torch
sys.path.insert(0, %r)
import torch.nn.functional as F
from ims_toucan_prosody_variance_b200 import ops
dev = torch.device("cuda:0")
def run(prec, Cin, Cout, K, L, dil=1):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, Cin, L, generator=g); w = torch.randn(Cout, Cin, K, generator=g) / (Cin*K)**0.5
    ref = F.conv1d(x, w, None, dilation=dil, padding=(K-1)//2*dil)
    layer = ops.ConvLayer(w.to(dev), None, dilation=dil, padding=(K-1)//2*dil, precision=prec)
    y = torch.zeros(1, Cout, L, device=dev)
    layer(x.to(dev), None, y); torch.cuda.synchronize()
    got = y.cpu()
    err = (got-ref).abs().max().item(); rms = ref.pow(2).mean().sqrt().item()
    print(f"  {prec} Cin={Cin} Cout={Cout} K={K} L={L} dil={dil}: max err {err:.4e} (ref rms {rms:.3f}) nonzero={int((got!=0).sum())}/{got.numel()}", flush=True)
    if err > 0.05:
        print("   got[0,:4,:6]", got[0,:4,:6].tolist()); print("   ref[0,:4,:6]", ref[0,:4,:6].tolist())
for prec in %r:
    run(prec, 16, 16, 1, 128)
    run(prec, 32, 32, 1, 128)
    run(prec, 64, 48, 3, 300)
    run(prec, 32, 32, 11, 400, 5)
    run(prec, 256, 256, 3, 300)
'''

for swap in ("0",):
    for precs in (("f16",), ("tf32",)):
        print(f"== TB200_DESC_SWAP={swap} {precs}", flush=True)
        env = dict(os.environ, TB200_DESC_SWAP=swap)
        try:
            r = subprocess.run([sys.executable, "-c", CASE % (ROOT, precs)], env=env, capture_output=True, text=True, timeout=240)
            print(r.stdout[-3000:]); print(r.stderr[-1500:])
        except subprocess.TimeoutExpired:
            print("  TIMEOUT")
