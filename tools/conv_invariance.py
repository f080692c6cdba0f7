"""Diagnostics: is one conv launch invariant to the batch size (tiling)?  python tools/conv_invariance.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ims_toucan_prosody_variance_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (C, K, d, L) in ((128, 7, 1, 24000), (128, 11, 5, 24000), (128, 3, 1, 24000), (256, 7, 1, 4000), (64, 11, 1, 96000)):
    w = (torch.randn(C, C, K, generator=g) * 0.05).to(dev)
    layer = ops.ConvLayer(w, torch.zeros(C, device=dev), dilation=d, padding=(K - 1) // 2 * d, precision="f16")
    x = torch.randn(16, C, L, generator=g).to(dev)
    outs = {}
    for B in (1, 8, 16):
        y = torch.zeros(B, C, L, device=dev)
        lens = torch.full((B,), L, dtype=torch.int32, device=dev)
        layer(x[:B].contiguous(), lens, y, act=1, slope=0.1)
        torch.cuda.synchronize()
        outs[B] = y[0].clone()
    for B in (8, 16):
        dd = (outs[B] - outs[1]).abs()
        print(f"C={C} K={K} d={d}: B={B} vs B=1 max abs diff {float(dd.max()):.3e} differing {int((dd > 0).sum())} rel rms {float(dd.pow(2).mean().sqrt() / outs[1].pow(2).mean().sqrt()):.2e}", flush=True)
