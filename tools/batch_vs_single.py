"""Diagnostics: which intermediate of the generator first differs between two batch sizes (utterance 0)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ims_toucan_prosody_variance_b200 as tb
from oracle import factory, restate
dev = torch.device("cuda:0")
kind = sys.argv[1] if len(sys.argv) > 1 else "hifigan"
sd = factory.make_state_dict(kind, 1234)
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "g.pt"); torch.save({"generator": sd}, path)
    cls = tb.BigVGAN if kind == "bigvgan" else tb.HiFiGANGenerator
    m = cls(path, precision="f16").to(dev); m.remove_weight_norm()
F = 500
melF = factory.make_mel(64, F, seed=100).to(dev)
snap = {}
for B in (8, 16):
    m.forward_batch(melF[:B].contiguous(), torch.full((B,), F, dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    ws = m._buffers_cache[(B, F, str(dev))]
    snap[B] = {k: v[0].float().clone() for k, v in ws.items()}
for k in snap[8]:
    a, b = snap[8][k], snap[16][k]
    d = (a - b).abs()
    print(f"{k:6s} shape {tuple(a.shape)} max abs diff {float(d.max()):.3e}  rel rms {float(d.pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-20)):.3e}  differing {int((d > 0).sum())}")
