"""Attention microbenchmark (for ncu): python tools/attn_micro.py B L [tc=1] -- config-3 decoder shape by default."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ims_toucan_prosody_variance_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
tc = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
dev = torch.device("cuda:0")
heads, dk = 4, 48
d = heads * dk
rng = random.Random(3)
lens = [rng.randint(L // 10, L) for _ in range(B)]
qkv = torch.randn(B, 3 * d, (L + 3) // 4 * 4, device=dev)
out = torch.zeros(B, d, (L + 3) // 4 * 4, device=dev)
cap = 256
while cap < L:
    cap *= 2
pos = torch.randn(d, 2 * cap - 1, device=dev) * 0.5
bu, bv = torch.randn(heads, dk, device=dev) * 0.3, torch.randn(heads, dk, device=dev) * 0.3
lt = torch.tensor(lens, dtype=torch.int32, device=dev)
pos16, center = ops.pack_relpos_table(pos, cap - 1, heads)
posp = torch.zeros(d, (2 * cap - 1 + 3) // 4 * 4, device=dev)
posp[:, :2 * cap - 1] = pos


def step():
    if tc:
        ops.relpos_attention_tc(qkv, lt, out, pos16, center, bu, bv, heads, L)
    else:
        ops.relpos_attention(qkv, lt, out, posp, cap - 1, bu, bv, heads, L)


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
scores = sum(n * n for n in lens) * heads
print(f"B={B} L<={L} heads={heads} dk={dk} {'tcgen05' if tc else 'simt fp32'}: {ms:.3f} ms  {scores * 6 * dk / ms / 1e9:.1f} TFLOP/s "
      f"(6 dk FLOP per score)  {ms * 1e-3 * 1.9e9 * 148 / scores:.3f} SM-cycles/score @1.9GHz")
