"""Kernel-level time breakdown of one acoustic-model batch (config 3 shape): python tools/profile_tts.py [n_utt] [precision]"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import ims_toucan_prosody_variance_b200 as tb  # noqa: E402
from oracle import factory  # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 128
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
dev = torch.device("cuda:0")
tts = tb.ToucanTTS(weights=factory.make_state_dict("toucantts", 1234), precision=prec).to(dev)
tts.store_inverse_all()
rng = random.Random(3)
lens = [rng.randint(20, 200) for _ in range(n_utt)]
text = torch.zeros((n_utt, max(lens), 62))
for i, n in enumerate(lens):
    text[i, :n] = factory.make_phoneme_tensor(n, i)
text = text.to(dev)
emb = torch.stack([factory.make_utterance_embedding(i) for i in range(n_utt)]).to(dev)
tlen = torch.tensor(lens, dtype=torch.int32)
lang = torch.full((n_utt,), 12, dtype=torch.int64)
for _ in range(2):
    r = tts.synthesize_batch(text, tlen, utterance_embedding=emb, lang_ids=lang, noise="device")
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
r = tts.synthesize_batch(text, tlen, utterance_embedding=emb, lang_ids=lang, noise="device")
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {t_host * 1e3:.1f} ms, wall {t_all * 1e3:.1f} ms, frames {int(r['frames_host'].sum())}")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tts.synthesize_batch(text, tlen, utterance_embedding=emb, lang_ids=lang, noise="device")
    torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg.setdefault(ev.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"GPU kernel time {tot / 1e3:.2f} ms")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}% {n:5d}x  {name}")

# ---- per-shape breakdown of the conv launches (CUDA events around each call) ----
from ims_toucan_prosody_variance_b200 import ops  # noqa: E402

records = []
orig = ops.ConvLayer.__call__


def timed(self, x, lengths, out, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = orig(self, x, lengths, out, **kw)
    e1.record()
    records.append((self, kw.get("l_in_max", x.shape[2]), x.shape[0], e0, e1))
    return res


ops.ConvLayer.__call__ = timed
tts.synthesize_batch(text, tlen, utterance_embedding=emb, lang_ids=lang, noise="device")
torch.cuda.synchronize()
ops.ConvLayer.__call__ = orig
agg = {}
for layer, L, B, e0, e1 in records:
    key = (layer.c_in, layer.c_out, layer.k, L, layer.precision)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"conv launches by shape (total {tot:.2f} ms): Cin Cout K Lmax prec  n  ms  ms/launch  TFLOP/s(dense over B*Lmax)")
for (cin, cout, k, L, pr), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    fl = 2.0 * cin * cout * k * n_utt * L * n
    print(f"{cin:5d} {cout:5d} {k:2d} {L:5d} {pr}  {n:4d} {ms:8.3f} {ms / n:7.3f} {fl / ms / 1e9:8.1f}")
