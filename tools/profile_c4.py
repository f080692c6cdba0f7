"""Where one rank's share of config 4 goes: python tools/profile_c4.py [n_utt=512] [max_batch=128]
Per batch of TextToWave: wall time of the acoustic call (host enqueue + its one D2H sync), GPU time of the acoustic model
and of the vocoder (CUDA events), wall time of the whole batch."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ims_toucan_prosody_variance_b200 as tb  # noqa: E402
from oracle import factory  # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 512
max_batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
voc_kind = sys.argv[3] if len(sys.argv) > 3 else "bigvgan"
graphs = len(sys.argv) > 4 and sys.argv[4] == "graphs"
dev = torch.device("cuda:0")
tts = tb.ToucanTTS(weights=factory.make_state_dict("toucantts", 1234), precision="tf32").to(dev)
tts.store_inverse_all()
if graphs:
    tts.enable_cuda_graphs()
path = f"/tmp/{voc_kind}_c4.pt"
torch.save({"generator": factory.make_state_dict(voc_kind, 1234)}, path)
voc = (tb.BigVGAN if voc_kind == "bigvgan" else tb.HiFiGANGenerator)(path, precision="f16", activation_dtype="f16").to(dev)
voc.remove_weight_norm()
eng = tb.TextToWave(tts, voc, max_batch=max_batch)
rng = random.Random(4)
lens = [rng.randint(20, 200) for _ in range(n_utt)]
texts = [factory.make_phoneme_tensor(n, 5000 + i) for i, n in enumerate(lens)]
emb = torch.stack([factory.make_utterance_embedding(i) for i in range(n_utt)])
lang = torch.full((n_utt,), 12, dtype=torch.int64)

rows = []
o_syn, o_voc = tts.synthesize_batch, voc.forward_batch


def syn(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    r = o_syn(*a, **k)
    e1.record()
    rows.append({"a_wall": (time.perf_counter() - t0) * 1e3, "a_ev": (e0, e1), "B": a[0].shape[0], "T": a[0].shape[1],
                 "frames": int(r["frames_host"].sum())})
    return r


def vocf(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    w = o_voc(*a, **k)
    e1.record()
    rows[-1].update({"v_wall": (time.perf_counter() - t0) * 1e3, "v_ev": (e0, e1)})
    return w


eng.synthesize(texts, emb, lang_ids=lang, noise="device", device=dev)
torch.cuda.synchronize()
tts.synthesize_batch, voc.forward_batch = syn, vocf
t0 = time.perf_counter()
waves = eng.synthesize(texts, emb, lang_ids=lang, noise="device", device=dev)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
audio = sum(int(w.numel()) for w in waves) / 24000
print(f"{n_utt} utterances, {audio:.1f} s audio, wall {wall:.1f} ms = {audio / wall * 1e3:.0f} audio-s/s, max_batch {max_batch}, graphs {graphs}")
print("batch   B   T  frames | acoustic wall  acoustic GPU | vocoder enqueue  vocoder GPU")
ta = tv = 0.0
for i, r in enumerate(rows):
    a = r["a_ev"][0].elapsed_time(r["a_ev"][1])
    v = r["v_ev"][0].elapsed_time(r["v_ev"][1])
    ta += a
    tv += v
    print(f"{i:5d} {r['B']:3d} {r['T']:3d} {r['frames']:7d} | {r['a_wall']:13.1f} {a:13.1f} | {r['v_wall']:15.1f} {v:12.1f}")
print(f"GPU time acoustic {ta:.1f} ms, vocoder {tv:.1f} ms, sum {ta + tv:.1f} ms, wall {wall:.1f} ms")
