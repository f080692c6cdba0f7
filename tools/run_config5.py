"""BASELINE.json configs[4]: UtteranceCloner-style prosody override on a long-form 2,000-phoneme input
(external durations / pitch / energy), text -> mel -> wave on one GPU.  Prints timings and sanity checks."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ims_toucan_prosody_variance_b200 as tb  # noqa: E402
from oracle import factory, restate  # noqa: E402

dev = torch.device("cuda:0")
n_ph = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
tts = tb.ToucanTTS(weights=factory.make_state_dict("toucantts", 1234)).to(dev)
tts.store_inverse_all()
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "v.pt")
    torch.save({"generator": factory.make_state_dict("bigvgan", 1234)}, path)
    voc = tb.BigVGAN(path).to(dev)
voc.remove_weight_norm()
text = factory.make_phoneme_tensor(n_ph, 5)
emb = factory.make_utterance_embedding(5)
d, p, e = factory.make_gold_prosody(text, 5)
ref_d, _, _ = restate.edit_prosody(text, d, p.reshape(-1), e.reshape(-1), 1.2, 1.1, 1.2, 0.8)


def run():
    mel, dur, pitch, energy = tts(text.to(dev), durations=d.clone(), pitch=p.clone(), energy=e.clone(), utterance_embedding=emb.to(dev),
                                  lang_id=torch.tensor([12]).to(dev), return_duration_pitch_energy=True, duration_scaling_factor=1.1,
                                  pause_duration_scaling_factor=1.2, pitch_variance_scale=1.2, energy_variance_scale=0.8)
    wave = voc(mel.transpose(0, 1).contiguous())
    return mel, dur, wave


for _ in range(2):
    mel, dur, wave = run()
torch.cuda.synchronize()
t0 = time.perf_counter()
mel, dur, wave = run()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
assert torch.equal(dur.cpu(), ref_d), "durations differ from the oracle's edit loop"
assert mel.shape == (2 * (int(ref_d.sum()) // 2), 80) and wave.numel() == mel.shape[0] * 384
assert torch.isfinite(wave).all() and float(wave.abs().max()) <= 1.0
audio = wave.numel() / 24000.0
print(f"config 5: {n_ph} phonemes -> {mel.shape[0]} frames -> {audio:.1f} s of audio in {dt * 1e3:.1f} ms wall "
      f"({audio / dt:.0f} audio-s/s, RTF {dt / audio:.5f}); durations bit-exact vs the oracle edit loop; "
      f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
