"""Generate tests/golden/*.pt from the LIVE reference (TEST INFRASTRUCTURE; authoring container only).

    python -m oracle.make_golden

The reference has no golden vectors of its own, so these fixtures are outputs of
the reference's unmodified inference modules (InferenceToucanTTS.ToucanTTS, InferenceAvocodo.
HiFiGANGenerator, InferenceBigVGAN.BigVGAN) loaded with oracle.factory weights, on oracle.factory
inputs.  Waveforms are stored as fp16 (quantisation noise ~ -70 dB, far below the 40 dB bound),
mels as fp32."""
import os
import tempfile

import torch

from oracle import factory, shim


def main():
    cls = shim.reference_classes()
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # ---- vocoders: one 24-frame mel ----
    voc = {"frames": 24, "seed": 77}
    mel = factory.make_mel(1, voc["frames"], seed=voc["seed"])[0]
    for kind, name in (("hifigan", "InfHiFiGAN"), ("bigvgan", "InfBigVGAN")):
        sd = factory.make_state_dict(kind, 1234)
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "g.pt")
            torch.save({"generator": sd}, path)
            model = cls[name](path_to_weights=path)
        model.remove_weight_norm()
        model.eval()
        with torch.inference_mode():
            voc[kind] = model(mel).clone().half()
    torch.save(voc, os.path.join(out_dir, "vocoder.pt"))

    # ---- acoustic model: predicted prosody, scaled prosody, external (cloner-shaped) prosody ----
    sd = factory.make_state_dict("toucantts", 1234)
    tts = cls["InfToucanTTS"](weights=sd)
    tts.store_inverse_all()
    cases = []
    for n_ph, seed, kw in ((14, 1, {}),
                           (19, 2, dict(duration_scaling_factor=1.1, pause_duration_scaling_factor=1.2,
                                        pitch_variance_scale=1.2, energy_variance_scale=0.8)),
                           (11, 3, "gold")):
        text = factory.make_phoneme_tensor(n_ph, seed)
        emb = factory.make_utterance_embedding(seed)
        args = {}
        if kw == "gold":
            d, p, e = factory.make_gold_prosody(text, seed)
            args = dict(durations=d.clone(), pitch=p.clone(), energy=e.clone())
            kw = {}
        torch.manual_seed(1000 + seed)  # the PostFlow noise comes from the global CPU generator (Glow.py:363)
        mel_out, dur, pitch, energy = tts(text, utterance_embedding=emb, lang_id=torch.tensor([12]),
                                          return_duration_pitch_energy=True, **args, **kw)
        cases.append(dict(n_ph=n_ph, seed=seed, kw=kw, gold=bool(args), noise_seed=1000 + seed, mel=mel_out.clone(),
                          durations=dur.clone(), pitch=pitch.clone(), energy=energy.clone()))
    torch.save({"cases": cases, "lang_id": 12}, os.path.join(out_dir, "toucantts.pt"))
    for f in os.listdir(out_dir):
        print(f, os.path.getsize(os.path.join(out_dir, f)), "bytes")


if __name__ == "__main__":
    main()
