"""End to end on the GPU: phoneme tensors -> waveform through TextToWave / ToucanTTSInterface, against the oracle
(acoustic restatement + vocoder restatement on the same weights, inputs and flow noise)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
_CACHE = {}


def _models(cuda, tmp_path):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory, restate
    if "m" not in _CACHE:
        tsd = factory.make_state_dict("toucantts", 1234)
        vsd = factory.make_state_dict("hifigan", 1234)
        vpath = os.path.join(tmp_path, "voc.pt")
        tpath = os.path.join(tmp_path, "tts.pt")
        torch.save({"generator": vsd}, vpath)
        torch.save({"model": tsd, "default_emb": factory.make_utterance_embedding(0)}, tpath)
        _CACHE["m"] = (tpath, vpath, restate.fold_weight_norm(tsd), restate.fold_weight_norm(vsd))
    return _CACHE["m"]


class _Frontend:
    """Stand-in text frontend: 'N:seed' -> oracle.factory phoneme tensor of N phonemes."""

    def string_to_tensor(self, text, input_phonemes=False):
        from oracle import factory
        n, seed = (int(v) for v in text.split(":"))
        return factory.make_phoneme_tensor(n, seed)


@pytest.mark.parametrize("kind", ["hifigan", "bigvgan"])
def test_text_to_wave_batch_vs_oracle(cuda, tmp_path, kind):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory, restate
    tpath, vpath, tfsd, vfsd = _models(cuda, tmp_path)
    ckpt = torch.load(tpath)
    tts = tb.ToucanTTS(weights=ckpt["model"], precision="fp32").to(cuda)
    tts.store_inverse_all()
    if kind == "bigvgan":
        bsd = factory.make_state_dict("bigvgan", 1234)
        vpath = os.path.join(tmp_path, "bvoc.pt")
        torch.save({"generator": bsd}, vpath)
        vfsd = restate.fold_weight_norm(bsd)
        voc = tb.BigVGAN(vpath, precision="f16").to(cuda)
    else:
        voc = tb.HiFiGANGenerator(vpath, precision="f16").to(cuda)
    voc.remove_weight_norm()
    eng = tb.TextToWave(tts, voc)
    lens = [17, 9, 26]
    texts = [factory.make_phoneme_tensor(n, 60 + i) for i, n in enumerate(lens)]
    embs = torch.stack([factory.make_utterance_embedding(60 + i) for i in range(3)])
    noise = torch.randn(3, 80, 400, generator=torch.Generator().manual_seed(11))
    batch = torch.zeros(3, max(lens), 62)
    for i, t in enumerate(texts):
        batch[i, :lens[i]] = t
    wave, wlen, r = eng.synthesize_padded(batch.to(cuda), torch.tensor(lens), embs.to(cuda), lang_ids=torch.tensor([12] * 3), noise=noise)
    torch.cuda.synchronize()
    for i in range(3):
        with torch.inference_mode():
            ref = restate.toucantts_forward(tfsd, texts[i], embs[i], lang_id=12, noise=noise[i, :, :int(r["frames_host"][i])])
            ref_wave = (restate.bigvgan_forward if kind == "bigvgan" else restate.hifigan_forward)(vfsd, ref["mel"].t())
        assert torch.equal(r["durations"][i, :lens[i]].cpu(), ref["durations"])
        assert int(wlen[i]) == ref_wave.numel()
        snr = restate.snr_db(wave[i, :int(wlen[i])].cpu(), ref_wave)
        assert snr >= 40.0, f"utterance {i}: text->wave SNR {snr:.1f} dB"


def test_interface_forward_batch_and_read_to_file(cuda, tmp_path):
    import wave as wavmod

    import ims_toucan_prosody_variance_b200 as tb
    tpath, vpath, _, _ = _models(cuda, tmp_path)
    tts = tb.ToucanTTSInterface(device="cuda", tts_model_path=tpath, vocoder_model_path=vpath, faster_vocoder=True,
                                language="en", text2phone=_Frontend())
    torch.manual_seed(5)
    single = tts("12:3")
    assert single.dim() == 1 and single.numel() % 384 == 0 and torch.isfinite(single).all()
    waves = tts.forward_batch(["12:3", "20:4", "7:5"], noise="device")
    assert len(waves) == 3 and waves[0].numel() == single.numel()      # durations do not depend on the flow noise
    out = os.path.join(tmp_path, "out.wav")
    wav = tts.read_to_file(["12:3", "", "20:4"], out, silent=True, increased_compatibility_mode=True)
    assert wav.numel() == 10600 * 3 + waves[0].numel() + waves[1].numel()
    with wavmod.open(out, "rb") as f:
        assert f.getframerate() == 48000 and f.getsampwidth() == 2 and f.getnframes() == 2 * wav.numel()


def test_read_to_file_content_parity(cuda, tmp_path):
    """read_to_file (ToucanTTSInterface.py:231-285): the file holds 10 600 samples of silence, then each non-empty
    sentence followed by silence; in increased-compatibility mode every sample twice as 16-bit PCM at 48 kHz
    (float2pcm, utils.py:20-33).  The written samples are compared bit for bit with the oracle's float2pcm of the
    same waveforms (same flow-noise seed), not only by their count."""
    import wave as wavmod

    import numpy as np

    import ims_toucan_prosody_variance_b200 as tb
    from oracle import restate
    tpath, vpath, _, _ = _models(cuda, tmp_path)
    tts = tb.ToucanTTSInterface(device="cuda", tts_model_path=tpath, vocoder_model_path=vpath, faster_vocoder=True,
                                language="en", text2phone=_Frontend())
    texts = ["12:3", "", "20:4", "9:1"]
    torch.manual_seed(11)
    waves = tts.forward_batch([t for t in texts if t.strip()])
    out = os.path.join(tmp_path, "content.wav")
    torch.manual_seed(11)
    wav = tts.read_to_file(texts, out, silent=True, increased_compatibility_mode=True).cpu()
    silence = torch.zeros(10600)
    expect = torch.cat([silence] + [p for w in waves for p in (w.reshape(-1).cpu(), silence)])
    assert torch.equal(wav, expect)
    with wavmod.open(out, "rb") as f:
        assert f.getframerate() == 48000 and f.getsampwidth() == 2 and f.getnchannels() == 1
        pcm = np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16)
    ref = restate.float2pcm(np.repeat(expect.numpy(), 2))
    assert pcm.shape == ref.shape and np.array_equal(pcm, ref)


def test_cloner_override_path(cuda, tmp_path):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    tpath, vpath, _, _ = _models(cuda, tmp_path)
    tts = tb.ToucanTTSInterface(device="cuda", tts_model_path=tpath, vocoder_model_path=vpath, text2phone=_Frontend())
    tts.set_utterance_embedding(embedding=factory.make_utterance_embedding(3))
    tts.set_accent_language("en")

    def extractor(transcript, path, lang):
        d, p, e = factory.make_gold_prosody(_Frontend().string_to_tensor(transcript), 3)
        return d, p, e, 100, 50

    tts._frontend_factory = lambda lang: _Frontend()
    cloner = tb.UtteranceCloner(model_id=None, device="cuda", tts=tts, prosody_extractor=extractor)
    tts.set_utterance_embedding = lambda **kw: None   # voice reference audio is outside the engine
    out = cloner.clone_utterance("intonation.wav", "voice.wav", "15:3", filename_of_result=os.path.join(tmp_path, "c.wav"), lang="en")
    d, _, _ = factory.make_gold_prosody(factory.make_phoneme_tensor(15, 3), 3)
    text = factory.make_phoneme_tensor(15, 3)
    d[text[:, factory.FEAT_WORD_BOUNDARY] == 1] = 0
    assert out.shape[0] == 150 * 3 + 2 * (int(d.sum()) // 2) * 384


def test_forward_batch_through_the_vectorised_tensoriser(cuda, tmp_path):
    """A frontend that exposes the reference's `phone_to_vector` table (and a feature index) is tensorised by
    frontend.PhoneTensoriser for the whole batch at once; the waveforms must equal those of the same frontend run
    through its own per-sentence string_to_tensor loop (here the oracle's restatement of TextFrontend.py:213-288)."""
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import restate
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend.pt"))

    class Loop:
        def string_to_tensor(self, text, input_phonemes=False):
            assert input_phonemes
            return restate.string_to_tensor(text, gold["phone_to_vector"], gold["feature_to_index"])

    class Vectorised(Loop):
        phone_to_vector = gold["phone_to_vector"]
        feature_to_index = gold["feature_to_index"]

        def string_to_tensor(self, text, input_phonemes=False):
            raise AssertionError("the batched path must not call the per-sentence loop")

    tpath, vpath, _, _ = _models(cuda, tmp_path)
    sentences = ["hˈɛloʊ wˈɜːld~#", "ðɪs ɪz ɐ tˈɛst.~#", "aː˥ b̃ˈa~#"]
    results = []
    for fe in (Loop(), Vectorised()):
        tts = tb.ToucanTTSInterface(device="cuda", tts_model_path=tpath, vocoder_model_path=vpath, faster_vocoder=True,
                                    language="en", text2phone=fe)
        torch.manual_seed(3)
        results.append(tts.forward_batch(sentences, input_is_phones=True))
    for a, b in zip(*results):
        assert a.numel() > 0 and torch.equal(a, b)


def test_empty_and_one_phoneme_utterances_inside_a_batch(cuda, tmp_path):
    """Ragged edge cases: an utterance of length 0 yields an empty waveform, a 1-phoneme utterance a short one, and
    neither disturbs its neighbours (same waveforms as the batch without them)."""
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    tpath, vpath, _, _ = _models(cuda, tmp_path)
    tts = tb.ToucanTTS(weights=torch.load(tpath)["model"]).to(cuda)
    tts.store_inverse_all()
    voc = tb.HiFiGANGenerator(vpath).to(cuda)
    voc.remove_weight_norm()
    eng = tb.TextToWave(tts, voc)
    texts = [factory.make_phoneme_tensor(14, 70), torch.zeros((0, 62)), factory.make_phoneme_tensor(1, 71),
             factory.make_phoneme_tensor(9, 72)]
    emb = torch.stack([factory.make_utterance_embedding(70 + i) for i in range(4)])
    noise = torch.randn((4, 80, 400), generator=torch.Generator().manual_seed(8))
    order = sorted(range(4), key=lambda i: (-texts[i].shape[0], i))                 # the batch is length-sorted inside
    waves = eng.synthesize(texts, emb, lang_ids=12, noise=noise[order])
    assert waves[1].numel() == 0
    assert waves[2].numel() % 384 == 0 and torch.isfinite(waves[2]).all()
    keep = [0, 3]
    ref = eng.synthesize([texts[i] for i in keep], emb[keep], lang_ids=12, noise=noise[keep])
    for i, r in zip(keep, ref):
        assert waves[i].shape == r.shape and torch.allclose(waves[i], r, atol=2e-4), i
