"""CPU: the oracle restatement against the committed golden fixtures (outputs of the live reference),
against the live reference itself where it exists, and the third-party alias_free_torch restatement."""
import os

import pytest
import torch

from oracle import factory, restate, shim

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_vocoder_oracle_matches_golden():
    g = torch.load(os.path.join(GOLD, "vocoder.pt"))
    mel = factory.make_mel(1, g["frames"], seed=g["seed"])[0]
    for kind, fwd in (("hifigan", restate.hifigan_forward), ("bigvgan", restate.bigvgan_forward)):
        wave = fwd(restate.fold_weight_norm(factory.make_state_dict(kind, 1234)), mel)
        assert wave.shape == g[kind].shape
        assert restate.snr_db(wave, g[kind].float()) > 60.0  # fixture is fp16-quantised


def test_toucantts_oracle_matches_golden():
    g = torch.load(os.path.join(GOLD, "toucantts.pt"))
    fsd = restate.fold_weight_norm(factory.make_state_dict("toucantts", 1234))
    for case in g["cases"]:
        text = factory.make_phoneme_tensor(case["n_ph"], case["seed"])
        emb = factory.make_utterance_embedding(case["seed"])
        kw = dict(case["kw"])
        if case["gold"]:
            d, p, e = factory.make_gold_prosody(text, case["seed"])
            kw.update(durations=d, pitch=p, energy=e)
        torch.manual_seed(case["noise_seed"])
        out = restate.toucantts_forward(fsd, text, emb, g["lang_id"], **kw)
        assert torch.equal(out["durations"], case["durations"])  # integer exact
        assert restate.rel_l1(out["mel"], case["mel"]) < 1e-5
        assert restate.rel_l1(out["pitch"], case["pitch"]) < 1e-4
        assert restate.rel_l1(out["energy"], case["energy"]) < 1e-4


def test_aa_filter_taps_pinned():
    """SURVEY.md 8c probe values of kaiser_sinc_filter1d(0.25, 0.3, 12); symmetric, unit sum."""
    f = restate.aa_filter().flatten()
    probe = torch.tensor([0.0020290, 0.0093895, -0.0255435, -0.0576574, 0.1285726, 0.4432098])
    assert torch.allclose(f[:6], probe, atol=2e-7)
    assert torch.allclose(f, f.flip(0), atol=1e-8)
    assert abs(float(f.sum()) - 1.0) < 1e-6


def test_aa_filter_cross_check_transformers():
    """Independent copy of the same published algorithm (transformers qwen2_5_omni)."""
    mod = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    fn = getattr(mod, "kaiser_sinc_filter1d", None)
    if fn is None:
        pytest.skip("transformers build without kaiser_sinc_filter1d")
    other = fn(0.25, 0.3, 12).flatten().float()
    assert torch.allclose(restate.aa_filter().flatten(), other, atol=1e-7)


def test_squeeze_unsqueeze_roundtrip():
    x = torch.randn(80, 11)
    s = restate.squeeze2(x)
    assert s.shape == (160, 5)
    assert torch.equal(restate.unsqueeze2(s), x[:, :10])
    assert torch.equal(s[80 + 3, 2], x[3, 5])


def test_duration_rounding_half_even():
    logd = torch.log(torch.tensor([1.5, 2.5, 3.5, 0.2]))
    assert restate.durations_from_log(logd).tolist() == [0, 2, 2, 0]


def test_edit_prosody_order_and_rescue():
    text = factory.make_phoneme_tensor(12, 3)
    d = torch.full((12,), 3)
    d2, p2, e2 = restate.edit_prosody(text, d, torch.ones(12), torch.ones(12), 1.5, 1.1, 1.0, 1.0)
    wb = text[:, factory.FEAT_WORD_BOUNDARY] == 1
    sil = text[:, factory.FEAT_SILENCE] == 1
    assert torch.all(d2[wb] == 0)
    assert torch.all(d2[sil] == round(round(3 * 1.5) * 1.1))
    assert torch.all(p2[text[:, factory.FEAT_VOICED] == 0] == 0)
    up, used = restate.length_regulate(torch.randn(4, 3), torch.zeros(4, dtype=torch.long))
    assert up.shape[0] == 4 and used.tolist() == [1, 1, 1, 1]


@pytest.mark.reference
@pytest.mark.skipif(not shim.available(), reason="live reference not present (GPU box)")
def test_restatement_vs_live_reference():
    cls = shim.reference_classes()
    sd = factory.make_state_dict("toucantts", 4321)
    tts = cls["InfToucanTTS"](weights=sd)
    tts.store_inverse_all()
    fsd = restate.fold_weight_norm(sd)
    text = factory.make_phoneme_tensor(17, 8)
    emb = factory.make_utterance_embedding(8)
    kw = dict(duration_scaling_factor=0.9, pause_duration_scaling_factor=1.3, pitch_variance_scale=0.7,
              energy_variance_scale=1.3)
    torch.manual_seed(3)
    mel, dur, pitch, energy = tts(text, utterance_embedding=emb, lang_id=torch.tensor([12]),
                                  return_duration_pitch_energy=True, **kw)
    torch.manual_seed(3)
    out = restate.toucantts_forward(fsd, text, emb, 12, **kw)
    assert torch.equal(out["durations"], dur)
    assert restate.rel_l1(out["mel"], mel) < 1e-5


@pytest.mark.reference
@pytest.mark.skipif(not shim.available(), reason="live reference not present (GPU box)")
def test_reference_random_init_vocoders(tmp_path):
    """The reference's OWN random init (train-side classes), not the factory's."""
    cls = shim.reference_classes()
    torch.manual_seed(1234)
    for train, inf, fwd in (("TrainHiFiGAN", "InfHiFiGAN", restate.hifigan_forward),
                            ("TrainBigVGAN", "InfBigVGAN", restate.bigvgan_forward)):
        sd = cls[train]().state_dict()
        path = os.path.join(tmp_path, train + ".pt")
        torch.save({"generator": sd}, path)
        model = cls[inf](path_to_weights=path)
        model.remove_weight_norm()
        mel = factory.make_mel(1, 19, seed=4)[0]
        with torch.inference_mode():
            ref = model(mel)
        assert restate.snr_db(fwd(restate.fold_weight_norm(sd), mel), ref) > 90.0
