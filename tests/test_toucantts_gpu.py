"""ToucanTTS acoustic model on the GPU against the oracle (stage by stage), the golden fixtures of the
live reference, and batch-vs-single consistency.

Tolerances (north_star): integer durations and frame counts bit-exact; mel relative L1 <= 1e-3 in the
fp32-accumulate modes ("fp32" = CUDA-core fp32, "tf32" = tcgen05 kind::tf32 with fp32 accumulation in TMEM);
the fp16-operand mode (fp32 accumulation, the mantissa of tf32) is held to the same bound."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEL_TOL = {"fp32": 1e-3, "tf32": 1e-3, "f16": 1e-3}   # fp16 operands carry tf32's mantissa: same bound (measured 2.0e-4)
_ENGINES = {}


def _engine(cuda, prec):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory, restate
    if prec not in _ENGINES:
        sd = factory.make_state_dict("toucantts", 1234)
        model = tb.ToucanTTS(weights=sd, precision=prec).to(cuda)
        model.store_inverse_all()
        _ENGINES[prec] = (model, restate.fold_weight_norm(sd))
    return _ENGINES[prec]


def _rel_l1(a, b):
    return ((a - b).abs().mean() / b.abs().mean().clamp_min(1e-12)).item()


def _log(lines):
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "toucantts_stage_errors.txt"), "a") as f:
            f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


@pytest.mark.parametrize("prec", ["fp32", "tf32", "f16"])
def test_stagewise_vs_oracle(cuda, prec):
    from oracle import factory, restate
    model, fsd = _engine(cuda, prec)
    n_ph, seed = 23, 5
    text = factory.make_phoneme_tensor(n_ph, seed)
    emb = factory.make_utterance_embedding(seed)
    otaps = {}
    with torch.inference_mode():
        ref = restate.toucantts_forward(fsd, text, emb, lang_id=12, generator=torch.Generator().manual_seed(99), taps=otaps)
    frames = int(ref["durations"].sum())
    noise = torch.randn((1, 80, frames), generator=torch.Generator().manual_seed(99))

    taps = {}
    r = model.synthesize_batch(text.unsqueeze(0).to(cuda), torch.tensor([n_ph]), utterance_embedding=emb.unsqueeze(0).to(cuda),
                               lang_ids=torch.tensor([12]), noise=noise, taps=taps)
    torch.cuda.synchronize()
    lines = [f"== stagewise {prec}: {n_ph} phonemes, {frames} frames"]
    enc_err = _rel_l1(taps["encoder"][0, :, :n_ph].t().cpu(), otaps["encoder"])
    lines.append(f"encoder rel-L1 {enc_err:.3e}")
    lines.append(f"log-durations max abs err {(r['log_durations'][0, :n_ph].cpu() - ref['log_durations']).abs().max().item():.3e}")
    lines.append(f"pitch rel-L1 {_rel_l1(r['pitch'][0, :n_ph].cpu(), ref['pitch']):.3e}  energy rel-L1 "
                 f"{_rel_l1(r['energy'][0, :n_ph].cpu(), ref['energy']):.3e}")
    dur_equal = torch.equal(r["durations"][0, :n_ph].cpu(), ref["durations"])
    lines.append(f"durations equal: {dur_equal}  frames engine {int(r['frames'][0])} oracle {frames}")
    errs = {}
    if dur_equal:
        for name, key in (("upsampled", "upsampled"), ("decoder", "decoder"), ("decoded", "decoded"), ("refined", "refined")):
            errs[name] = _rel_l1(taps[name][0, :, :frames].t().cpu(), otaps[key])
            lines.append(f"{name} rel-L1 {errs[name]:.3e}")
        for i, xb in enumerate(taps["flow_blocks"]):
            b = 17 - i
            e = _rel_l1(xb[0, :, :frames // 2].cpu(), otaps[f"post_flow.block{b}"])
            if i % 6 == 5 or i == 0:
                lines.append(f"flow block {b} rel-L1 {e:.3e}")
        m = 2 * (frames // 2)
        errs["mel"] = _rel_l1(r["mel_ncl"][0, :, :m].t().cpu(), ref["mel"])
        lines.append(f"mel rel-L1 {errs['mel']:.3e}  (bound {MEL_TOL[prec]:.0e})")
    _log(lines)
    assert int(r["frames"][0]) == int(r["durations"][0, :n_ph].sum())
    if prec == "fp32":
        assert dur_equal, "fp32 mode must reproduce the oracle's integer durations"
    if dur_equal:
        assert enc_err < (1e-4 if prec == "fp32" else 5e-3)
        assert errs["upsampled"] < (1e-4 if prec == "fp32" else 5e-3)
        assert errs["mel"] <= MEL_TOL[prec], f"mel rel-L1 {errs['mel']:.3e}"
    else:
        pytest.fail(f"{prec}: durations differ from the oracle's (see gpurun_out/toucantts_stage_errors.txt)")


@pytest.mark.parametrize("prec", ["fp32", "tf32", "f16"])
def test_golden_fixture_cases(cuda, prec):
    """Outputs of the reference's unmodified InferenceToucanTTS (tests/golden/toucantts.pt, oracle/make_golden.py):
    predicted prosody, scaled prosody, external (cloner-shaped) prosody; noise from the global CPU generator."""
    from oracle import factory
    model, _ = _engine(cuda, prec)
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "toucantts.pt"))
    lines = [f"== golden {prec}"]
    failures = []
    for case in gold["cases"]:
        text = factory.make_phoneme_tensor(case["n_ph"], case["seed"])
        emb = factory.make_utterance_embedding(case["seed"])
        args = {}
        if case["gold"]:
            d, p, e = factory.make_gold_prosody(text, case["seed"])
            args = dict(durations=d.clone(), pitch=p.clone(), energy=e.clone())
        torch.manual_seed(case["noise_seed"])
        mel, dur, pitch, energy = model(text.to(cuda), utterance_embedding=emb.to(cuda), lang_id=torch.tensor([gold["lang_id"]]).to(cuda),
                                        return_duration_pitch_energy=True, **args, **case["kw"])
        dur_ok = torch.equal(dur.cpu(), case["durations"])
        line = f"case n_ph={case['n_ph']} gold={case['gold']} kw={case['kw']}: durations equal {dur_ok}"
        if dur_ok:
            err = _rel_l1(mel.cpu(), case["mel"])
            line += (f" mel rel-L1 {err:.3e} pitch {_rel_l1(pitch.cpu(), case['pitch']):.3e} "
                     f"energy {_rel_l1(energy.cpu(), case['energy']):.3e}")
            if err > MEL_TOL[prec]:
                failures.append(line)
        else:
            failures.append(line)
        lines.append(line)
    _log(lines)
    assert not failures, failures


def test_ragged_batch_equals_single_calls(cuda):
    """A ragged batch must equal per-utterance batch-1 calls (no leakage through conv halos, attention, norms)."""
    from oracle import factory
    model, _ = _engine(cuda, "tf32")
    lens = [31, 12, 20]
    texts = [factory.make_phoneme_tensor(n, 40 + i) for i, n in enumerate(lens)]
    embs = torch.stack([factory.make_utterance_embedding(40 + i) for i in range(len(lens))])
    batch = torch.zeros(len(lens), max(lens), 62)
    for i, t in enumerate(texts):
        batch[i, :lens[i]] = t
    noise = torch.randn(len(lens), 80, 600, generator=torch.Generator().manual_seed(3))
    rb = model.synthesize_batch(batch.to(cuda), torch.tensor(lens), utterance_embedding=embs.to(cuda),
                                lang_ids=torch.tensor([12, 12, 12]), noise=noise)
    for i, n in enumerate(lens):
        rs = model.synthesize_batch(texts[i].unsqueeze(0).to(cuda), torch.tensor([n]), utterance_embedding=embs[i:i + 1].to(cuda),
                                    lang_ids=torch.tensor([12]), noise=noise[i:i + 1])
        assert torch.equal(rb["durations"][i, :n].cpu(), rs["durations"][0, :n].cpu())
        m = int(rs["mel_lengths"][0])
        assert int(rb["mel_lengths"][i]) == m
        a, b = rb["mel_ncl"][i, :, :m].cpu(), rs["mel_ncl"][0, :, :m].cpu()
        assert _rel_l1(a, b) < 1e-5, f"utterance {i}: batch vs single rel-L1 {_rel_l1(a, b):.3e}"


def test_long_form_external_prosody(cuda):
    """Config-5 shaped input at reduced length: 400 phonemes with external durations/pitch/energy (UtteranceCloner
    override path, InferenceToucanTTS.py:204-212): frames = sum(edited durations), output finite."""
    from oracle import factory, restate
    model, _ = _engine(cuda, "tf32")
    n_ph = 400
    text = factory.make_phoneme_tensor(n_ph, 9)
    emb = factory.make_utterance_embedding(9)
    d, p, e = factory.make_gold_prosody(text, 9)
    ref_d, ref_p, ref_e = restate.edit_prosody(text, d, p.reshape(-1), e.reshape(-1), 1.2, 1.1, 1.2, 0.8)
    mel, dur, pitch, energy = model(text.to(cuda), durations=d.clone(), pitch=p.clone(), energy=e.clone(),
                                    utterance_embedding=emb.to(cuda), lang_id=torch.tensor([12]).to(cuda),
                                    return_duration_pitch_energy=True, duration_scaling_factor=1.1,
                                    pause_duration_scaling_factor=1.2, pitch_variance_scale=1.2, energy_variance_scale=0.8)
    assert torch.equal(dur.cpu(), ref_d)
    assert torch.allclose(pitch.cpu(), ref_p, atol=1e-5) and torch.allclose(energy.cpu(), ref_e, atol=1e-5)
    assert mel.shape == (2 * (int(ref_d.sum()) // 2), 80)
    assert torch.isfinite(mel).all()


@pytest.mark.parametrize("prec", ["fp32", "tf32", "f16"])
def test_config3_length_mel_parity(cuda, prec):
    """BASELINE.json configs[2] at its longest utterance: 200 phonemes -> ~1 000 frames.  The decoder's attention then
    runs 16 key tiles per query tile and the k = 31 depthwise conv sees full-length rows; mel rel-L1 <= 1e-3 against the
    oracle on the same flow noise, durations bit-exact (fp32) / frame count equal."""
    from oracle import factory, restate
    model, fsd = _engine(cuda, prec)
    n_ph, seed = 200, 21
    text = factory.make_phoneme_tensor(n_ph, seed)
    emb = factory.make_utterance_embedding(seed)
    with torch.inference_mode():
        ref = restate.toucantts_forward(fsd, text, emb, lang_id=12, generator=torch.Generator().manual_seed(7))
    frames = int(ref["durations"].sum())
    assert frames >= 800
    noise = torch.randn((1, 80, frames), generator=torch.Generator().manual_seed(7))
    r = model.synthesize_batch(text.unsqueeze(0).to(cuda), torch.tensor([n_ph]), utterance_embedding=emb.unsqueeze(0).to(cuda),
                               lang_ids=torch.tensor([12]), noise=noise)
    torch.cuda.synchronize()
    flips = int((r["durations"][0, :n_ph].cpu() != ref["durations"]).sum())
    _log([f"== config-3 length, {prec}: {n_ph} phonemes, {frames} frames, duration flips {flips}"])
    if prec == "fp32":
        assert flips == 0
    if flips == 0:
        mel = r["mel_ncl"][0, :, :2 * (frames // 2)].t().cpu()
        err = _rel_l1(mel, ref["mel"])
        _log([f"   mel rel-L1 {err:.3e}"])
        assert err <= MEL_TOL[prec]
    else:   # tf32 encoder feeding the fp32 duration head: a log-duration on a .5 boundary may flip by one frame
        assert flips <= 2 and abs(int(r["frames"][0]) - frames) <= 2


def test_long_form_1000_phonemes_mel_parity(cuda):
    """Config-5 shaped input: 1 000 phonemes with external durations / pitch / energy (UtteranceCloner override path,
    InferenceToucanTTS.py:204-212) -> ~3 000 frames; the mel itself is checked against the oracle (fp32 mode), not
    only durations and finiteness.  The relative positional table is regrown past its initial 256 positions."""
    from oracle import factory, restate
    model, fsd = _engine(cuda, "fp32")
    n_ph = 1000
    text = factory.make_phoneme_tensor(n_ph, 31)
    emb = factory.make_utterance_embedding(31)
    d, p, e = factory.make_gold_prosody(text, 31)
    d = d.clamp(max=4)
    with torch.inference_mode():
        ref = restate.toucantts_forward(fsd, text, emb, lang_id=12, durations=d.clone(), pitch=p.clone(), energy=e.clone(),
                                        generator=torch.Generator().manual_seed(3))
    frames = int(ref["durations"].sum())
    assert frames >= 2000
    noise = torch.randn((1, 80, frames), generator=torch.Generator().manual_seed(3))
    r = model.synthesize_batch(text.unsqueeze(0).to(cuda), torch.tensor([n_ph]), utterance_embedding=emb.unsqueeze(0).to(cuda),
                               lang_ids=torch.tensor([12]), gold_durations=d.unsqueeze(0).to(cuda),
                               gold_pitch=p.reshape(1, -1, 1).to(cuda), gold_energy=e.reshape(1, -1, 1).to(cuda), noise=noise)
    torch.cuda.synchronize()
    assert torch.equal(r["durations"][0, :n_ph].cpu(), ref["durations"])
    mel = r["mel_ncl"][0, :, :2 * (frames // 2)].t().cpu()
    err = _rel_l1(mel, ref["mel"])
    _log([f"== long form: {n_ph} phonemes, {frames} frames, mel rel-L1 {err:.3e}"])
    assert err <= 1e-3


def test_all_zero_durations_rescue_and_tiny_utterance(cuda):
    """LengthRegulator.py:52-53: an utterance whose durations sum to 0 gets one frame per phoneme; a 3-phoneme
    utterance (shorter than every conv kernel) still goes through the whole path."""
    from oracle import factory, restate
    model, fsd = _engine(cuda, "fp32")
    text = factory.make_phoneme_tensor(6, 21)
    emb = factory.make_utterance_embedding(21)
    zeros = torch.zeros(6, dtype=torch.int64)
    p = torch.ones(6, 1)
    noise = torch.randn((1, 80, 6), generator=torch.Generator().manual_seed(1))
    r = model.synthesize_batch(text.unsqueeze(0).to(cuda), torch.tensor([6]), gold_durations=zeros.unsqueeze(0),
                               gold_pitch=p.unsqueeze(0), gold_energy=p.unsqueeze(0), utterance_embedding=emb.unsqueeze(0).to(cuda),
                               lang_ids=torch.tensor([12]), noise=noise)
    assert int(r["frames"][0]) == 6 and r["durations"][0, :6].tolist() == [1] * 6
    with torch.inference_mode():
        ref = restate.toucantts_forward(fsd, text, emb, lang_id=12, durations=zeros.clone(), pitch=p.clone(), energy=p.clone(),
                                        noise=noise[0])
    assert _rel_l1(r["mel_ncl"][0, :, :6].t().cpu(), ref["mel"]) < 1e-3

    text3 = factory.make_phoneme_tensor(3, 22)
    with torch.inference_mode():
        ref3 = restate.toucantts_forward(fsd, text3, emb, lang_id=12, generator=torch.Generator().manual_seed(2))
    f3 = int(ref3["durations"].sum())
    noise3 = torch.randn((1, 80, f3), generator=torch.Generator().manual_seed(2))
    r3 = model.synthesize_batch(text3.unsqueeze(0).to(cuda), torch.tensor([3]), utterance_embedding=emb.unsqueeze(0).to(cuda),
                                lang_ids=torch.tensor([12]), noise=noise3)
    assert torch.equal(r3["durations"][0, :3].cpu(), ref3["durations"])
    m = 2 * (f3 // 2)
    assert _rel_l1(r3["mel_ncl"][0, :, :m].t().cpu(), ref3["mel"]) < 1e-3


def test_single_speaker_single_language_variant(cuda):
    """ToucanTTSInterface.py:53-62 falls back to lang_embs=None / utt_embed_dim=None checkpoints: predictors then use
    plain LayerNorm (VariancePredictor.py:42-47) and the encoder has no embedding projection."""
    import ims_toucan_prosody_variance_b200 as tb
    from ims_toucan_prosody_variance_b200 import layouts
    from oracle import factory, restate
    lay, alias = layouts.toucantts_layout(utt_embed_dim=None, lang_embs=None)
    full = factory.make_state_dict("toucantts", 1234)
    g = torch.Generator().manual_seed(11)
    sd = {}
    for k, shape in lay.items():
        if k in alias:
            sd[k] = sd[alias[k]]
        elif k in full and tuple(full[k].shape) == tuple(shape):
            sd[k] = full[k].clone()
        elif k.endswith("norms.0.weight") or ".norms." in k and k.endswith(".weight"):
            sd[k] = torch.ones(shape) + 0.1 * torch.randn(shape, generator=g)
        else:
            sd[k] = 0.1 * torch.randn(shape, generator=g)
    model = tb.ToucanTTS(weights=sd, utt_embed_dim=None, lang_embs=None, precision="fp32").to(cuda)
    model.store_inverse_all()
    text = factory.make_phoneme_tensor(17, 8)
    fsd = restate.fold_weight_norm(sd)
    with torch.inference_mode():
        ref = restate.toucantts_forward(fsd, text, None, lang_id=None, generator=torch.Generator().manual_seed(5))
    frames = int(ref["durations"].sum())
    assert frames >= 2
    noise = torch.randn((1, 80, frames), generator=torch.Generator().manual_seed(5))
    r = model.synthesize_batch(text.unsqueeze(0).to(cuda), torch.tensor([17]), noise=noise)
    assert torch.equal(r["durations"][0, :17].cpu(), ref["durations"])
    m = 2 * (frames // 2)
    assert _rel_l1(r["mel_ncl"][0, :, :m].t().cpu(), ref["mel"]) < 1e-3


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_cuda_graph_replay_matches_eager(cuda, prec):
    """enable_cuda_graphs(): both device segments are captured per padded shape (phonemes to 8, frames to 64) and
    replayed.  Same inputs -> same integers and (padding only changes masked positions) the same mel as the eager path;
    a second batch of the same shape re-uses the capture; a different shape captures anew; results of the eager path
    are untouched by switching graphs on and off."""
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory
    model = tb.ToucanTTS(weights=factory.make_state_dict("toucantts", 1234), precision=prec).to(cuda)
    model.store_inverse_all()

    def batch(lens, seed):
        t = torch.zeros((len(lens), max(lens), 62))
        for i, n in enumerate(lens):
            t[i, :n] = factory.make_phoneme_tensor(n, seed + i)
        emb = torch.stack([factory.make_utterance_embedding(seed + i) for i in range(len(lens))])
        return t.to(cuda), torch.tensor(lens, dtype=torch.int32), emb.to(cuda), torch.full((len(lens),), 12)

    def run(args, noise):
        t, tl, emb, lang = args
        r = model.synthesize_batch(t, tl, utterance_embedding=emb, lang_ids=lang, noise=noise)
        torch.cuda.synchronize()
        n = [int(v) for v in r["mel_lengths"].cpu()]
        return ([r["mel_ncl"][i, :, :n[i]].clone() for i in range(len(n))], r["durations"][:, :t.shape[1]].clone(),
                r["frames_host"].clone())

    cases = [batch([29, 12, 5], 40), batch([27, 29, 3], 50), batch([41, 40], 60)]
    noises = [torch.randn((len(c[1]), 80, 600), generator=torch.Generator().manual_seed(7 + i)) for i, c in enumerate(cases)]
    eager = [run(c, n) for c, n in zip(cases, noises)]
    model.enable_cuda_graphs(max_cached=2)
    for rep in range(2):                      # second round: every shape is a replay (case 0 was evicted: LRU of 2 -> recapture)
        for c, n, e in zip(cases, noises, eager):
            mels, dur, frames = run(c, n)
            assert torch.equal(frames, e[2]) and torch.equal(dur, e[1])
            for got, ref in zip(mels, e[0]):
                assert got.shape == ref.shape
                assert _rel_l1(got.cpu(), ref.cpu()) < (1e-6 if prec == "fp32" else 1e-4)
    assert len(model._graphs) == 2
    model.disable_cuda_graphs()
    mels, dur, frames = run(cases[0], noises[0])
    assert all(torch.equal(a, b) for a, b in zip(mels, eager[0][0]))


def test_duration_flip_rate_of_tensor_core_modes(cuda):
    """Durations are round(exp(log_d) - 1) of the duration predictor's output.  The kernel that rounds is bit-exact,
    but in the tf32 / f16 modes the encoder feeding the predictor runs on the tensor cores, so a log-duration that
    lands within ~1e-3 of a rounding boundary can round the other way.  Measure how often, against the engine's own
    fp32 mode (which matches the oracle's integers in every parity test): 96 utterances of 20..200 phonemes."""
    import random

    from oracle import factory
    rng = random.Random(17)
    lens = [rng.randint(20, 200) for _ in range(96)]
    text = torch.zeros((len(lens), max(lens), 62))
    for i, n in enumerate(lens):
        text[i, :n] = factory.make_phoneme_tensor(n, 900 + i)
    emb = torch.stack([factory.make_utterance_embedding(900 + i) for i in range(len(lens))])
    tl = torch.tensor(lens, dtype=torch.int32)
    lang = torch.full((len(lens),), 12)
    durs = {}
    for prec in ("fp32", "tf32", "f16"):
        model, _ = _engine(cuda, prec)
        r = model.synthesize_batch(text.to(cuda), tl, utterance_embedding=emb.to(cuda), lang_ids=lang, noise="device")
        durs[prec] = torch.cat([r["durations"][i, :n].cpu() for i, n in enumerate(lens)])
    total = durs["fp32"].numel()
    lines = [f"== duration flips vs fp32 mode over {total} phonemes"]
    for prec in ("tf32", "f16"):
        diff = (durs[prec] - durs["fp32"]).abs()
        flips = int((diff > 0).sum())
        lines.append(f"{prec}: {flips} of {total} phonemes differ ({100.0 * flips / total:.3f} %), max |delta| {int(diff.max())} frame(s), "
                     f"total frames {int(durs[prec].sum())} vs {int(durs['fp32'].sum())}")
        assert int(diff.max()) <= 1, f"{prec}: a duration moved by more than one frame"
        assert flips <= 0.01 * total, f"{prec}: {flips} of {total} durations flipped"
    _log(lines)
