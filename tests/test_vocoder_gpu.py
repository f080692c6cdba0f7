"""Vocoder generators on the GPU against the oracle (restatement of InferenceAvocodo / InferenceBigVGAN).

north_star tolerance: waveform SNR >= 40 dB.  The engine's default operand type is fp16 (10-bit
mantissa, fp32 accumulate); the exact fp32 mode must be far above the bound."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

SNR_MIN = {"fp32": 80.0, "f16": 40.0, "tf32": 40.0}


def _make(kind, prec, cuda, tmp_path):
    import ims_toucan_prosody_variance_b200 as tb
    from oracle import factory, restate
    sd = factory.make_state_dict(kind, 1234)
    path = os.path.join(tmp_path, kind + ".pt")
    torch.save({"generator": sd}, path)
    cls = tb.HiFiGANGenerator if kind == "hifigan" else tb.BigVGAN
    model = cls(path, precision=prec).to(cuda)
    model.remove_weight_norm()
    return model, restate.fold_weight_norm(sd)


@pytest.mark.parametrize("kind", ["hifigan", "bigvgan"])
@pytest.mark.parametrize("prec", ["fp32", "f16", "tf32"])
def test_generator_ragged_batch(cuda, tmp_path, kind, prec):
    from oracle import factory, restate
    model, fsd = _make(kind, prec, cuda, str(tmp_path))
    lens = [37, 21, 1]
    mel = factory.make_mel(len(lens), max(lens), seed=2)
    wave = model.forward_batch(mel.to(cuda), torch.tensor(lens)).cpu()
    fwd = restate.hifigan_forward if kind == "hifigan" else restate.bigvgan_forward
    for b, n in enumerate(lens):
        ref = fwd(fsd, mel[b, :, :n])
        got = wave[b, :n * 384]
        snr = restate.snr_db(got, ref)
        assert snr >= SNR_MIN[prec], f"{kind}/{prec} utterance {b}: SNR {snr:.1f} dB"


@pytest.mark.parametrize("kind", ["hifigan", "bigvgan"])
def test_generator_batch1_signature(cuda, tmp_path, kind):
    """forward(c: (80,T)) -> (T*384,) exactly like the reference's inference modules."""
    from oracle import factory, restate
    model, fsd = _make(kind, "f16", cuda, str(tmp_path))
    mel = factory.make_mel(1, 130, seed=9)[0]
    wave = model(mel.to(cuda))
    assert wave.shape == (130 * 384,) and wave.is_cuda
    ref = (restate.hifigan_forward if kind == "hifigan" else restate.bigvgan_forward)(fsd, mel)
    assert restate.snr_db(wave.cpu(), ref) >= 40.0
    assert float(wave.abs().max()) <= 1.0


def test_golden_vocoder_fixture(cuda, tmp_path):
    """Committed fixture generated from the LIVE reference modules (oracle/make_golden.py)."""
    golden = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vocoder.pt"))
    from oracle import factory, restate
    for kind in ("hifigan", "bigvgan"):
        model, _ = _make(kind, "f16", cuda, str(tmp_path))
        mel = factory.make_mel(1, golden["frames"], seed=golden["seed"])[0]
        wave = model(mel.to(cuda)).cpu()
        snr = restate.snr_db(wave, golden[kind].float())
        assert snr >= 40.0, f"{kind}: SNR vs reference fixture {snr:.1f} dB"


def test_forward_is_cuda_graph_capturable(cuda, tmp_path):
    """The C ABI enqueues on the caller's stream and never synchronises: a whole generator forward can be captured
    in a CUDA graph and replayed (workspaces are allocated by the warm-up call)."""
    from oracle import factory, restate
    model, fsd = _make("hifigan", "f16", cuda, str(tmp_path))
    mel = factory.make_mel(2, 16, seed=5).to(cuda)
    lens = torch.tensor([16, 9], dtype=torch.int32, device=cuda)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            eager = model.forward_batch(mel, lens).clone()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = model.forward_batch(mel, lens)
    mel.copy_(factory.make_mel(2, 16, seed=6).to(cuda))      # new input, same buffers
    graph.replay()
    torch.cuda.synchronize()
    ref = restate.hifigan_forward(fsd, mel[0, :, :16].cpu())
    assert restate.snr_db(out[0, :16 * 384].cpu(), ref) >= 40.0
    assert not torch.equal(out, eager)
