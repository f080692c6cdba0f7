"""K7/K12 on the GPU against the oracle: integer durations, prefix sums and frame indices bit-exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(B, Tmax, seed):
    from oracle import factory
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(3, Tmax + 1, (B,), generator=g).tolist()
    lens[0] = Tmax
    text = torch.zeros(B, Tmax, 62)
    for b, n in enumerate(lens):
        text[b, :n] = factory.make_phoneme_tensor(n, seed * 100 + b)
    return text, lens, g


@pytest.mark.parametrize("pause,scale", [(1.0, 1.0), (1.2, 1.1), (0.7, 0.9)])
def test_duration_finalize_bit_exact(cuda, pause, scale):
    from ims_toucan_prosody_variance_b200 import ops
    from oracle import restate
    B, Tmax = 9, 200
    text, lens, g = _batch(B, Tmax, 11)
    # log-durations spread over [-1, 3.2] incl. exact .5 boundaries of exp(x)-1
    logd = torch.rand(B, Tmax, generator=g) * 4.2 - 1.0
    logd[1, :8] = torch.log(torch.tensor([1.5, 2.5, 3.5, 4.5, 5.5, 6.5, 0.5, 1.0]))
    logd[2, :lens[2]] = -5.0  # all-zero utterance -> LengthRegulator rescue
    dur, cum, frames = ops.duration_finalize(text.to(cuda), torch.tensor(lens, dtype=torch.int32, device=cuda),
                                             log_dur=logd.to(cuda), pause_scale=pause, duration_scale=scale)
    flips = 0
    for b, n in enumerate(lens):
        d = restate.durations_from_log(logd[b, :n])
        d, _, _ = restate.edit_prosody(text[b, :n], d, torch.zeros(n), torch.zeros(n), pause, scale)
        _, d = restate.length_regulate(torch.zeros(n, 1), d)
        got = dur[b, :n].cpu()
        flips += int((got != d).sum())
        assert torch.equal(cum[b, :n].cpu().long(), torch.cumsum(got, 0))
        assert int(frames[b]) == int(got.sum())
    assert flips == 0, f"{flips} duration mismatches"


def test_duration_gold_and_expand_indices(cuda):
    from ims_toucan_prosody_variance_b200 import ops
    from oracle import factory, restate
    B, Tmax, C = 5, 120, 192
    text, lens, g = _batch(B, Tmax, 5)
    gold = torch.zeros(B, Tmax, dtype=torch.int64)
    pitch = torch.zeros(B, Tmax)
    energy = torch.zeros(B, Tmax)
    for b, n in enumerate(lens):
        d, p, e = factory.make_gold_prosody(text[b, :n], b)
        gold[b, :n], pitch[b, :n], energy[b, :n] = d, p[:, 0], e[:, 0]
    tl = torch.tensor(lens, dtype=torch.int32, device=cuda)
    dur, cum, frames = ops.duration_finalize(text.to(cuda), tl, gold_dur=gold.to(cuda), pause_scale=1.3, duration_scale=1.0)
    pd, ed = pitch.to(cuda), energy.to(cuda)
    ops.variance_edit(pd, text.to(cuda), tl, 0, 1.2)
    ops.variance_edit(ed, text.to(cuda), tl, 1, 0.8)
    enc = torch.randn(B, C, Tmax, generator=g)
    wp, bp, we, be = (torch.randn(C, generator=g) for _ in range(4))
    fmax = int(frames.max())
    out, f2p = ops.length_regulate(enc.to(cuda), cum, tl, frames, fmax, pd, ed, wp.to(cuda), bp.to(cuda), we.to(cuda),
                                   be.to(cuda), want_index=True)
    for b, n in enumerate(lens):
        d, p, e = restate.edit_prosody(text[b, :n], gold[b, :n], pitch[b, :n], energy[b, :n], 1.3, 1.0, 1.2, 0.8)
        assert torch.equal(dur[b, :n].cpu(), d)
        assert torch.allclose(pd[b, :n].cpu(), p, rtol=1e-5, atol=1e-6)
        assert torch.allclose(ed[b, :n].cpu(), e, rtol=1e-5, atol=1e-6)
        idx = restate.frame_to_phoneme(d)
        F_b = int(frames[b])
        assert F_b == idx.numel()
        assert torch.equal(f2p[b, :F_b].cpu().long(), idx)  # bit-exact frame indices
        enriched = enc[b, :, :n].t() + p.unsqueeze(1) * wp + bp + e.unsqueeze(1) * we + be
        ref, _ = restate.length_regulate(enriched, d)
        assert torch.allclose(out[b, :, :F_b].cpu().t(), ref, rtol=1e-5, atol=1e-5)
        assert torch.all(out[b, :, F_b:] == 0)  # masked padding


def test_expand_large_roundtrip(cuda):
    """Size-independent property at benchmark scale: expand then segment-mean recovers the rows."""
    from ims_toucan_prosody_variance_b200 import ops
    B, T, C = 16, 2000, 192
    g = torch.Generator().manual_seed(3)
    dur = torch.randint(0, 12, (B, T), generator=g)
    dur[:, 0] = 1
    cum = torch.cumsum(dur, 1).int().to(cuda)
    frames = cum[:, -1].contiguous()
    enc = torch.randn(B, C, T, generator=g).to(cuda)
    tl = torch.full((B,), T, dtype=torch.int32, device=cuda)
    out, f2p = ops.length_regulate(enc, cum, tl, frames, int(frames.max()), want_index=True)
    for b in (0, B - 1):
        F_b = int(frames[b])
        idx = f2p[b, :F_b].long()
        assert torch.equal(idx.cpu(), torch.repeat_interleave(torch.arange(T), dur[b]))
        assert torch.equal(out[b, :, :F_b], enc[b].index_select(1, idx))
        assert bool((idx[1:] >= idx[:-1]).all())
