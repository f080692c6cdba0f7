"""tb200_respair (one residual pair in one launch) against the same pair computed with torch CPU ops, one utterance
at a time on its own length (batch-1 semantics: zero padding of the convs and replicate padding of the anti-aliasing
filters at the utterance's own ends).  Reference loop bodies: BigVGAN AMP.py:51-60, HiFiGAN ResidualBlock.py:93-97.

Tolerance: fp16 operands (10-bit mantissa), fp32 accumulation, and the value between the two convolutions rounded to
fp16 (as in the two-launch path) -> 2.5e-3 of the output RMS; fp16 output adds its own 2^-11 rounding."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _pad8(n):
    return (n + 7) // 8 * 8


def _reference(x, lens, w1, b1, w2, b2, k, dil, snake, act1, act2, slope, out_alpha, res_beta, y0, accumulate):
    from oracle import restate
    ref = y0.clone()
    for b in range(x.shape[0]):
        n = lens[b]
        if n == 0:
            continue
        xb = x[b:b + 1, :, :n]
        a = restate.aa_snake(xb, *act1) if snake else F.leaky_relu(xb, slope)
        u = F.conv1d(a, w1, b1, dilation=dil, padding=(k - 1) // 2 * dil)
        v = restate.aa_snake(u, *act2) if snake else F.leaky_relu(u, slope)
        o = F.conv1d(v, w2, b2, padding=(k - 1) // 2)
        o = out_alpha * o + res_beta * xb
        ref[b, :, :n] = o[0] + (y0[b, :, :n] if accumulate else 0)
    return ref


def _run(cuda, B, C, K, dil, L, lens=None, snake=False, slope=0.1, out_alpha=1.0, res_beta=1.0, accumulate=False,
         half=False, seed=0, alpha_scale=0.3, tol=2.5e-3):
    from ims_toucan_prosody_variance_b200 import ops
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, L, generator=g)
    if half:
        x = x.half().float()
    w1 = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    w2 = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b1 = torch.randn(C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    act1 = (torch.randn(C, generator=g) * alpha_scale, torch.randn(C, generator=g) * 0.3)
    act2 = (torch.randn(C, generator=g) * alpha_scale, torch.randn(C, generator=g) * 0.3)
    y0 = torch.randn(B, C, L, generator=g) if accumulate else torch.zeros(B, C, L)
    if half:
        y0 = y0.half().float()
    lens = lens or [L] * B
    ref = _reference(x, lens, w1, b1, w2, b2, K, dil, snake, act1, act2, slope, out_alpha, res_beta, y0, accumulate)

    c1 = ops.ConvLayer(w1.to(cuda), b1.to(cuda), dilation=dil, padding=(K - 1) // 2 * dil, precision="f16")
    c2 = ops.ConvLayer(w2.to(cuda), b2.to(cuda), dilation=1, padding=(K - 1) // 2, precision="f16")
    pair = ops.ResPair(c1, c2, tuple(t.to(cuda) for t in act1) if snake else None,
                       tuple(t.to(cuda) for t in act2) if snake else None)
    dt = torch.float16 if half else torch.float32
    xd = torch.zeros(B, C, _pad8(L), dtype=dt, device=cuda)
    xd[:, :, :L] = x.to(cuda)
    yd = torch.zeros(B, C, _pad8(L), dtype=dt, device=cuda)
    yd[:, :, :L] = y0.to(cuda)
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    pair(xd, lt, yd, l_max=L, slope=slope, out_alpha=out_alpha, res_beta=res_beta, accumulate=accumulate)
    torch.cuda.synchronize()
    got = yd[:, :, :L].float().cpu()
    if half:
        tol = max(tol, 1.5e-3)
    for b in range(B):
        n = lens[b]
        if n:
            scale = ref[b, :, :n].pow(2).mean().sqrt().item() + 1e-6
            err = (got[b, :, :n] - ref[b, :, :n]).abs().max().item()
            rel = ((got[b, :, :n] - ref[b, :, :n]).pow(2).mean().sqrt() / scale).item()
            assert rel < tol, f"b={b} rel rms err {rel:.3e} (max {err:.3e}, rms {scale:.3e})"
            assert err / scale < tol * 10, f"b={b} max err {err:.3e} vs rms {scale:.3e}"
        # positions past the utterance's length must be left untouched
        assert torch.equal(got[b, :, n:], y0[b, :, n:])
    return got


@pytest.mark.parametrize("C", [32, 64, 128])
@pytest.mark.parametrize("K,dil", [(3, 1), (7, 3), (11, 5)])
def test_leaky_pair(cuda, C, K, dil):
    # (C = 128 with an fp32 stream does not fit in shared memory for K = 11: the generators run it on fp16 streams)
    _run(cuda, B=3, C=C, K=K, dil=dil, L=700, lens=[700, 333, 5], seed=C + K, half=C == 128)


@pytest.mark.parametrize("C", [32, 64, 128])
@pytest.mark.parametrize("K,dil", [(3, 1), (7, 3), (11, 5)])
def test_snake_pair(cuda, C, K, dil):
    _run(cuda, B=3, C=C, K=K, dil=dil, L=700, lens=[700, 333, 5], snake=True, seed=C + K + 1, half=C == 128)


@pytest.mark.parametrize("snake", [False, True])
def test_pair_mrf_epilogue_and_fp16_stream(cuda, snake):
    # the multi-receptive-field mean folded into the last pair of a block: y = y_old + (x + pair(x)) / 3
    _run(cuda, B=2, C=64, K=7, dil=5, L=900, lens=[900, 411], snake=snake, out_alpha=1 / 3, res_beta=1 / 3, accumulate=True)
    _run(cuda, B=2, C=32, K=11, dil=3, L=1300, lens=[1300, 129], snake=snake, half=True)
    _run(cuda, B=2, C=128, K=3, dil=5, L=300, lens=[300, 17], snake=snake, half=True, out_alpha=1 / 3, res_beta=1 / 3,
         accumulate=True)


@pytest.mark.parametrize("snake", [False, True])
def test_pair_ragged_many_tiles(cuda, snake):
    # several tiles per utterance, tile boundaries inside and at the ends of utterances, empty utterances,
    # CTAs whose tiles are all skipped (followed by a second launch on the same SMs)
    lens = [4000, 0, 1, 127, 128, 129, 2047, 3999, 12, 0, 950, 234, 235, 468]
    for _ in range(2):
        _run(cuda, B=len(lens), C=32, K=3, dil=1, L=4000, lens=lens, snake=snake, seed=5)
    _run(cuda, B=len(lens), C=64, K=11, dil=5, L=4000, lens=lens, snake=snake, seed=6)


def test_snake_pair_large_alpha(cuda):
    # trained checkpoints reach |alpha| of 2-3: sin arguments of tens of radians (sin.approx range reduction)
    _run(cuda, B=1, C=32, K=3, dil=1, L=600, snake=True, alpha_scale=1.5, seed=9, tol=4e-3)


@pytest.mark.parametrize("snake", [False, True])
def test_pair_batch_stride_beyond_32_bits(cuda, snake):
    """Utterances more than 2^31 elements apart (a 256-utterance text->wave bucket at the C = 32 stage is 3.1 G elements):
    offsets across utterances are 64-bit.  Two utterances placed 2^31 + 64 elements apart inside one 4.3 GB buffer must
    give the same result as the same two utterances in a compact tensor."""
    from ims_toucan_prosody_variance_b200 import ops
    C, K, L = 32, 3, 1000
    g = torch.Generator().manual_seed(3)
    w1 = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    w2 = torch.randn(C, C, K, generator=g) / (C * K) ** 0.5
    b1, b2 = torch.randn(C, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1
    act = (torch.randn(C, generator=g) * 0.3, torch.randn(C, generator=g) * 0.3)
    c1 = ops.ConvLayer(w1.to(cuda), b1.to(cuda), dilation=1, padding=1, precision="f16")
    c2 = ops.ConvLayer(w2.to(cuda), b2.to(cuda), dilation=1, padding=1, precision="f16")
    a = tuple(t.to(cuda) for t in act) if snake else None
    pair = ops.ResPair(c1, c2, a, a)
    x = torch.randn(2, C, L, generator=g).half().to(cuda)
    lens = torch.tensor([L, 777], dtype=torch.int32, device=cuda)
    y_ref = torch.zeros_like(x)
    pair(x, lens, y_ref, l_max=L)
    far = (1 << 31) + 64
    xbuf = torch.zeros(far + C * L, dtype=torch.float16, device=cuda)
    ybuf = torch.zeros(far + C * L, dtype=torch.float16, device=cuda)
    xs = xbuf.as_strided((2, C, L), (far, L, 1))
    ys = ybuf.as_strided((2, C, L), (far, L, 1))
    xs.copy_(x)
    pair(xs, lens, ys, l_max=L)
    torch.cuda.synchronize()
    assert torch.equal(ys, y_ref)
    del xbuf, ybuf
    torch.cuda.empty_cache()


def test_pair_rejects_unsupported(cuda):
    from ims_toucan_prosody_variance_b200 import ops
    from ims_toucan_prosody_variance_b200._lib import EngineError
    w = torch.randn(64, 64, 3, device=cuda)
    c1 = ops.ConvLayer(w, None, dilation=1, padding=1, precision="f16")
    c2 = ops.ConvLayer(w, None, dilation=1, padding=1, precision="f16")
    pair = ops.ResPair(c1, c2)
    x = torch.zeros(1, 64, 256, device=cuda)
    with pytest.raises(EngineError):
        pair(x, None, x)          # in place
    with pytest.raises(EngineError):
        pair(x[:, :, 1:], None, torch.zeros(1, 64, 255, device=cuda))   # misaligned rows
