"""tb200_conv1d against torch's CPU conv on the same seeded inputs (all three precisions).

Tolerances: fp32 SIMT 2e-5 relative to the output scale; tf32 / f16 operands have a 10-bit
mantissa -> 2e-3 of the output RMS (accumulation is fp32 in TMEM)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "tf32": 2e-3, "f16": 2e-3}


def _run(cuda, prec, B, Cin, Cout, K, dil, L, lens=None, up=0, act=0, slope=0.0, snake=False, residual=False,
         out_act=0, out_alpha=1.0, res_beta=1.0, accumulate=False, x_half=False, y_half=False, seed=0,
         y_misalign=False, alpha_scale=0.3, staged=False):
    from ims_toucan_prosody_variance_b200 import ops
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, L, generator=g)
    if x_half:
        x = x.half().float()
    wshape = (Cin, Cout, 2 * up) if up else (Cout, Cin, K)
    w = torch.randn(wshape, generator=g) / (Cin * (2 if up else K)) ** 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    alpha = torch.randn(Cin, generator=g) * alpha_scale
    beta = torch.randn(Cin, generator=g) * 0.3
    Lout = L * up if up else L
    res = torch.randn(B, Cout, Lout, generator=g) if residual else None
    y0 = torch.randn(B, Cout, Lout, generator=g) if accumulate else torch.zeros(B, Cout, Lout)
    lens = lens or [L] * B
    pad = (K - 1) // 2 * dil

    # reference, one utterance at a time on its own length (batch-1 semantics)
    from oracle import restate
    ref = y0.clone()
    for b in range(B):
        n = lens[b]
        if n == 0:
            continue
        xb = x[b:b + 1, :, :n]
        if snake:
            xb = restate.aa_snake(xb, alpha, beta)
        elif act == 1:
            xb = F.leaky_relu(xb, slope)
        if up:
            o = F.conv_transpose1d(xb, w, bias, stride=up, padding=up // 2)
        else:
            o = F.conv1d(xb, w, bias, dilation=dil, padding=pad)
        if out_act == 1:
            o = torch.tanh(o)
        o = o * out_alpha
        no = n * up if up else n
        if residual:
            o = o + res_beta * res[b:b + 1, :, :no]
        ref[b, :, :no] = o[0] + (y0[b, :, :no] if accumulate else 0)

    layer = ops.ConvLayer(w.to(cuda), bias.to(cuda), dilation=dil, padding=pad, transposed_stride=up, precision=prec)
    xd = x.to(cuda)
    if x_half:
        xd = xd.half()
    yd = y0.to(cuda)
    if y_half:
        yd = yd.half()
    if staged:   # 16-byte aligned rows, as the generators' workspaces have
        def pad8(t):
            full = torch.zeros(t.shape[0], t.shape[1], (t.shape[2] + 7) // 8 * 8, dtype=t.dtype, device=cuda)
            full[:, :, :t.shape[2]] = t
            return full[:, :, :t.shape[2]]
        xd, yd = pad8(xd), pad8(yd)
    if y_misalign:   # a view whose rows start one element off: no 8-byte (fp32) / 4-byte (fp16) alignment for paired stores
        full = torch.zeros(B, Cout, Lout + 2 + (Lout & 1), dtype=yd.dtype, device=cuda)
        full[:, :, 1:1 + Lout] = yd
        yd = full[:, :, 1:1 + Lout]
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    resd = res.to(cuda) if residual else None
    if staged and residual:
        resd = pad8(resd.half() if y_half else resd)
        if y_half:
            ref = ref + 0   # (the fp16 rounding of the residual is inside the tolerance of fp16 outputs)
    if staged:
        assert layer.staged_ok(xd, yd, resd, 2 if snake else act, out_act), "the staged kernel must accept this case"
    layer(xd, lt, yd, act=2 if snake else act, slope=slope, alpha=alpha.to(cuda) if snake else None,
          beta=beta.to(cuda) if snake else None, out_act=out_act, out_alpha=out_alpha,
          residual=resd, res_beta=res_beta, accumulate=accumulate, staged=staged)
    torch.cuda.synchronize()
    got = yd.float().cpu()
    tol = max(TOL[prec], 1e-3) if y_half else TOL[prec]  # fp16 output rounding: 2^-11 relative
    for b in range(B):
        no = lens[b] * up if up else lens[b]
        if no == 0:
            continue
        err = (got[b, :, :no] - ref[b, :, :no]).abs().max().item()
        scale = ref[b, :, :no].pow(2).mean().sqrt().item() + 1e-6
        assert err / scale < tol * 8, f"{prec} b={b} max err {err:.3e} vs rms {scale:.3e}"
        rel = ((got[b, :, :no] - ref[b, :, :no]).pow(2).mean().sqrt() / scale).item()
        assert rel < tol, f"{prec} b={b} rel rms err {rel:.3e}"
        # positions past the utterance's length must be left untouched
        assert torch.equal(got[b, :, no:], (y0.half().float() if y_half else y0)[b, :, no:])


@pytest.mark.parametrize("prec", ["fp32", "f16", "tf32"])
def test_pointwise_linear(cuda, prec):
    _run(cuda, prec, B=2, Cin=64, Cout=48, K=1, dil=1, L=300)


@pytest.mark.parametrize("prec", ["fp32", "f16", "tf32"])
def test_dilated_conv_residual_ragged(cuda, prec):
    _run(cuda, prec, B=3, Cin=32, Cout=32, K=11, dil=5, L=400, lens=[400, 131, 7], act=1, slope=0.1, residual=True)


@pytest.mark.parametrize("prec", ["fp32", "f16", "tf32"])
def test_conv_k7_odd_channels(cuda, prec):
    _run(cuda, prec, B=2, Cin=80, Cout=512, K=7, dil=1, L=150, lens=[150, 90])


@pytest.mark.parametrize("prec", ["fp32", "f16", "tf32"])
@pytest.mark.parametrize("u,cin,cout", [(8, 64, 32), (6, 32, 16), (2, 64, 32)])
def test_transposed(cuda, prec, u, cin, cout):
    _run(cuda, prec, B=2, Cin=cin, Cout=cout, K=2 * u, dil=1, L=200, lens=[200, 77], up=u, act=1, slope=0.1)


@pytest.mark.parametrize("prec", ["fp32", "f16", "tf32"])
def test_aa_snake_prologue(cuda, prec):
    _run(cuda, prec, B=2, Cin=32, Cout=32, K=3, dil=3, L=300, lens=[300, 45], snake=True, residual=True)


def test_aa_snake_trained_scale_alpha(cuda):
    """Trained BigVGAN checkpoints reach |alpha| of 2-3 (log scale): sin arguments of tens of radians.  The staging
    uses sin.approx (range reduction in fp32, absolute error growing with |x e^alpha|): still within the f16 tolerance."""
    _run(cuda, "f16", B=2, Cin=32, Cout=32, K=3, dil=1, L=300, lens=[300, 64], snake=True, alpha_scale=1.5, seed=4)
    _run(cuda, "f16", B=1, Cin=64, Cout=64, K=7, dil=3, L=600, snake=True, alpha_scale=1.5, x_half=True, seed=5)


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_epilogue_variants(cuda, prec):
    _run(cuda, prec, B=2, Cin=32, Cout=16, K=7, dil=1, L=260, lens=[260, 129], act=1, slope=0.01, out_act=1)
    _run(cuda, prec, B=2, Cin=64, Cout=64, K=3, dil=1, L=260, lens=[260, 128], residual=True, out_alpha=1 / 3,
         res_beta=1 / 3, accumulate=True)
    _run(cuda, prec, B=1, Cin=64, Cout=64, K=3, dil=1, L=260, x_half=True, y_half=True)


@pytest.mark.parametrize("prec", ["f16", "tf32"])
def test_streamed_weights_wide(cuda, prec):
    # 256x256x11 does not fit in shared memory: exercises the weight ring and multi-chunk K loop
    _run(cuda, prec, B=2, Cin=256, Cout=256, K=11, dil=1, L=300, lens=[300, 200], act=1, slope=0.1)


@pytest.mark.parametrize("prec", ["f16"])
def test_transposed_wide_multi_ntile(cuda, prec):
    # N = 256*8 = 2048 -> 8 accumulator tiles per input tile
    _run(cuda, prec, B=1, Cin=512, Cout=256, K=16, dil=1, L=140, up=8)


def test_single_tap_tiny(cuda):
    for prec in ("fp32", "f16", "tf32"):
        _run(cuda, prec, B=1, Cin=16, Cout=16, K=1, dil=1, L=5)


@pytest.mark.parametrize("prec", ["fp32", "f16"])
def test_zero_length_utterance_in_batch(cuda, prec):
    # an empty utterance inside a batch is skipped: nothing read, nothing written
    _run(cuda, prec, B=3, Cin=32, Cout=32, K=7, dil=3, L=200, lens=[200, 0, 64], act=1, slope=0.1, residual=True)
    _run(cuda, prec, B=3, Cin=32, Cout=32, K=3, dil=1, L=200, lens=[0, 200, 3], snake=True)


def test_snake_edges_short_and_unaligned_lengths(cuda):
    # utterance lengths around the 8-step block size and the filter reach of the anti-aliased snake
    for n in (1, 2, 5, 7, 8, 9, 15, 17, 31, 33):
        _run(cuda, "f16", B=2, Cin=32, Cout=32, K=3, dil=1, L=40, lens=[40, n], snake=True, seed=n)
    _run(cuda, "f16", B=2, Cin=64, Cout=64, K=11, dil=5, L=700, lens=[700, 513], snake=True, residual=True)


@pytest.mark.parametrize("prec", ["f16", "tf32"])
def test_wide_linear_relu_and_scaled_residual(cuda, prec):
    # the acoustic model's FFN pair: 192 -> 1536 with ReLU, 1536 -> 192 scaled by 0.5 plus residual
    from ims_toucan_prosody_variance_b200 import ops
    from ims_toucan_prosody_variance_b200._lib import OUT_RELU
    g = torch.Generator().manual_seed(4)
    B, L, lens = 3, 200, [200, 77, 130]
    x = torch.randn(B, 192, L, generator=g)
    w1, b1 = torch.randn(1536, 192, generator=g) / 192 ** 0.5, torch.randn(1536, generator=g) * 0.1
    w2, b2 = torch.randn(192, 1536, generator=g) / 1536 ** 0.5, torch.randn(192, generator=g) * 0.1
    l1 = ops.ConvLayer(w1.to(cuda), b1.to(cuda), precision=prec)
    l2 = ops.ConvLayer(w2.to(cuda), b2.to(cuda), precision=prec)
    xd = x.to(cuda)
    h = torch.zeros(B, 1536, L, device=cuda)
    y = xd.clone()
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    l1(xd, lt, h, out_act=OUT_RELU)
    l2(h, lt, y, out_alpha=0.5, residual=y)
    torch.cuda.synchronize()
    for b, n in enumerate(lens):
        hb = torch.relu(torch.einsum("oc,cl->ol", w1, x[b, :, :n]) + b1[:, None])
        ref = x[b, :, :n] + 0.5 * (torch.einsum("oc,cl->ol", w2, hb) + b2[:, None])
        got = y[b, :, :n].cpu()
        rel = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        assert rel < 2e-3, f"{prec} b={b} rel rms err {rel:.3e}"
        assert torch.equal(y[b, :, n:].cpu(), x[b, :, n:])


@pytest.mark.parametrize("u,cin,cout", [(4, 64, 32), (8, 32, 16), (2, 32, 16), (6, 64, 64)])
@pytest.mark.parametrize("y_half", [False, True])
def test_transposed_store_only_epilogue(cuda, u, cin, cout, y_half):
    """Every up-sampling stride of the vocoders through the stride-templated epilogue: fp32 / fp16 outputs, ragged batch
    with a 1-frame utterance, odd lengths."""
    _run(cuda, "f16", B=3, Cin=cin, Cout=cout, K=2 * u, dil=1, L=333, lens=[333, 1, 150], up=u, y_half=y_half)


@pytest.mark.parametrize("u", [4, 8])
@pytest.mark.parametrize("y_half", [False, True])
def test_transposed_unaligned_output_rows(cuda, u, y_half):
    """Output rows that are not 8-byte aligned take the scalar-store variant of the same epilogue."""
    _run(cuda, "f16", B=2, Cin=32, Cout=16, K=2 * u, dil=1, L=130, lens=[130, 61], up=u, y_half=y_half, y_misalign=True)


def test_transposed_with_residual_keeps_generic_epilogue(cuda):
    _run(cuda, "f16", B=2, Cin=64, Cout=32, K=8, dil=1, L=150, lens=[150, 40], up=4, residual=True, res_beta=0.5)


def test_ragged_ctas_without_valid_tiles_then_relaunch(cuda):
    """Resident weights + a ragged batch in which whole CTAs own only skipped tiles (they must not leave bulk copies in
    flight when they exit), followed by more launches on the same SMs."""
    lens = [2000] + [3] * 200 + [2000]
    for seed in range(3):
        _run(cuda, "f16", B=len(lens), Cin=32, Cout=32, K=3, dil=1, L=2000, lens=lens, act=1, slope=0.1, residual=True, seed=seed)


@pytest.mark.parametrize("snake", [False, True])
@pytest.mark.parametrize("C,K,dil", [(32, 3, 1), (64, 7, 3), (64, 11, 5), (128, 11, 5), (128, 3, 1)])
def test_staged_single_conv(cuda, snake, C, K, dil):
    """tb200_conv1d_staged (one conv through the TMA-fed pipeline) == tb200_conv1d's contract: ragged, with and without
    residual / accumulate, fp32 and fp16 streams."""
    kw = dict(B=3, Cin=C, Cout=C, K=K, dil=dil, L=900, lens=[900, 257, 6], snake=snake, act=0 if snake else 1, slope=0.1,
              staged=True)
    _run(cuda, "f16", seed=C + K, x_half=True, y_half=True, **kw)                                  # conv1 of a pair
    _run(cuda, "f16", seed=C + K + 1, x_half=True, y_half=True, residual=True, **kw)               # conv2 of a pair
    _run(cuda, "f16", seed=C + K + 2, x_half=True, y_half=True, residual=True, out_alpha=1 / 3, res_beta=1 / 3,
         accumulate=True, **kw)                                                                    # ... folded into the MRF mean
    if C <= 64:
        _run(cuda, "f16", seed=C + K + 3, residual=True, **kw)                                     # fp32 streams


def test_staged_many_tiles_and_empty_utterances(cuda):
    lens = [5000, 0, 1, 127, 128, 129, 2047, 4999, 12, 0, 950, 234, 235, 468]
    for snake in (False, True):
        for _ in range(2):
            _run(cuda, "f16", B=len(lens), Cin=32, Cout=32, K=3, dil=1, L=5000, lens=lens, snake=snake, act=0 if snake else 1,
                 slope=0.1, residual=True, x_half=True, y_half=True, staged=True, seed=7)
