"""The non-GEMM kernels of the acoustic model, each alone against a plain fp32 torch reference of the same reference
op on the same seeded inputs (ragged batches; positions past an utterance's length must stay untouched).

  tb200_relpos_attention  Layers/Attention.py:159-198, rel_shift :138-157, forward_attention :66-92
  tb200_glu_dwconv        Layers/Convolution.py:43-52
  tb200_group_norm        Layers/PostNet.py:45-59,62-74
  tb200_channel_norm      Layers/LayerNorm.py:17, ConditionalLayerNorm.py:52-67
  tb200_wn_gate           ToucanTTS/wavenet.py:29-35,102-111
  tb200_flow_close        ToucanTTS/Glow.py:260-263,116-128,30-32
  tb200_squeeze2          ToucanTTS/glow_utils.py:28-53
  tb200_cln_mlp           Layers/ConditionalLayerNorm.py:27-50

Tolerances: all kernels compute in fp32 -> 2e-5 of the output scale (attention: 5e-5, sums over up to 4 097 keys in a
different order than torch)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _pad4(n):
    return (n + 3) // 4 * 4


def _ncl(t, cuda, fill=0.0):
    """(B,C,L) CPU -> CUDA NCL tensor with a 16-byte aligned row pitch."""
    b, c, l = t.shape
    out = torch.full((b, c, _pad4(l)), fill, dtype=torch.float32, device=cuda)
    out[:, :, :l] = t.to(cuda)
    return out


def _close(got, ref, tol, what):
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


# ---------------------------------------------------------------------------------------------
# relative-position attention
# ---------------------------------------------------------------------------------------------
def _attention_reference(q, k, v, p, bu, bv):
    """One utterance, all keys valid.  q,k,v (H,T,dk); p (H,2T-1,dk) with row m <-> relative position T-1-m
    (the table order of RelPositionalEncoding); bu, bv (H,dk).  score(i,j) = ((q_i+u).k_j + (q_i+v).p_{i-j}) / sqrt(dk)."""
    h, t, dk = q.shape
    ac = torch.matmul(q + bu.unsqueeze(1), k.transpose(1, 2))
    bd = torch.matmul(q + bv.unsqueeze(1), p.transpose(1, 2))          # (H,T,2T-1)
    idx = (t - 1 - torch.arange(t).unsqueeze(1)) + torch.arange(t).unsqueeze(0)   # rel_shift: bd[i][T-1-i+j]
    bd = torch.gather(bd, 2, idx.unsqueeze(0).expand(h, t, t))
    attn = torch.softmax((ac + bd) / math.sqrt(dk), dim=-1)
    return torch.matmul(attn, v)                                      # (H,T,dk)


@pytest.mark.parametrize("tensor_core", [False, True])
@pytest.mark.parametrize("lens", [[1], [63], [64], [65], [127, 128, 129], [1000, 17, 129], [4097]])
def test_relpos_attention(cuda, lens, tensor_core):
    """tensor_core=False: the fp32 CUDA-core kernel (tolerance 5e-5); True: tcgen05 flash attention with fp16 operands
    (Q, K, the positional band, the probabilities and V are rounded to 11 bits: tolerance 4e-3 of the output scale)."""
    from ims_toucan_prosody_variance_b200 import ops
    heads, dk = 4, 48
    d = heads * dk
    b, l_max = len(lens), max(lens)
    g = torch.Generator().manual_seed(l_max)
    qkv = torch.randn(b, 3 * d, l_max, generator=g)
    cap = 256
    while cap < l_max:
        cap *= 2
    # projected positional table: column (cap-1-r) holds the vector of relative position r, r in (-cap, cap)
    pos = torch.randn(d, 2 * cap - 1, generator=g) * 0.5
    bu, bv = torch.randn(heads, dk, generator=g) * 0.3, torch.randn(heads, dk, generator=g) * 0.3
    qkv_d, out_d = _ncl(qkv, cuda), torch.full((b, d, _pad4(l_max)), 7.0, device=cuda)
    pos_d = torch.zeros(d, _pad4(2 * cap - 1), device=cuda)
    pos_d[:, :2 * cap - 1] = pos.to(cuda)
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    if tensor_core:
        pos16, center = ops.pack_relpos_table(pos_d[:, :2 * cap - 1], cap - 1, heads)
        ops.relpos_attention_tc(qkv_d, lt, out_d, pos16, center, bu.to(cuda), bv.to(cuda), heads, l_max)
    else:
        ops.relpos_attention(qkv_d, lt, out_d, pos_d, cap - 1, bu.to(cuda), bv.to(cuda), heads, l_max)
    torch.cuda.synchronize()
    got = out_d.cpu()
    for i, t in enumerate(lens):
        q, k, v = (qkv[i, j * d:(j + 1) * d, :t].reshape(heads, dk, t).transpose(1, 2).double() for j in range(3))
        # rows m = 0..2t-2 of the utterance's own table <-> relative positions t-1-m <-> columns cap-1-(t-1-m)
        cols = cap - 1 - (t - 1 - torch.arange(2 * t - 1))
        p = pos[:, cols].reshape(heads, dk, 2 * t - 1).transpose(1, 2).double()
        ref = _attention_reference(q, k, v, p, bu.double(), bv.double())        # (H,T,dk)
        ref = ref.transpose(1, 2).reshape(d, t).float()
        _close(got[i, :, :t], ref, 4e-3 if tensor_core else 5e-5, f"attention T={t} tensor_core={tensor_core}")
        assert torch.all(got[i, :, t:l_max] == 7.0), "rows past the utterance's length were written"


# ---------------------------------------------------------------------------------------------
# GLU + depthwise conv + BatchNorm(eval) + Swish
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [7, 31])
def test_glu_dwconv(cuda, k):
    from ims_toucan_prosody_variance_b200 import ops
    c, lens = 192, [300, 41, 1, 130]
    b, l_max = len(lens), max(lens)
    g = torch.Generator().manual_seed(k)
    x = torch.randn(b, 2 * c, l_max, generator=g)
    w, bias = torch.randn(c, 1, k, generator=g) * 0.2, torch.randn(c, generator=g) * 0.1
    mean, var = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    gamma, beta = torch.randn(c, generator=g) * 0.2 + 1.0, torch.randn(c, generator=g) * 0.1
    out_d = torch.full((b, c, _pad4(l_max)), 7.0, device=cuda)
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    ops.glu_dwconv(_ncl(x, cuda), lt, out_d, w.reshape(c, k).contiguous().to(cuda), bias.to(cuda), mean.to(cuda), var.to(cuda),
                   gamma.to(cuda), beta.to(cuda), l_max)
    torch.cuda.synchronize()
    got = out_d.cpu()
    for i, t in enumerate(lens):
        h = F.glu(x[i:i + 1, :, :t], dim=1)
        h = F.conv1d(h, w, bias, padding=(k - 1) // 2, groups=c)
        h = F.batch_norm(h, mean, var, gamma, beta, training=False, eps=1e-5)
        h = h * torch.sigmoid(h)
        _close(got[i, :, :t], h[0], 2e-5, f"glu_dwconv k={k} T={t}")
        assert torch.all(got[i, :, t:l_max] == 7.0)


# ---------------------------------------------------------------------------------------------
# GroupNorm (+ tanh, + residual), statistics over the utterance's own frames
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,groups,tanh,residual", [(256, 32, True, False), (80, 20, False, True), (80, 20, False, False)])
def test_group_norm(cuda, c, groups, tanh, residual):
    from ims_toucan_prosody_variance_b200 import ops
    lens = [500, 37, 2, 1000]
    b, l_max = len(lens), max(lens)
    g = torch.Generator().manual_seed(c + groups)
    x = torch.randn(b, c, l_max, generator=g) * 2 + 0.3
    gamma, beta = torch.randn(c, generator=g) * 0.3 + 1.0, torch.randn(c, generator=g) * 0.2
    res = torch.randn(b, c, l_max, generator=g) if residual else None
    out_d = torch.full((b, c, _pad4(l_max)), 7.0, device=cuda)
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    ops.group_norm(_ncl(x, cuda), lt, out_d, gamma.to(cuda), beta.to(cuda), groups, l_max,
                   residual=_ncl(res, cuda) if residual else None, tanh=tanh)
    torch.cuda.synchronize()
    got = out_d.cpu()
    for i, t in enumerate(lens):
        h = F.group_norm(x[i:i + 1, :, :t], groups, gamma, beta, eps=1e-5)
        if tanh:
            h = torch.tanh(h)
        if residual:
            h = h + res[i:i + 1, :, :t]
        _close(got[i, :, :t], h[0], 3e-5, f"group_norm C={c} T={t}")
        assert torch.all(got[i, :, t:l_max] == 7.0)


# ---------------------------------------------------------------------------------------------
# LayerNorm over channels / ConditionalLayerNorm (variance, not std, and no epsilon: reference quirk)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lens", [[257, 3, 64], [33, 5, 1], [1000, 999, 130, 65]])
@pytest.mark.parametrize("c", [192, 256, 100])
def test_channel_norm_both_modes(cuda, c, lens):
    # C = 100 leaves part of the per-warp channel slots empty
    from ims_toucan_prosody_variance_b200 import ops
    b, l_max = len(lens), max(lens)
    g = torch.Generator().manual_seed(c)
    x = torch.randn(b, c, l_max, generator=g) * 1.5 + 0.2
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    # mode 0: LayerNorm, eps 1e-12
    gamma, beta = torch.randn(c, generator=g) * 0.3 + 1.0, torch.randn(c, generator=g) * 0.2
    out_d = torch.full((b, c, _pad4(l_max)), 7.0, device=cuda)
    ops.channel_norm(_ncl(x, cuda), lt, out_d, gamma.to(cuda), beta.to(cuda), l_max)
    got = out_d.cpu()
    for i, t in enumerate(lens):
        ref = F.layer_norm(x[i, :, :t].t(), (c,), gamma, beta, 1e-12).t()
        _close(got[i, :, :t], ref, 2e-5, f"layer_norm C={c} T={t}")
        assert torch.all(got[i, :, t:l_max] == 7.0)
    # mode 1: ConditionalLayerNorm, per-utterance scale / bias: y = scale * (x - mean) / var + bias
    scale, bias = torch.randn(b, c, generator=g) * 0.3 + 1.0, torch.randn(b, c, generator=g) * 0.2
    out_d = torch.full((b, c, _pad4(l_max)), 7.0, device=cuda)
    ops.channel_norm(_ncl(x, cuda), lt, out_d, scale.to(cuda), bias.to(cuda), l_max, conditional=True)
    got = out_d.cpu()
    for i, t in enumerate(lens):
        xi = x[i, :, :t].t()                                           # (T,C)
        mean = xi.mean(dim=-1, keepdim=True)
        var = ((xi - mean) ** 2).mean(dim=-1, keepdim=True)
        ref = (scale[i] * ((xi - mean) / var) + bias[i]).t()
        _close(got[i, :, :t], ref, 5e-5, f"conditional_layer_norm C={c} T={t}")


# ---------------------------------------------------------------------------------------------
# WaveNet gate, flow close, squeeze / unsqueeze
# ---------------------------------------------------------------------------------------------
def test_wn_gate(cuda):
    from ims_toucan_prosody_variance_b200 import ops
    hidden, lens = 192, [333, 5, 64]
    b, l_max = len(lens), max(lens)
    a = torch.randn(b, 2 * hidden, l_max, generator=torch.Generator().manual_seed(3)) * 2
    out_d = torch.full((b, hidden, _pad4(l_max)), 7.0, device=cuda)
    ops.wn_gate(_ncl(a, cuda), torch.tensor(lens, dtype=torch.int32, device=cuda), out_d, l_max)
    got = out_d.cpu()
    for i, t in enumerate(lens):
        ref = torch.tanh(a[i, :hidden, :t]) * torch.sigmoid(a[i, hidden:, :t])
        _close(got[i, :, :t], ref, 2e-5, f"wn_gate T={t}")
        assert torch.all(got[i, :, t:l_max] == 7.0)


def test_flow_close(cuda):
    """coupling^-1, InvConvNear^-1 (channel regrouping of Glow.py:102-103,126-127) and ActNorm^-1 in one pass."""
    from ims_toucan_prosody_variance_b200 import ops
    c, lens = 160, [250, 9, 1]
    half = c // 2
    b, l_max = len(lens), max(lens)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(b, c, l_max, generator=g)
    ml = torch.randn(b, c, l_max, generator=g) * 0.3           # rows [0, C/2) m, [C/2, C) logs
    w_inv = torch.randn(4, 4, generator=g)
    an_bias, an_logs = torch.randn(c, generator=g) * 0.2, torch.randn(c, generator=g) * 0.2
    x_d = _ncl(x, cuda, fill=7.0)
    ops.flow_close(x_d, _ncl(ml, cuda), torch.tensor(lens, dtype=torch.int32, device=cuda), l_max, w_inv.contiguous().to(cuda),
                   an_bias.to(cuda), an_logs.to(cuda))
    got = x_d.cpu()
    for i, t in enumerate(lens):
        xi = x[i, :, :t]
        z = torch.cat([xi[:half], (xi[half:] - ml[i, :half, :t]) * torch.exp(-ml[i, half:, :t])], dim=0)
        zg = z.reshape(2, c // 4, 2, t).permute(0, 2, 1, 3).reshape(4, c // 4, t)
        zg = torch.einsum("pq,qmt->pmt", w_inv, zg)
        z = zg.reshape(2, 2, c // 4, t).permute(0, 2, 1, 3).reshape(c, t)
        ref = (z - an_bias.reshape(-1, 1)) * torch.exp(-an_logs.reshape(-1, 1))
        _close(got[i, :, :t], ref, 2e-5, f"flow_close T={t}")
        assert torch.equal(got[i, :, t:l_max], x[i, :, t:l_max]), "positions past the length must keep their value"


def test_squeeze2_roundtrip_and_layout(cuda):
    from ims_toucan_prosody_variance_b200 import ops
    from oracle import restate
    c, lens = 80, [101, 100, 2, 1]
    b, l_max = len(lens), max(lens)
    x = torch.randn(b, c, l_max, generator=torch.Generator().manual_seed(5))
    lt = torch.tensor(lens, dtype=torch.int32, device=cuda)
    sq_d = torch.full((b, 2 * c, _pad4(l_max // 2 + 1)), 7.0, device=cuda)
    ops.squeeze2(_ncl(x, cuda), lt, sq_d, l_max)
    got = sq_d.cpu()
    for i, t in enumerate(lens):
        ref = restate.squeeze2(x[i, :, :t])
        assert torch.equal(got[i, :, :t // 2], ref), f"squeeze T={t}"
        assert torch.all(got[i, :, t // 2:l_max // 2] == 7.0)
    # inverse: squeezed lengths; an odd last frame is gone (Glow's output length is 2 * floor(F / 2))
    back_d = torch.full((b, c, _pad4(l_max)), 7.0, device=cuda)
    ops.squeeze2(sq_d, lt // 2, back_d, l_max // 2, inverse=True)
    back = back_d.cpu()
    for i, t in enumerate(lens):
        assert torch.equal(back[i, :, :2 * (t // 2)], x[i, :, :2 * (t // 2)]), f"unsqueeze T={t}"


def test_cln_mlp(cuda):
    """All conditioning MLPs of the variance predictors in one launch: W4 tanh(W2 tanh(W0 e + b0) + b2) + b4."""
    from ims_toucan_prosody_variance_b200 import ops
    _cln_case(cuda, 24, 5)
    _cln_case(cuda, 3, 19)      # several utterance groups per MLP, the last one partial


def _cln_case(cuda, n, b):
    from ims_toucan_prosody_variance_b200 import ops
    e_dim, cc = 64, 256
    g = torch.Generator().manual_seed(2)
    e = torch.randn(b, e_dim, generator=g)
    w0, b0 = torch.randn(n, e_dim, e_dim, generator=g) / 8, torch.randn(n, e_dim, generator=g) * 0.1
    w2, b2 = torch.randn(n, cc, e_dim, generator=g) / 8, torch.randn(n, cc, generator=g) * 0.1
    w4, b4 = torch.randn(n, cc, cc, generator=g) / 16, torch.randn(n, cc, generator=g) * 0.1
    out = ops.cln_mlp(*(t.contiguous().to(cuda) for t in (e, w0, b0, w2, b2, w4, b4))).cpu()
    for i in range(n):
        ref = F.linear(torch.tanh(F.linear(torch.tanh(F.linear(e, w0[i], b0[i])), w2[i], b2[i])), w4[i], b4[i])
        _close(out[i], ref, 2e-5, f"cln_mlp {i}")
