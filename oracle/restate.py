"""Functional fp32 restatement of the reference hot path (TEST INFRASTRUCTURE).

Plain torch CPU ops driven directly by a reference-layout ``state_dict``; one
utterance at a time, exactly like the reference (its inference API is strictly
batch-1, InferenceToucanTTS.py:293-316, InferenceAvocodo.py:72,
InferenceBigVGAN.py:73).  File:line citations are relative to /root/reference.

This file is the checker for the CUDA engine and the "port" CPU baseline of
bench.py.  It is pinned against the live reference in
tests/test_oracle_vs_reference.py and through tests/golden (see oracle/__init__).
"""
import math

import torch
import torch.nn.functional as F

from oracle.alias_free_torch import kaiser_sinc_filter1d

# -----------------------------------------------------------------------------------------
# load-time folding (A16: store_inverse_all / remove_weight_norm)
# -----------------------------------------------------------------------------------------


def fold_weight_norm(sd):
    """weight = g * v / ||v|| with the norm over every dim but 0 -- what
    torch.nn.utils.remove_weight_norm leaves behind (InferenceToucanTTS.py:321-330,
    InferenceAvocodo.py:82-89, InferenceBigVGAN.py:97-105)."""
    out = {}
    for key, value in sd.items():
        if key.endswith(".weight_g"):
            v = sd[key[:-2] + "_v"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(value.shape)
            out[key[:-2]] = v * (value / norm)
        elif key.endswith(".weight_v"):
            continue
        else:
            out[key] = value
    return out


def invconv_inverse(sd, prefix):
    """Glow.py:130-139: W = P (L*mask + I)(U*mask^T + diag(sign_s * exp(log_s))), cached inverse."""
    l = sd[prefix + "l"] * sd[prefix + "l_mask"] + sd[prefix + "eye"]
    u = sd[prefix + "u"] * sd[prefix + "l_mask"].transpose(0, 1).contiguous() \
        + torch.diag(sd[prefix + "sign_s"] * torch.exp(sd[prefix + "log_s"]))
    weight = torch.matmul(sd[prefix + "p"], torch.matmul(l, u))
    return torch.inverse(weight.float())


# -----------------------------------------------------------------------------------------
# K7 (A7 tail, A8, A10): durations, prosody edits, length regulation -- integer exact
# -----------------------------------------------------------------------------------------


def durations_from_log(log_d):
    """DurationPredictor.py:79: clamp(round(exp(x) - 1), min=0).long(); round = half-to-even."""
    return torch.clamp(torch.round(log_d.exp() - 1.0), min=0).long()


def edit_prosody(text, durations, pitch, energy, pause_scale=1.0, duration_scale=1.0,
                 pitch_variance_scale=1.0, energy_variance_scale=1.0):
    """InferenceToucanTTS.py:214-227 on one utterance.  text (T,62); durations (T,) int64;
    pitch/energy (T,) fp32.  Returns new tensors (the reference mutates in place)."""
    from oracle.factory import FEAT_PHONEME, FEAT_SILENCE, FEAT_VOICED, FEAT_WORD_BOUNDARY
    durations, pitch, energy = durations.clone(), pitch.clone(), energy.clone()
    pitch[text[:, FEAT_VOICED] == 0] = 0.0
    energy[text[:, FEAT_PHONEME] == 0] = 0.0
    durations[text[:, FEAT_WORD_BOUNDARY] == 1] = 0
    if pause_scale != 1.0:
        sil = text[:, FEAT_SILENCE] == 1
        durations[sil] = torch.round(durations[sil].float() * pause_scale).long()
    if duration_scale != 1.0:
        durations = torch.round(durations.float() * duration_scale).long()
    pitch = scale_variance(pitch, pitch_variance_scale)
    energy = scale_variance(energy, energy_variance_scale)
    return durations, pitch, energy


def scale_variance(seq, scale):
    """InferenceToucanTTS.py:333-343."""
    if scale == 1.0:
        return seq
    average = seq[seq != 0.0].mean()
    seq = (seq - average) * scale + average
    return torch.where(seq < 0.0, torch.zeros_like(seq), seq)


def length_regulate(x, durations):
    """LengthRegulator.py:37-61 for one utterance: x (T,D), durations (T,) -> (sum d, D).
    The "all zero -> 1" rescue (:52-53) fires when the whole batch sums to 0."""
    if durations.sum() == 0:
        durations = torch.ones_like(durations)
    return torch.repeat_interleave(x, durations, dim=0), durations


def frame_to_phoneme(durations):
    """Index form of repeat_interleave: frame f <- phoneme searchsorted(cumsum(d), f, right)."""
    return torch.repeat_interleave(torch.arange(durations.numel()), durations)


# -----------------------------------------------------------------------------------------
# Conformer (A2-A6, A11)
# -----------------------------------------------------------------------------------------


def rel_positional_table(t, d_model=192):
    """PositionalEncoding.py:95-130: row k of the (2t-1, d) table is the sinusoid of relative
    position t-1-k."""
    position = torch.arange(t - 1, -t, -1, dtype=torch.float32).unsqueeze(1)
    # reference builds positive / negative halves separately with sin(-x), cos(-x); identical values
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(2 * t - 1, d_model)
    pos = torch.arange(0, t, dtype=torch.float32).unsqueeze(1)
    pe_pos = torch.zeros(t, d_model)
    pe_neg = torch.zeros(t, d_model)
    pe_pos[:, 0::2] = torch.sin(pos * div_term)
    pe_pos[:, 1::2] = torch.cos(pos * div_term)
    pe_neg[:, 0::2] = torch.sin(-1 * pos * div_term)
    pe_neg[:, 1::2] = torch.cos(-1 * pos * div_term)
    pe = torch.cat([torch.flip(pe_pos, [0]), pe_neg[1:]], dim=0)
    del position
    return pe


def layer_norm(x, sd, prefix, eps=1e-12):
    """LayerNorm.py:17 (eps 1e-12)."""
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], eps)


def feed_forward(x, sd, prefix):
    """MultiLayeredConv1d.py:40-51, kernel size 1: Linear-ReLU-Linear over channels."""
    h = torch.relu(F.linear(x, sd[prefix + "w_1.weight"].squeeze(-1), sd[prefix + "w_1.bias"]))
    return F.linear(h, sd[prefix + "w_2.weight"].squeeze(-1), sd[prefix + "w_2.bias"])


def rel_pos_attention(x, pos, sd, prefix, heads=4):
    """Attention.py:159-198 + rel_shift :138-157 + forward_attention :66-92, batch-1, no mask
    effect (all keys valid).  x (T,D), pos (2T-1,D)."""
    t, d = x.shape
    dk = d // heads
    q = F.linear(x, sd[prefix + "linear_q.weight"], sd[prefix + "linear_q.bias"]).view(t, heads, dk)
    k = F.linear(x, sd[prefix + "linear_k.weight"], sd[prefix + "linear_k.bias"]).view(t, heads, dk)
    v = F.linear(x, sd[prefix + "linear_v.weight"], sd[prefix + "linear_v.bias"]).view(t, heads, dk)
    p = F.linear(pos, sd[prefix + "linear_pos.weight"]).view(2 * t - 1, heads, dk)
    q_u = (q + sd[prefix + "pos_bias_u"]).transpose(0, 1)  # (H,T,dk)
    q_v = (q + sd[prefix + "pos_bias_v"]).transpose(0, 1)
    ac = torch.matmul(q_u, k.permute(1, 2, 0))  # (H,T,T)
    bd = torch.matmul(q_v, p.permute(1, 2, 0))  # (H,T,2T-1)
    # rel_shift: out[i][j] = bd[i][T-1-i+j]
    idx = (t - 1 - torch.arange(t).unsqueeze(1)) + torch.arange(t).unsqueeze(0)
    bd = torch.gather(bd, 2, idx.unsqueeze(0).expand(heads, t, t))
    attn = torch.softmax((ac + bd) / math.sqrt(dk), dim=-1)
    ctx = torch.matmul(attn, v.transpose(0, 1))  # (H,T,dk)
    ctx = ctx.transpose(0, 1).reshape(t, d)
    return F.linear(ctx, sd[prefix + "linear_out.weight"], sd[prefix + "linear_out.bias"])


def conv_module(x, sd, prefix):
    """Convolution.py:31-55: pointwise 192->384, GLU, depthwise k, BatchNorm(eval), Swish, pointwise."""
    h = x.transpose(0, 1).unsqueeze(0)  # (1,C,T)
    h = F.conv1d(h, sd[prefix + "pointwise_conv1.weight"], sd[prefix + "pointwise_conv1.bias"])
    h = F.glu(h, dim=1)
    w = sd[prefix + "depthwise_conv.weight"]
    h = F.conv1d(h, w, sd[prefix + "depthwise_conv.bias"], padding=(w.shape[-1] - 1) // 2, groups=w.shape[0])
    h = F.batch_norm(h, sd[prefix + "norm.running_mean"], sd[prefix + "norm.running_var"],
                     sd[prefix + "norm.weight"], sd[prefix + "norm.bias"], training=False, eps=1e-5)
    h = h * torch.sigmoid(h)
    h = F.conv1d(h, sd[prefix + "pointwise_conv2.weight"], sd[prefix + "pointwise_conv2.bias"])
    return h.squeeze(0).transpose(0, 1)


def conformer_block(x, pos, sd, prefix):
    """EncoderLayer.py:62-144 (macaron, normalize_before, conv module, no concat)."""
    x = x + 0.5 * feed_forward(layer_norm(x, sd, prefix + "norm_ff_macaron."), sd, prefix + "feed_forward_macaron.")
    x = x + rel_pos_attention(layer_norm(x, sd, prefix + "norm_mha."), pos, sd, prefix + "self_attn.")
    x = x + conv_module(layer_norm(x, sd, prefix + "norm_conv."), sd, prefix + "conv_module.")
    x = x + 0.5 * feed_forward(layer_norm(x, sd, prefix + "norm_ff."), sd, prefix + "feed_forward.")
    return layer_norm(x, sd, prefix + "norm_final.")


def conformer(x, sd, prefix, n_blocks=6, taps=None):
    """Conformer.py:116-123 after the input layer: x*sqrt(D), relative PE, blocks."""
    t, d = x.shape
    x = x * math.sqrt(d)
    pos = rel_positional_table(t, d)
    for i in range(n_blocks):
        x = conformer_block(x, pos, sd, f"{prefix}encoders.{i}.")
        if taps is not None:
            taps[f"{prefix}encoders.{i}"] = x
    return x


def encoder(text, utt_emb_n, lang_id, sd, taps=None):
    """Conformer.py:92-134 with the articulatory embedding (InferenceToucanTTS.py:86).
    utt_emb_n is the already once-normalised embedding (InferenceToucanTTS.py:202); it is
    normalised a second time at Conformer.py:132."""
    x = F.linear(text, sd["encoder.embed.0.weight"], sd["encoder.embed.0.bias"])
    x = F.linear(torch.tanh(x), sd["encoder.embed.2.weight"], sd["encoder.embed.2.bias"])
    if lang_id is not None:
        x = x + sd["encoder.language_embedding.weight"][int(lang_id)]
    x = conformer(x, sd, "encoder.", taps=taps)
    x = layer_norm(x, sd, "encoder.output_norm.")
    if utt_emb_n is not None:
        e = F.normalize(utt_emb_n.unsqueeze(0)).expand(x.shape[0], -1)
        x = F.linear(torch.cat([x, e], dim=-1), sd["encoder.hs_emb_projection.weight"],
                     sd["encoder.hs_emb_projection.bias"])
    return x


# -----------------------------------------------------------------------------------------
# variance predictors (A7)
# -----------------------------------------------------------------------------------------


def _cln_mlp(e, sd, prefix):
    h = torch.tanh(F.linear(e, sd[prefix + "0.weight"], sd[prefix + "0.bias"]))
    h = torch.tanh(F.linear(h, sd[prefix + "2.weight"], sd[prefix + "2.bias"]))
    return F.linear(h, sd[prefix + "4.weight"], sd[prefix + "4.bias"])


def conditional_layer_norm(x, e, sd, prefix):
    """ConditionalLayerNorm.py:52-67: divides by the VARIANCE, no epsilon.  x (T,C)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return _cln_mlp(e, sd, prefix + "W_scale.") * ((x - mean) / var) + _cln_mlp(e, sd, prefix + "W_bias.")


def variance_predictor(x, e, sd, prefix, n_layers):
    """VariancePredictor.py:53-80 / DurationPredictor.py:63-77: (T,192) -> (T,) pre-activation."""
    h = x
    for i in range(n_layers):
        w = sd[f"{prefix}conv.{i}.0.weight"]
        h = F.conv1d(h.transpose(0, 1).unsqueeze(0), w, sd[f"{prefix}conv.{i}.0.bias"],
                     padding=(w.shape[-1] - 1) // 2).squeeze(0).transpose(0, 1)
        h = torch.relu(h)
        h = conditional_layer_norm(h, e, sd, f"{prefix}norms.{i}.") if e is not None else \
            F.layer_norm(h, (h.shape[-1],), sd[f"{prefix}norms.{i}.weight"], sd[f"{prefix}norms.{i}.bias"], 1e-12)
    return F.linear(h, sd[prefix + "linear.weight"], sd[prefix + "linear.bias"]).squeeze(-1)


# -----------------------------------------------------------------------------------------
# PostNet (A12) and Glow PostFlow (A13)
# -----------------------------------------------------------------------------------------


def postnet(mel, sd):
    """PostNet.py:62-74: mel (F,80) -> residual (F,80)."""
    h = mel.transpose(0, 1).unsqueeze(0)
    for i in range(5):
        w = sd[f"conv_postnet.postnet.{i}.0.weight"]
        h = F.conv1d(h, w, None, padding=(w.shape[-1] - 1) // 2)
        h = F.group_norm(h, 32 if i < 4 else 20, sd[f"conv_postnet.postnet.{i}.1.weight"],
                         sd[f"conv_postnet.postnet.{i}.1.bias"], 1e-5)
        if i < 4:
            h = torch.tanh(h)
    return h.squeeze(0).transpose(0, 1)


def squeeze2(x):
    """glow_utils.py:28-40 with n_sqz=2 on (C,T): out[s*C+c][tau] = x[c][2 tau + s]; odd tail dropped."""
    c, t = x.shape
    t2 = t // 2
    return x[:, :2 * t2].reshape(c, t2, 2).permute(2, 0, 1).reshape(2 * c, t2)


def unsqueeze2(x):
    """glow_utils.py:43-53: inverse of squeeze2."""
    c2, t2 = x.shape
    return x.reshape(2, c2 // 2, t2).permute(1, 2, 0).reshape(c2 // 2, 2 * t2)


def wavenet(h, cond, sd, prefix, n_layers=4, hidden=192):
    """wavenet.py:89-122 (nonpadding == 1): returns the skip sum (hidden, T)."""
    out = torch.zeros_like(h)
    for i in range(n_layers):
        w = sd[f"{prefix}in_layers.{i}.weight"]
        a = F.conv1d(h.unsqueeze(0), w, sd[f"{prefix}in_layers.{i}.bias"], padding=(w.shape[-1] - 1) // 2).squeeze(0)
        a = a + cond[i * 2 * hidden:(i + 1) * 2 * hidden]
        acts = torch.tanh(a[:hidden]) * torch.sigmoid(a[hidden:])
        rs = F.conv1d(acts.unsqueeze(0), sd[f"{prefix}res_skip_layers.{i}.weight"],
                      sd[f"{prefix}res_skip_layers.{i}.bias"]).squeeze(0)
        if i < n_layers - 1:
            h = h + rs[:hidden]
            out = out + rs[hidden:]
        else:
            out = out + rs
    return out


def glow_reverse(mel, enc_up, noise, sd, n_blocks=18, taps=None):
    """Glow.forward(infer=True) + _forward(reverse=True) (Glow.py:342-391).
    mel (F,80) refined spectrogram, enc_up (F,192), noise (80,F) = randn*0.8 drawn by the caller
    from the CPU generator with shape (1,80,F) (Glow.py:363).  sd must be weight-norm folded.
    Returns (2*floor(F/2), 80)."""
    g = torch.cat([mel.transpose(0, 1), enc_up.transpose(0, 1)], dim=0).unsqueeze(0)
    g = F.conv1d(g, sd["post_flow.g_proj.weight"], sd["post_flow.g_proj.bias"], padding=2).squeeze(0)
    x = squeeze2(noise)
    g = squeeze2(g)
    half = x.shape[0] // 2
    for b in reversed(range(n_blocks)):
        an, ic, cp = f"post_flow.flows.{3 * b}.", f"post_flow.flows.{3 * b + 1}.", f"post_flow.flows.{3 * b + 2}."
        # coupling^-1 (Glow.py:248-269)
        x0, x1 = x[:half], x[half:]
        h = F.conv1d(x0.unsqueeze(0), sd[cp + "start.weight"], sd[cp + "start.bias"]).squeeze(0)
        cond = F.conv1d(g.unsqueeze(0), sd[cp + "wn.cond_layer.weight"], sd[cp + "wn.cond_layer.bias"]).squeeze(0)
        out = wavenet(h, cond, sd, cp + "wn.")
        out = F.conv1d(out.unsqueeze(0), sd[cp + "end.weight"], sd[cp + "end.bias"]).squeeze(0)
        x = torch.cat([x0, (x1 - out[:half]) * torch.exp(-out[half:])], dim=0)
        # invconv^-1 (Glow.py:93-128): channel c = a*80 + 2m + r mixes over q = 2a + r
        c, t = x.shape
        w_inv = invconv_inverse(sd, ic)
        xg = x.reshape(2, c // 4, 2, t).permute(0, 2, 1, 3).reshape(4, c // 4, t)
        xg = torch.einsum("pq,qmt->pmt", w_inv, xg)
        x = xg.reshape(2, 2, c // 4, t).permute(0, 2, 1, 3).reshape(c, t)
        # actnorm^-1 (Glow.py:30-32)
        x = (x - sd[an + "bias"].reshape(-1, 1)) * torch.exp(-sd[an + "logs"].reshape(-1, 1))
        if taps is not None:
            taps[f"post_flow.block{b}"] = x
    return unsqueeze2(x).transpose(0, 1)


# -----------------------------------------------------------------------------------------
# ToucanTTS.forward for one utterance (A1)
# -----------------------------------------------------------------------------------------


def toucantts_forward(sd_folded, text, utterance_embedding, lang_id=None, durations=None, pitch=None,
                      energy=None, duration_scaling_factor=1.0, pitch_variance_scale=1.0,
                      energy_variance_scale=1.0, pause_duration_scaling_factor=1.0, noise=None,
                      generator=None, taps=None):
    """InferenceToucanTTS.py:183-319 for one utterance.  text (T,62).  Returns a dict with
    mel (F',80), durations (T,), pitch (T,), energy (T,), decoded (F,80), log_durations.
    noise: optional (80,F) *unscaled* standard normal; else drawn like Glow.py:363 from
    `generator` (or the global CPU RNG)."""
    sd = sd_folded
    e = F.normalize(utterance_embedding.unsqueeze(0)).squeeze(0) if utterance_embedding is not None else None
    enc = encoder(text, e, lang_id, sd, taps=taps)
    e1 = e.unsqueeze(0) if e is not None else None
    out = {}
    if pitch is None:
        pitch = variance_predictor(enc, e1, sd, "pitch_predictor.", 7)
    else:
        pitch = pitch.reshape(-1).float()
    if energy is None:
        energy = variance_predictor(enc, e1, sd, "energy_predictor.", 2)
    else:
        energy = energy.reshape(-1).float()
    if durations is None:
        log_d = variance_predictor(enc, e1, sd, "duration_predictor.", 3)
        out["log_durations"] = log_d
        durations = durations_from_log(log_d)
    durations, pitch, energy = edit_prosody(text, durations, pitch, energy, pause_duration_scaling_factor,
                                            duration_scaling_factor, pitch_variance_scale, energy_variance_scale)
    # pitch_embed / energy_embed: Conv1d(1,192,k=1)  (InferenceToucanTTS.py:230-232)
    enriched = enc + pitch.unsqueeze(1) * sd["pitch_embed.0.weight"].reshape(1, -1) + sd["pitch_embed.0.bias"] \
        + energy.unsqueeze(1) * sd["energy_embed.0.weight"].reshape(1, -1) + sd["energy_embed.0.bias"]
    up, used_d = length_regulate(enriched, durations)
    dec = conformer(up, sd, "decoder.", taps=taps)
    decoded = F.linear(dec, sd["feat_out.weight"], sd["feat_out.bias"])
    refined = decoded + postnet(decoded, sd)
    frames = refined.shape[0]
    if noise is None:
        noise = torch.randn((1, 80, frames), generator=generator).squeeze(0)
    mel = glow_reverse(refined, up, noise * 0.8, sd, taps=taps)
    if taps is not None:
        taps.update(encoder=enc, enriched=enriched, upsampled=up, decoder=dec, decoded=decoded, refined=refined)
    out.update(mel=mel, durations=durations, pitch=pitch, energy=energy, decoded=decoded)
    return out


# -----------------------------------------------------------------------------------------
# vocoders (A14, A15)
# -----------------------------------------------------------------------------------------

UPSAMPLE_SCALES = (8, 6, 4, 2)
UPSAMPLE_KERNELS = (16, 12, 8, 4)
RESBLOCK_KERNELS = (3, 7, 11)
RESBLOCK_DILATIONS = (1, 3, 5)


def hifigan_forward(sd, mel, taps=None):
    """InferenceAvocodo.py:69-80 + ResidualBlock.py:83-98.  sd weight-norm folded; mel (80,F) or
    (B,80,F) -> wave (F*384,) or (B,F*384)."""
    single = mel.dim() == 2
    c = mel.unsqueeze(0) if single else mel
    c = F.conv1d(c, sd["input_conv.weight"], sd["input_conv.bias"], padding=3)
    for i, (u, k) in enumerate(zip(UPSAMPLE_SCALES, UPSAMPLE_KERNELS)):
        c = F.conv_transpose1d(F.leaky_relu(c, 0.1), sd[f"upsamples.{i}.1.weight"], sd[f"upsamples.{i}.1.bias"],
                               stride=u, padding=(k - u) // 2)
        if taps is not None:
            taps[f"up{i}"] = c
        cs = 0.0
        for j, kr in enumerate(RESBLOCK_KERNELS):
            x = c
            blk = f"blocks.{i * 3 + j}."
            for n, d in enumerate(RESBLOCK_DILATIONS):
                xt = F.conv1d(F.leaky_relu(x, 0.1), sd[f"{blk}convs1.{n}.1.weight"], sd[f"{blk}convs1.{n}.1.bias"],
                              dilation=d, padding=(kr - 1) // 2 * d)
                xt = F.conv1d(F.leaky_relu(xt, 0.1), sd[f"{blk}convs2.{n}.1.weight"], sd[f"{blk}convs2.{n}.1.bias"],
                              padding=(kr - 1) // 2)
                x = xt + x
            cs = cs + x
        c = cs / 3
        if taps is not None:
            taps[f"stage{i}"] = c
    c = F.conv1d(F.leaky_relu(c, 0.01), sd["output_conv.1.weight"], sd["output_conv.1.bias"], padding=3)
    c = torch.tanh(c).squeeze(1)
    return c.squeeze(0) if single else c


_AA_FILTER = None


def aa_filter():
    global _AA_FILTER
    if _AA_FILTER is None:
        _AA_FILTER = kaiser_sinc_filter1d(0.25, 0.3, 12)
    return _AA_FILTER


def snake_beta(x, alpha, beta):
    """Snake.py:56-69 with alpha_logscale=True: x + 1/(e^beta + 1e-9) * sin^2(x e^alpha)."""
    a = torch.exp(alpha).reshape(1, -1, 1)
    b = torch.exp(beta).reshape(1, -1, 1)
    return x + (1.0 / (b + 0.000000001)) * torch.pow(torch.sin(x * a), 2)


def aa_snake(x, alpha, beta):
    """alias_free_torch.Activation1d(SnakeBeta): 2x kaiser-sinc upsample, snake, 2x downsample."""
    c = x.shape[1]
    f = aa_filter().expand(c, -1, -1)
    y = F.pad(x, (5, 5), mode="replicate")
    y = 2 * F.conv_transpose1d(y, f, stride=2, groups=c)[..., 15:-15]
    y = snake_beta(y, alpha, beta)
    y = F.pad(y, (5, 6), mode="replicate")
    return F.conv1d(y, f, stride=2, groups=c)


def bigvgan_forward(sd, mel, taps=None):
    """InferenceBigVGAN.py:72-95 + AMP.py:51-60.  sd weight-norm folded."""
    single = mel.dim() == 2
    x = mel.unsqueeze(0) if single else mel
    x = F.conv1d(x, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(UPSAMPLE_SCALES, UPSAMPLE_KERNELS)):
        x = F.conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"], stride=u, padding=(k - u) // 2)
        if taps is not None:
            taps[f"up{i}"] = x
        xs = None
        for j, kr in enumerate(RESBLOCK_KERNELS):
            blk = f"resblocks.{i * 3 + j}."
            y = x
            for n, d in enumerate(RESBLOCK_DILATIONS):
                a1, a2 = f"{blk}activations.{2 * n}.act.", f"{blk}activations.{2 * n + 1}.act."
                xt = aa_snake(y, sd[a1 + "alpha"], sd[a1 + "beta"])
                xt = F.conv1d(xt, sd[f"{blk}convs1.{n}.weight"], sd[f"{blk}convs1.{n}.bias"], dilation=d,
                              padding=int((kr * d - d) / 2))
                xt = aa_snake(xt, sd[a2 + "alpha"], sd[a2 + "beta"])
                xt = F.conv1d(xt, sd[f"{blk}convs2.{n}.weight"], sd[f"{blk}convs2.{n}.bias"], padding=int((kr - 1) / 2))
                y = xt + y
            xs = y if xs is None else xs + y
        x = xs / 3
        if taps is not None:
            taps[f"stage{i}"] = x
    x = aa_snake(x, sd["activation_post.act.alpha"], sd["activation_post.act.beta"])
    x = torch.tanh(F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)).squeeze(1)
    return x.squeeze(0) if single else x


# -----------------------------------------------------------------------------------------
# metrics used by every parity test
# -----------------------------------------------------------------------------------------


_PREVIOUS_PHONE_MODIFIERS = {
    "\u02D0": "lengthened", "\u02D1": "half-length", "\u0306": "shortened", "\u0303": "nasal",
    "\u02E5": "very-high-tone", "\u02E6": "high-tone", "\u02E7": "mid-tone", "\u02E8": "low-tone", "\u02E9": "very-low-tone",
    "\u2B67": "rising-tone", "\u2B68": "falling-tone", "\u2B81": "peaking-tone", "\u2B83": "dipping-tone",
}


def string_to_tensor(phones, phone_to_vector, feature_to_index, handle_missing=True):
    """Preprocessing/TextFrontend.py:213-288 with input_phonemes=True, character by character: primary stress marks
    the next appended vector (:232-234, :282-284), 13 modifier characters set one feature of the previous vector
    (:236-274), everything else is looked up (:276-280); an unknown character is skipped when handle_missing but
    still resolves a pending stress mark (the `if stressed_flag` of :282 sits after the try / except)."""
    phones = phones.replace("\u025A", "\u0259").replace("\u1D7B", "\u0268")          # :223
    vectors = []
    stressed = False
    for char in phones:
        if char == "\u02C8":
            stressed = True
        elif char in _PREVIOUS_PHONE_MODIFIERS:
            vectors[-1][feature_to_index[_PREVIOUS_PHONE_MODIFIERS[char]]] = 1
        else:
            if handle_missing:
                if char in phone_to_vector:
                    vectors.append(list(phone_to_vector[char]))
            else:
                vectors.append(list(phone_to_vector[char]))
            if stressed:
                stressed = False
                vectors[-1][feature_to_index["stressed"]] = 1
    return torch.tensor(vectors, dtype=torch.float32).reshape(len(vectors), -1)


def float2pcm(sig):
    """Utility/utils.py:20-33 (dtype int16): (sig * 32768).clip(-32768, 32767) truncated toward zero by astype."""
    import numpy as np
    sig = np.asarray(sig)
    if sig.dtype.kind != "f":
        raise TypeError("'sig' must be a float array")
    return (sig * 32768.0 + 0).clip(-32768, 32767).astype(np.int16)


def rel_l1(a, b):
    """mean |a-b| / mean |b| (north_star: mel relative L1 <= 1e-3 in the fp32-accumulate mode)."""
    return float((a.double() - b.double()).abs().mean() / b.double().abs().mean().clamp_min(1e-30))


def snr_db(test, ref):
    """10 log10(sum ref^2 / sum (test-ref)^2) (north_star: waveform SNR >= 40 dB)."""
    num = float((ref.double() ** 2).sum())
    den = float(((test.double() - ref.double()) ** 2).sum())
    return float("inf") if den == 0 else 10.0 * math.log10(num / max(den, 1e-300))
