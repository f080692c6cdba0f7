"""Deterministic random-init state_dicts and synthetic inputs (TEST INFRASTRUCTURE).

The reference needs weights at construction (InferenceToucanTTS.py:75,180,
InferenceAvocodo.py:67, InferenceBigVGAN.py:70) and there is no network for
checkpoints, so every parity test and the benchmark run on random-init weights
with the reference's state_dict layout (oracle/manifests/*.json, SURVEY.md
appendix B).  Values are a pure function of (model, seed, key) so the live
reference (authoring container), the restated oracle and the CUDA engine (GPU
box) all load bit-identical tensors without shipping 250 MB of weights.

Differences from the reference's own init, on purpose:
  * vocoder convs: the reference draws N(0, 0.01) (HiFiGAN.py:125-137,
    BigVGAN.py init_weights), which makes every residual branch ~1e-2 of the
    skip path -- a kernel bug inside a ResBlock would hide below the SNR bound.
    Here branches are O(1) of the skip path so SNR actually tests them.
  * PostFlow ``end`` convs are zero-init in the reference (Glow.py:241-243, the
    flow is an identity coupling); here they are small but non-zero.
  * biases, norm affine terms, BatchNorm running stats, Snake alpha/beta and the
    ConditionalLayerNorm MLPs are perturbed so none of them is a no-op.
  * ``duration_predictor.linear`` is calibrated (weight*0.02, bias 1.8) exactly as
    SURVEY.md section 8d prescribes -> about 4-5 frames per phoneme.
"""
import json
import math
import os
import random
import zlib

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_MANIFESTS = {}


def manifest(model):
    if model not in _MANIFESTS:
        with open(os.path.join(_HERE, "manifests", model + ".json")) as f:
            _MANIFESTS[model] = json.load(f)
    return _MANIFESTS[model]


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def _randn(shape, g, std=1.0, mean=0.0):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean


def _uniform(shape, g, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _fans(shape, transposed=False, stride=1):
    if len(shape) == 2:
        return shape[1], shape[0]
    k = shape[2] if len(shape) > 2 else 1
    if transposed:  # ConvTranspose1d weight is (Cin, Cout, k)
        return shape[0] * k // max(stride, 1), shape[1] * k
    return shape[1] * k, shape[0] * k


_UP_STRIDES = (8, 6, 4, 2)


def _vocoder_tensor(key, shape, g):
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "filter":  # alias_free_torch kaiser-sinc buffers (persistent in 0.0.6)
        from oracle.alias_free_torch import kaiser_sinc_filter1d
        return kaiser_sinc_filter1d(0.25, 0.3, 12)
    if leaf in ("alpha", "beta"):
        return _randn(shape, g, 0.3)
    if leaf == "bias":
        return _randn(shape, g, 0.05)
    transposed = key.startswith("ups.") or key.startswith("upsamples.")
    stride = _UP_STRIDES[int(key.split(".")[1])] if transposed else 1
    if leaf in ("weight_v", "weight"):
        fan_in, _ = _fans(shape, transposed, stride)
        if ".convs1." in key or ".convs2." in key:
            gain = 0.8  # residual branch ~0.35x (HiFiGAN) / ~0.75x (BigVGAN) of the skip path per block
        elif key.startswith(("conv_pre", "input_conv")):
            gain = 0.3  # mel input is ~N(-5, 2): keep stage activations O(1)
        elif key.startswith("output_conv"):
            gain = 0.6  # keep tanh out of saturation (wave rms ~0.15)
        elif key.startswith("conv_post"):
            gain = 0.08
        else:
            gain = 1.0
        return _uniform(shape, g, gain * math.sqrt(3.0 / fan_in))
    raise KeyError(key)


def _toucan_tensor(key, shape, g):
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "running_mean":
        return _randn(shape, g, 0.1)
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    if key.startswith("post_flow.flows."):
        if leaf == "logs":
            return _randn(shape, g, 0.1)
        if leaf == "l_mask":
            return torch.tril(torch.ones(shape), -1)
        if leaf == "eye":
            return torch.eye(shape[0])
        if leaf == "p":
            perm = torch.randperm(shape[0], generator=g)
            return torch.eye(shape[0])[perm]
        if leaf == "sign_s":
            return (torch.rand(shape, generator=g) > 0.5).float() * 2 - 1
        if leaf == "log_s":
            return _randn(shape, g, 0.1)
        if leaf in ("l", "u") and len(shape) == 2 and shape[0] == shape[1] == 4:
            return _randn(shape, g, 0.3)
        if ".end." in key and leaf == "weight":
            return _randn(shape, g, 0.05 / math.sqrt(shape[1]))
    if ".norms." in key:  # ConditionalLayerNorm MLPs (ConditionalLayerNorm.py:23-51)
        if leaf == "weight":
            return _randn(shape, g, 0.3 / math.sqrt(shape[1]))
        return _randn(shape, g, 0.05, 1.0 if ".W_scale." in key else 0.0)
    if key == "encoder.language_embedding.weight":
        return _randn(shape, g, 1.0)
    if key == "duration_predictor.linear.bias":
        return torch.full(shape, 1.8)
    if len(shape) >= 2 and leaf in ("weight", "weight_v", "pos_bias_u", "pos_bias_v"):
        fan_in, fan_out = _fans(shape)
        w = _uniform(shape, g, math.sqrt(6.0 / (fan_in + fan_out)))
        if key == "duration_predictor.linear.weight":
            w = w * 0.02
        return w
    if leaf == "weight":  # 1-D affine terms of LayerNorm / GroupNorm / BatchNorm
        return _randn(shape, g, 0.1, 1.0)
    if leaf == "bias":
        return _randn(shape, g, 0.05)
    raise KeyError(key)


def make_state_dict(model, seed=1234):
    """model in {"toucantts", "hifigan", "bigvgan"} -> reference-layout state_dict (fp32 / int64)."""
    man = manifest(model)
    rule = _toucan_tensor if model == "toucantts" else _vocoder_tensor
    sd = {}
    for key, meta in man.items():
        if "alias" in meta:
            continue
        if key.endswith("weight_g"):
            continue
        sd[key] = rule(key, tuple(meta["shape"]), _gen(seed, key)).reshape(meta["shape"]).contiguous()
    for key, meta in man.items():  # weight-norm gains: ||v|| per dim-0 slice, perturbed
        if key.endswith("weight_g") and "alias" not in meta:
            v = sd[key[:-1] + "v"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(meta["shape"])
            sd[key] = (norm * (1.0 + _randn(tuple(meta["shape"]), _gen(seed, key), 0.1))).contiguous()
    for key, meta in man.items():
        if "alias" in meta:
            sd[key] = sd[meta["alias"]]
    return {k: sd[k] for k in man}


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------
# Indices into the 62-dim articulatory vector, from the reference's
# Preprocessing/articulatory_features.py:817-901 get_feature_to_index_lookup().
FEAT_PHONEME, FEAT_SILENCE, FEAT_EOS, FEAT_WORD_BOUNDARY, FEAT_VOICED = 15, 16, 17, 21, 61
N_FEATS = 62


def _phone_rows():
    """A small bank of 0/1 articulatory rows with the four control features set the way the
    frontend's table sets them (phoneme rows, unvoiced phoneme rows, silence '~', word boundary
    ' ', end-of-sentence '#').  Generated, not copied from the reference table."""
    rng = random.Random(7)
    rows = {}

    def base():
        r = [0.0] * N_FEATS
        for i in range(22, 61):
            r[i] = float(rng.random() < 0.2)
        return r

    voiced = []
    for _ in range(24):
        r = base(); r[FEAT_PHONEME] = 1.0; r[FEAT_VOICED] = 1.0; voiced.append(r)
    unvoiced = []
    for _ in range(12):
        r = base(); r[FEAT_PHONEME] = 1.0; unvoiced.append(r)
    sil = [0.0] * N_FEATS; sil[FEAT_SILENCE] = 1.0
    wb = [0.0] * N_FEATS; wb[FEAT_WORD_BOUNDARY] = 1.0
    eos = [0.0] * N_FEATS; eos[FEAT_EOS] = 1.0
    rows.update(voiced=voiced, unvoiced=unvoiced, sil=sil, wb=wb, eos=eos)
    return rows


_ROWS = None


def make_phoneme_tensor(n_phonemes, seed):
    """(T,62) float 0/1 tensor shaped like the frontend output: starts with '~', ends '~','#',
    word boundaries every few phonemes (TextFrontend.py:431-441 conventions)."""
    global _ROWS
    if _ROWS is None:
        _ROWS = _phone_rows()
    rng = random.Random(seed)
    seq = [_ROWS["sil"]]
    since_wb = 0
    while len(seq) < n_phonemes - 2:
        if since_wb >= 3 and rng.random() < 0.22:
            seq.append(_ROWS["wb"]); since_wb = 0
        elif rng.random() < 0.03:
            seq.append(_ROWS["sil"]); since_wb = 0
        else:
            seq.append(rng.choice(_ROWS["voiced"] if rng.random() < 0.7 else _ROWS["unvoiced"]))
            since_wb += 1
    seq = seq[:max(n_phonemes - 2, 1)] + [_ROWS["sil"], _ROWS["eos"]]
    return torch.tensor(seq[:n_phonemes] if n_phonemes >= 3 else seq[-n_phonemes:], dtype=torch.float32)


def make_utterance_embedding(seed):
    return _randn((64,), _gen(seed, "utt_emb"))


def make_mel(batch, frames, seed, n_mels=80):
    """log-mel-like synthetic vocoder input: randn*2-5."""
    return _randn((batch, n_mels, frames), _gen(seed, "mel"), 2.0, -5.0)


def make_gold_prosody(text, seed):
    """Cloner-shaped external prosody: int64 durations in [0,12) with
    zeros at word boundaries; |N(1,0.2)| pitch/energy, zero where unvoiced / non-phoneme."""
    g = _gen(seed, "prosody")
    t = text.shape[0]
    dur = torch.randint(0, 12, (t,), generator=g, dtype=torch.int64)
    dur[text[:, FEAT_WORD_BOUNDARY] == 1] = 0
    pitch = _randn((t, 1), g, 0.2, 1.0).abs()
    energy = _randn((t, 1), g, 0.2, 1.0).abs()
    pitch[text[:, FEAT_VOICED] == 0] = 0.0
    energy[text[:, FEAT_PHONEME] == 0] = 0.0
    return dur, pitch, energy
