"""state_dict layouts of the three checkpoints the engine must load unchanged.

The compatibility contract of the drop-in modules is the reference's state_dict (SURVEY.md
appendix B): same dotted keys, same shapes, weight-norm ``weight_g``/``weight_v`` pairs before
folding, BatchNorm running stats, shared PostFlow WaveNet layers.  The layouts are generated from
the hyper-parameters (constructor defaults of InferenceToucanTTS.py:18-75,
InferenceAvocodo.py:8-22, InferenceBigVGAN.py:22-30), and ``attach`` materialises them as nested
parameter containers so ``load_state_dict`` / ``state_dict`` behave like the reference's modules.
"""
import torch

BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked", "p", "sign_s", "l_mask", "eye", "filter")


def _wn(out, name, shape):
    """weight-normalised conv: bias, weight_g (dim-0 gains), weight_v."""
    out[name + ".bias"] = None  # filled by caller
    out[name + ".weight_g"] = (shape[0], 1, 1)
    out[name + ".weight_v"] = tuple(shape)


def hifigan_layout(in_channels=80, out_channels=1, channels=512, kernel_size=7, upsample_scales=(8, 6, 4, 2),
                   upsample_kernel_sizes=(16, 12, 8, 4), resblock_kernel_sizes=(3, 7, 11),
                   resblock_dilations=((1, 3, 5), (1, 3, 5), (1, 3, 5))):
    lay = {}

    def wn(name, shape, nbias):
        lay[name + ".bias"] = (nbias,)
        lay[name + ".weight_g"] = (shape[0], 1, 1)
        lay[name + ".weight_v"] = tuple(shape)

    wn("input_conv", (channels, in_channels, kernel_size), channels)
    for i, k in enumerate(upsample_kernel_sizes):
        wn(f"upsamples.{i}.1", (channels // 2 ** i, channels // 2 ** (i + 1), k), channels // 2 ** (i + 1))
    for i in range(len(upsample_kernel_sizes)):
        ch = channels // 2 ** (i + 1)
        for j, kr in enumerate(resblock_kernel_sizes):
            blk = f"blocks.{i * len(resblock_kernel_sizes) + j}"
            for n in range(len(resblock_dilations[j])):
                wn(f"{blk}.convs1.{n}.1", (ch, ch, kr), ch)
            for n in range(len(resblock_dilations[j])):
                wn(f"{blk}.convs2.{n}.1", (ch, ch, kr), ch)
    wn("output_conv.1", (out_channels, ch, kernel_size), out_channels)
    wn("out_proj_x1", (1, 512 // 4, 7), 1)
    wn("out_proj_x2", (1, 512 // 8, 7), 1)
    return lay, {}


def bigvgan_layout(num_mels=80, upsample_initial_channel=512, upsample_rates=(8, 6, 4, 2),
                   upsample_kernel_sizes=(16, 12, 8, 4), resblock_kernel_sizes=(3, 7, 11),
                   resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), filter_buffers=True):
    lay = {}

    def wn(name, shape, nbias):
        lay[name + ".bias"] = (nbias,)
        lay[name + ".weight_g"] = (shape[0], 1, 1)
        lay[name + ".weight_v"] = tuple(shape)

    def act(name, ch):
        lay[name + ".act.alpha"] = (ch,)
        lay[name + ".act.beta"] = (ch,)
        if filter_buffers:
            lay[name + ".upsample.filter"] = (1, 1, 12)
            lay[name + ".downsample.lowpass.filter"] = (1, 1, 12)

    c0 = upsample_initial_channel
    wn("conv_pre", (c0, num_mels, 7), c0)
    for i, k in enumerate(upsample_kernel_sizes):
        wn(f"ups.{i}.0", (c0 // 2 ** i, c0 // 2 ** (i + 1), k), c0 // 2 ** (i + 1))
    for i in range(len(upsample_rates)):
        ch = c0 // 2 ** (i + 1)
        for j, kr in enumerate(resblock_kernel_sizes):
            blk = f"resblocks.{i * len(resblock_kernel_sizes) + j}"
            nl = len(resblock_dilation_sizes[j])
            for n in range(nl):
                wn(f"{blk}.convs1.{n}", (ch, ch, kr), ch)
            for n in range(nl):
                wn(f"{blk}.convs2.{n}", (ch, ch, kr), ch)
            for n in range(2 * nl):
                act(f"{blk}.activations.{n}", ch)
    act("activation_post", ch)
    wn("conv_post", (1, ch, 7), 1)
    lay["out_proj_x1.weight"] = (1, 512 // 4, 7)
    lay["out_proj_x1.bias"] = (1,)
    lay["out_proj_x2.weight"] = (1, 512 // 8, 7)
    lay["out_proj_x2.bias"] = (1,)
    return lay, {}


def _conformer_layout(lay, prefix, adim, heads, units, blocks, dw_kernel):
    for i in range(blocks):
        p = f"{prefix}encoders.{i}."
        lay[p + "self_attn.pos_bias_u"] = (heads, adim // heads)
        lay[p + "self_attn.pos_bias_v"] = (heads, adim // heads)
        for name in ("linear_q", "linear_k", "linear_v", "linear_out"):
            lay[p + f"self_attn.{name}.weight"] = (adim, adim)
            lay[p + f"self_attn.{name}.bias"] = (adim,)
        lay[p + "self_attn.linear_pos.weight"] = (adim, adim)
        for ff in ("feed_forward", "feed_forward_macaron"):
            lay[p + ff + ".w_1.weight"] = (units, adim, 1)
            lay[p + ff + ".w_1.bias"] = (units,)
            lay[p + ff + ".w_2.weight"] = (adim, units, 1)
            lay[p + ff + ".w_2.bias"] = (adim,)
        lay[p + "conv_module.pointwise_conv1.weight"] = (2 * adim, adim, 1)
        lay[p + "conv_module.pointwise_conv1.bias"] = (2 * adim,)
        lay[p + "conv_module.depthwise_conv.weight"] = (adim, 1, dw_kernel)
        lay[p + "conv_module.depthwise_conv.bias"] = (adim,)
        lay[p + "conv_module.norm.weight"] = (adim,)
        lay[p + "conv_module.norm.bias"] = (adim,)
        lay[p + "conv_module.norm.running_mean"] = (adim,)
        lay[p + "conv_module.norm.running_var"] = (adim,)
        lay[p + "conv_module.norm.num_batches_tracked"] = ()
        lay[p + "conv_module.pointwise_conv2.weight"] = (adim, adim, 1)
        lay[p + "conv_module.pointwise_conv2.bias"] = (adim,)
        for name in ("norm_ff", "norm_mha", "norm_ff_macaron", "norm_conv", "norm_final"):
            lay[p + name + ".weight"] = (adim,)
            lay[p + name + ".bias"] = (adim,)


def _predictor_layout(lay, prefix, idim, layers, chans, kernel, utt):
    for i in range(layers):
        lay[f"{prefix}conv.{i}.0.weight"] = (chans, idim if i == 0 else chans, kernel)
        lay[f"{prefix}conv.{i}.0.bias"] = (chans,)
    for i in range(layers):
        if utt is None:
            lay[f"{prefix}norms.{i}.weight"] = (chans,)
            lay[f"{prefix}norms.{i}.bias"] = (chans,)
            continue
        for w in ("W_scale", "W_bias"):
            lay[f"{prefix}norms.{i}.{w}.0.weight"] = (utt, utt)
            lay[f"{prefix}norms.{i}.{w}.0.bias"] = (utt,)
            lay[f"{prefix}norms.{i}.{w}.2.weight"] = (chans, utt)
            lay[f"{prefix}norms.{i}.{w}.2.bias"] = (chans,)
            lay[f"{prefix}norms.{i}.{w}.4.weight"] = (chans, chans)
            lay[f"{prefix}norms.{i}.{w}.4.bias"] = (chans,)
    lay[prefix + "linear.weight"] = (1, chans)
    lay[prefix + "linear.bias"] = (1,)


def toucantts_layout(idim=62, odim=80, adim=192, heads=4, enc_layers=6, enc_units=1536, enc_kernel=7, dec_layers=6,
                     dec_units=1536, dec_kernel=31, dur_layers=3, dur_chans=256, dur_kernel=3, pitch_layers=7,
                     pitch_chans=256, pitch_kernel=5, energy_layers=2, energy_chans=256, energy_kernel=3,
                     utt_embed_dim=64, lang_embs=8000, flow_hidden=192, flow_kernel=5, flow_blocks=18, flow_layers=4,
                     flow_share_wn=4, postnet_layers=5, postnet_chans=256, postnet_filts=5):
    lay, alias = {}, {}
    lay["encoder.embed.0.weight"] = (100, idim)
    lay["encoder.embed.0.bias"] = (100,)
    lay["encoder.embed.2.weight"] = (adim, 100)
    lay["encoder.embed.2.bias"] = (adim,)
    lay["encoder.output_norm.weight"] = (adim,)
    lay["encoder.output_norm.bias"] = (adim,)
    if utt_embed_dim is not None:
        lay["encoder.hs_emb_projection.weight"] = (adim, adim + utt_embed_dim)
        lay["encoder.hs_emb_projection.bias"] = (adim,)
    if lang_embs is not None:
        lay["encoder.language_embedding.weight"] = (lang_embs, adim)
    _conformer_layout(lay, "encoder.", adim, heads, enc_units, enc_layers, enc_kernel)
    _predictor_layout(lay, "duration_predictor.", adim, dur_layers, dur_chans, dur_kernel, utt_embed_dim)
    _predictor_layout(lay, "pitch_predictor.", adim, pitch_layers, pitch_chans, pitch_kernel, utt_embed_dim)
    _predictor_layout(lay, "energy_predictor.", adim, energy_layers, energy_chans, energy_kernel, utt_embed_dim)
    lay["pitch_embed.0.weight"] = (adim, 1, 1)
    lay["pitch_embed.0.bias"] = (adim,)
    lay["energy_embed.0.weight"] = (adim, 1, 1)
    lay["energy_embed.0.bias"] = (adim,)
    _conformer_layout(lay, "decoder.", adim, heads, dec_units, dec_layers, dec_kernel)
    lay["feat_out.weight"] = (odim, adim)
    lay["feat_out.bias"] = (odim,)
    for i in range(postnet_layers):
        ic = odim if i == 0 else postnet_chans
        oc = odim if i == postnet_layers - 1 else postnet_chans
        lay[f"conv_postnet.postnet.{i}.0.weight"] = (oc, ic, postnet_filts)
        lay[f"conv_postnet.postnet.{i}.1.weight"] = (oc,)
        lay[f"conv_postnet.postnet.{i}.1.bias"] = (oc,)
    lay["post_flow.g_proj.weight"] = (adim, odim + adim, 5)
    lay["post_flow.g_proj.bias"] = (adim,)
    c2 = odim * 2
    for b in range(flow_blocks):
        an, ic, cp = f"post_flow.flows.{3 * b}.", f"post_flow.flows.{3 * b + 1}.", f"post_flow.flows.{3 * b + 2}."
        lay[an + "logs"] = (1, c2, 1)
        lay[an + "bias"] = (1, c2, 1)
        for name, shape in (("l", (4, 4)), ("log_s", (4,)), ("u", (4, 4)), ("p", (4, 4)), ("sign_s", (4,)),
                            ("l_mask", (4, 4)), ("eye", (4, 4))):
            lay[ic + name] = shape
        lay[cp + "start.bias"] = (flow_hidden,)
        lay[cp + "start.weight_g"] = (flow_hidden, 1, 1)
        lay[cp + "start.weight_v"] = (flow_hidden, c2 // 2, 1)
        lay[cp + "end.weight"] = (c2, flow_hidden, 1)
        lay[cp + "end.bias"] = (c2,)
        owner = f"post_flow.flows.{3 * (b - b % flow_share_wn) + 2}." if flow_share_wn > 0 else cp
        for n in range(flow_layers):
            rs = 2 * flow_hidden if n < flow_layers - 1 else flow_hidden
            for key, shape in ((f"wn.in_layers.{n}.bias", (2 * flow_hidden,)),
                               (f"wn.in_layers.{n}.weight_g", (2 * flow_hidden, 1, 1)),
                               (f"wn.in_layers.{n}.weight_v", (2 * flow_hidden, flow_hidden, flow_kernel))):
                lay[cp + key] = shape
                if owner != cp:
                    alias[cp + key] = owner + key
        for n in range(flow_layers):
            rs = 2 * flow_hidden if n < flow_layers - 1 else flow_hidden
            for key, shape in ((f"wn.res_skip_layers.{n}.bias", (rs,)),
                               (f"wn.res_skip_layers.{n}.weight_g", (rs, 1, 1)),
                               (f"wn.res_skip_layers.{n}.weight_v", (rs, flow_hidden, 1))):
                lay[cp + key] = shape
                if owner != cp:
                    alias[cp + key] = owner + key
        lay[cp + "wn.cond_layer.bias"] = (2 * flow_hidden * flow_layers,)
        lay[cp + "wn.cond_layer.weight_g"] = (2 * flow_hidden * flow_layers, 1, 1)
        lay[cp + "wn.cond_layer.weight_v"] = (2 * flow_hidden * flow_layers, 2 * adim, 1)
    return lay, alias


class _Node(torch.nn.Module):
    """Anonymous container; children and tensors are attached by dotted path."""


def attach(root, layout, alias=None):
    """Create nested containers under `root` so that root.state_dict() has exactly `layout`'s keys."""
    alias = alias or {}
    made = {}
    for key, shape in layout.items():
        *path, leaf = key.split(".")
        mod = root
        for part in path:
            if part not in mod._modules:
                mod.add_module(part, _Node())
            mod = mod._modules[part]
        if key in alias:
            mod.register_parameter(leaf, made[alias[key]])
            continue
        if leaf in BUFFER_LEAVES:
            dtype = torch.int64 if leaf == "num_batches_tracked" else torch.float32
            mod.register_buffer(leaf, torch.zeros(shape, dtype=dtype))
        else:
            prm = torch.nn.Parameter(torch.zeros(shape, dtype=torch.float32), requires_grad=False)
            mod.register_parameter(leaf, prm)
            made[key] = prm
    return root


def fold_weight_norm(sd):
    """g * v / ||v|| over all dims but 0 (what torch's remove_weight_norm leaves as `.weight`)."""
    out = {}
    for key, value in sd.items():
        if key.endswith(".weight_g"):
            v = sd[key[:-2] + "_v"]
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(value.shape)
            out[key[:-2]] = v * (value / norm)
        elif not key.endswith(".weight_v"):
            out[key] = value
    return out
