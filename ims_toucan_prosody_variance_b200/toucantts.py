"""Drop-in ToucanTTS acoustic model (text -> mel) on the B200 engine.

Same constructor keywords, `forward` / `_forward` / `store_inverse_all` signatures and the same
state_dict layout as InferenceInterfaces/InferenceArchitectures/InferenceToucanTTS.py:16-343 of the
reference.  `forward(text, ...)` keeps the reference's batch-1 contract; `_forward` accepts real
batches (the reference's `_forward` has batch-shaped parameters but silently breaks for B > 1, see
SURVEY.md 8b), and `synthesize_batch` is the additive ragged-batch entry that also returns the mel
in the NCL layout the vocoders consume.  A batched call equals per-utterance batch-1 reference calls:
every convolution zero-pads at the utterance's own ends, attention / norms / GroupNorm statistics see
only the utterance's own positions.

All activations live in HBM as fp32 NCL tensors (B, C, L) with L contiguous.  The dense contractions
(Linear, Conv1d k=1/3/5) are tb200_conv1d launches (tcgen05, operand precision `precision`), the rest
are the kernels of csrc/acoustic.cu and csrc/ragged.cu.  PyTorch only allocates memory here.
"""
import math

import torch

from . import layouts, ops
from ._lib import ACT_NONE, OUT_NONE, OUT_RELU, OUT_TANH, EngineError


def _pad4(n):
    return (int(n) + 3) // 4 * 4


class _Block:
    """Packed parameters of one Conformer block (EncoderLayer.py:62-144)."""


class ToucanTTS(torch.nn.Module):
    """InferenceToucanTTS.py:16-180 (constructor keywords kept; dropout rates are inert at inference)."""

    def __init__(self,
                 input_feature_dimensions=62, output_spectrogram_channels=80, attention_dimension=192, attention_heads=4,
                 positionwise_conv_kernel_size=1, use_scaled_positional_encoding=True, use_macaron_style_in_conformer=True,
                 use_cnn_in_conformer=True,
                 encoder_layers=6, encoder_units=1536, encoder_normalize_before=True, encoder_concat_after=False,
                 conformer_encoder_kernel_size=7, transformer_enc_dropout_rate=0.2,
                 transformer_enc_positional_dropout_rate=0.2, transformer_enc_attn_dropout_rate=0.2,
                 decoder_layers=6, decoder_units=1536, decoder_concat_after=False, conformer_decoder_kernel_size=31,
                 decoder_normalize_before=True, transformer_dec_dropout_rate=0.2,
                 transformer_dec_positional_dropout_rate=0.2, transformer_dec_attn_dropout_rate=0.2,
                 duration_predictor_layers=3, duration_predictor_chans=256, duration_predictor_kernel_size=3,
                 duration_predictor_dropout_rate=0.2,
                 pitch_predictor_layers=7, pitch_predictor_chans=256, pitch_predictor_kernel_size=5,
                 pitch_predictor_dropout=0.5, pitch_embed_kernel_size=1, pitch_embed_dropout=0.0,
                 energy_predictor_layers=2, energy_predictor_chans=256, energy_predictor_kernel_size=3,
                 energy_predictor_dropout=0.5, energy_embed_kernel_size=1, energy_embed_dropout=0.0,
                 utt_embed_dim=64, detach_postflow=True, lang_embs=8000, weights=None,
                 precision="tf32", duration_precision="fp32"):
        super().__init__()
        if not (positionwise_conv_kernel_size == 1 and use_macaron_style_in_conformer and use_cnn_in_conformer
                and encoder_normalize_before and decoder_normalize_before and not encoder_concat_after
                and not decoder_concat_after and pitch_embed_kernel_size == 1 and energy_embed_kernel_size == 1):
            raise EngineError("unsupported ToucanTTS variant (the engine covers the reference's inference configuration)")
        if (duration_predictor_chans != pitch_predictor_chans) or (pitch_predictor_chans != energy_predictor_chans):
            raise EngineError("the engine expects one channel width for the three variance predictors")
        self.input_feature_dimensions = input_feature_dimensions
        self.output_spectrogram_channels = output_spectrogram_channels
        self.attention_dimension = attention_dimension
        self.attention_heads = attention_heads
        self.detach_postflow = detach_postflow
        self.use_scaled_pos_enc = use_scaled_positional_encoding
        self.multilingual_model = lang_embs is not None
        self.multispeaker_model = utt_embed_dim is not None
        self.utt_embed_dim = utt_embed_dim
        self.encoder_layers, self.decoder_layers = encoder_layers, decoder_layers
        self.encoder_units, self.decoder_units = encoder_units, decoder_units
        self.predictor_layers = dict(duration=duration_predictor_layers, pitch=pitch_predictor_layers,
                                     energy=energy_predictor_layers)
        self.predictor_chans = pitch_predictor_chans
        self.flow_blocks, self.flow_layers, self.flow_hidden = 18, 4, 192
        self.precision, self.duration_precision = precision, duration_precision
        lay, alias = layouts.toucantts_layout(
            idim=input_feature_dimensions, odim=output_spectrogram_channels, adim=attention_dimension,
            heads=attention_heads, enc_layers=encoder_layers, enc_units=encoder_units,
            enc_kernel=conformer_encoder_kernel_size, dec_layers=decoder_layers, dec_units=decoder_units,
            dec_kernel=conformer_decoder_kernel_size, dur_layers=duration_predictor_layers,
            dur_chans=duration_predictor_chans, dur_kernel=duration_predictor_kernel_size,
            pitch_layers=pitch_predictor_layers, pitch_chans=pitch_predictor_chans,
            pitch_kernel=pitch_predictor_kernel_size, energy_layers=energy_predictor_layers,
            energy_chans=energy_predictor_chans, energy_kernel=energy_predictor_kernel_size, utt_embed_dim=utt_embed_dim,
            lang_embs=lang_embs)
        layouts.attach(self, lay, alias)
        self._packed = None
        self._pos_cache = {}
        self._graphs = None          # enable_cuda_graphs()
        if weights is not None:
            self.load_state_dict(weights)
        self.eval()

    # ------------------------------------------------------------------------------------------
    # load-time packing (A16)
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def store_inverse_all(self):
        """Reference: caches the InvConvNear inverses and folds weight norm (InferenceToucanTTS.py:321-330,
        Glow.py:137-139,271-272).  Here additionally: packs every Linear/Conv weight into its tensor-core
        operand image.  The parameters themselves (state_dict) are left untouched."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise EngineError("toucan_b200 ToucanTTS runs on CUDA only: call .to('cuda') before store_inverse_all()/forward()")
        if self._graphs is not None:
            self._graphs = {}        # captured graphs point into the previous packing
        # folded on the host (load time; a few hundred small tensors), so that the only device work of loading a model is
        # the packing kernels of this library and plain copies
        sd = {k: v.to(dev) for k, v in layouts.fold_weight_norm({k: v.detach().cpu() for k, v in self.state_dict().items()}).items()}
        prec = self.precision
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        cache = {}

        def conv(wkey, bkey=None, dilation=1, precision=None):
            w = sd[wkey]
            key = (w.data_ptr(), precision or prec)
            if key not in cache:
                pad = (w.shape[-1] - 1) // 2 * dilation if w.dim() == 3 else 0
                cache[key] = ops.ConvLayer(w, sd[bkey] if bkey else None, dilation=dilation, padding=pad,
                                           precision=precision or prec)
            return cache[key]

        pk = {}
        pk["embed0"] = conv("encoder.embed.0.weight", "encoder.embed.0.bias")
        pk["embed2"] = conv("encoder.embed.2.weight", "encoder.embed.2.bias")
        pk["out_norm"] = (f32(sd["encoder.output_norm.weight"]), f32(sd["encoder.output_norm.bias"]))
        if self.multispeaker_model:
            pk["hs_proj"] = conv("encoder.hs_emb_projection.weight", "encoder.hs_emb_projection.bias")
        if self.multilingual_model:
            pk["lang_emb"] = f32(sd["encoder.language_embedding.weight"])

        def block(prefix):
            blk = _Block()
            p = prefix
            for ff in ("feed_forward_macaron", "feed_forward"):
                setattr(blk, ff + "_w1", conv(f"{p}{ff}.w_1.weight", f"{p}{ff}.w_1.bias"))
                setattr(blk, ff + "_w2", conv(f"{p}{ff}.w_2.weight", f"{p}{ff}.w_2.bias"))
            wqkv = torch.cat([sd[f"{p}self_attn.linear_{n}.weight"] for n in "qkv"], dim=0)
            bqkv = torch.cat([sd[f"{p}self_attn.linear_{n}.bias"] for n in "qkv"], dim=0)
            blk.qkv = ops.ConvLayer(wqkv, bqkv, precision=prec)
            blk.pos = ops.ConvLayer(sd[f"{p}self_attn.linear_pos.weight"], None, precision=prec)
            blk.out = conv(f"{p}self_attn.linear_out.weight", f"{p}self_attn.linear_out.bias")
            blk.bias_u, blk.bias_v = f32(sd[f"{p}self_attn.pos_bias_u"]), f32(sd[f"{p}self_attn.pos_bias_v"])
            blk.pw1 = conv(f"{p}conv_module.pointwise_conv1.weight", f"{p}conv_module.pointwise_conv1.bias")
            blk.pw2 = conv(f"{p}conv_module.pointwise_conv2.weight", f"{p}conv_module.pointwise_conv2.bias")
            blk.dw_w = f32(sd[f"{p}conv_module.depthwise_conv.weight"]).reshape(self.attention_dimension, -1).contiguous()
            blk.dw_b = f32(sd[f"{p}conv_module.depthwise_conv.bias"])
            blk.bn = tuple(f32(sd[f"{p}conv_module.norm.{n}"]) for n in ("running_mean", "running_var", "weight", "bias"))
            blk.norms = {n: (f32(sd[f"{p}{n}.weight"]), f32(sd[f"{p}{n}.bias"]))
                         for n in ("norm_ff_macaron", "norm_mha", "norm_conv", "norm_ff", "norm_final")}
            blk.pos_table, blk.pos16, blk.pos_tables = None, None, {}
            return blk

        pk["enc"] = [block(f"encoder.encoders.{i}.") for i in range(self.encoder_layers)]
        pk["dec"] = [block(f"decoder.encoders.{i}.") for i in range(self.decoder_layers)]

        # variance predictors: convs + the stacked ConditionalLayerNorm MLPs (one launch for all of them)
        mlp_keys = []
        for name in ("duration", "pitch", "energy"):
            pprec = self.duration_precision if name == "duration" else prec
            layers = []
            for i in range(self.predictor_layers[name]):
                layers.append(conv(f"{name}_predictor.conv.{i}.0.weight", f"{name}_predictor.conv.{i}.0.bias", precision=pprec))
                if self.multispeaker_model:
                    mlp_keys += [f"{name}_predictor.norms.{i}.W_scale.", f"{name}_predictor.norms.{i}.W_bias."]
                else:
                    pk[f"{name}.ln{i}"] = (f32(sd[f"{name}_predictor.norms.{i}.weight"]), f32(sd[f"{name}_predictor.norms.{i}.bias"]))
            pk[f"{name}.convs"] = layers
            # the final projection to one value per phoneme stays fp32 (durations must be bit-exact)
            pk[f"{name}.linear"] = conv(f"{name}_predictor.linear.weight", f"{name}_predictor.linear.bias", precision="fp32")
        if self.multispeaker_model:
            pk["cln"] = tuple(torch.stack([f32(sd[k + leaf]) for k in mlp_keys]).contiguous()
                              for leaf in ("0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias"))
        adim = self.attention_dimension
        pk["pitch_embed"] = (f32(sd["pitch_embed.0.weight"]).reshape(adim), f32(sd["pitch_embed.0.bias"]))
        pk["energy_embed"] = (f32(sd["energy_embed.0.weight"]).reshape(adim), f32(sd["energy_embed.0.bias"]))

        pk["feat_out"] = conv("feat_out.weight", "feat_out.bias")
        pk["postnet"] = [(conv(f"conv_postnet.postnet.{i}.0.weight"), f32(sd[f"conv_postnet.postnet.{i}.1.weight"]),
                          f32(sd[f"conv_postnet.postnet.{i}.1.bias"])) for i in range(5)]

        pk["g_proj"] = conv("post_flow.g_proj.weight", "post_flow.g_proj.bias")
        flows = []
        hid = self.flow_hidden
        for b in range(self.flow_blocks):
            an, ic, cp = f"post_flow.flows.{3 * b}.", f"post_flow.flows.{3 * b + 1}.", f"post_flow.flows.{3 * b + 2}."
            fl = _Block()
            fl.an_bias, fl.an_logs = f32(sd[an + "bias"]).reshape(-1), f32(sd[an + "logs"]).reshape(-1)
            # Glow.py:130-139: W = P (L*mask + I)(U*mask^T + diag(sign_s*exp(log_s))), cached fp32 inverse
            ich = {k: sd[ic + k].float().cpu() for k in ("l", "l_mask", "eye", "u", "sign_s", "log_s", "p")}   # 4x4: on the host
            l = ich["l"] * ich["l_mask"] + ich["eye"]
            u = ich["u"] * ich["l_mask"].transpose(0, 1).contiguous() + torch.diag(ich["sign_s"] * torch.exp(ich["log_s"]))
            fl.w_inv = torch.inverse(torch.matmul(ich["p"], torch.matmul(l, u))).to(dev).contiguous()
            fl.start = conv(cp + "start.weight", cp + "start.bias")
            fl.end = conv(cp + "end.weight", cp + "end.bias")
            fl.cond = conv(cp + "wn.cond_layer.weight", cp + "wn.cond_layer.bias")
            fl.in_layers, fl.res, fl.skip = [], [], []
            for n in range(self.flow_layers):
                fl.in_layers.append(conv(f"{cp}wn.in_layers.{n}.weight", f"{cp}wn.in_layers.{n}.bias"))
                w, bias = sd[f"{cp}wn.res_skip_layers.{n}.weight"], sd[f"{cp}wn.res_skip_layers.{n}.bias"]
                key = (w.data_ptr(), "split")
                if key not in cache:
                    if n < self.flow_layers - 1:
                        cache[key] = (ops.ConvLayer(w[:hid].contiguous(), bias[:hid].contiguous(), precision=prec),
                                      ops.ConvLayer(w[hid:].contiguous(), bias[hid:].contiguous(), precision=prec))
                    else:
                        cache[key] = (None, ops.ConvLayer(w, bias, precision=prec))
                fl.res.append(cache[key][0])
                fl.skip.append(cache[key][1])
            flows.append(fl)
        pk["flows"] = flows
        self._packed = pk
        self._pos_cache = {}

    # ------------------------------------------------------------------------------------------
    # relative positional table (PositionalEncoding.py:95-130) and its per-layer projection
    # ------------------------------------------------------------------------------------------
    def _positions(self, blk, l_max, dev):
        """(D, 2*cap-1) tensor whose column (cap-1-r) holds linear_pos(PE(r)); computed once per layer and table size
        and kept for the model's lifetime (captured CUDA graphs hold pointers into these tables)."""
        cap = 256
        while cap < l_max:
            cap *= 2
        hit = blk.pos_tables.get((cap, str(dev)))
        if hit is not None:
            blk.pos_table, blk.pos16 = hit
            return blk.pos_table
        key = (cap, str(dev))
        pe = self._pos_cache.get(key)
        if pe is None:
            d = self.attention_dimension
            rel = torch.arange(cap - 1, -cap, -1, dtype=torch.float32).unsqueeze(1)  # row k <-> relative position cap-1-k
            div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
            tab = torch.zeros(2 * cap - 1, d)
            tab[:, 0::2] = torch.sin(rel * div)
            tab[:, 1::2] = torch.cos(rel * div)
            pe = torch.zeros((1, d, _pad4(2 * cap - 1)), dtype=torch.float32, device=dev)
            pe[0, :, :2 * cap - 1] = tab.t().to(dev)
            self._pos_cache = {key: pe}
        out = torch.zeros((1, self.attention_dimension, pe.shape[2]), dtype=torch.float32, device=dev)
        blk.pos(pe, None, out, l_in_max=2 * cap - 1)
        blk.pos_table = (out[0], cap)
        blk.pos16 = None
        if self.precision != "fp32":   # operand rows of the tensor-core attention, packed once per layer and table size
            blk.pos16 = ops.pack_relpos_table(out[0][:, :2 * cap - 1], cap - 1, self.attention_heads)
        blk.pos_tables[(cap, str(dev))] = (blk.pos_table, blk.pos16)
        return blk.pos_table

    # ------------------------------------------------------------------------------------------
    # building blocks
    # ------------------------------------------------------------------------------------------
    def _conformer_block(self, blk, x, lens, l_max, ws, dw_kernel):
        """EncoderLayer.py:62-144 in place on x (B,D,L)."""
        n, h, qkv, ctx = ws["n"], ws["h"], ws["qkv"], ws["ctx"]
        units = blk.feed_forward_w1.c_out
        hv = h[:, :units]
        # x += 0.5 * FFN(LN(x))   (macaron)
        ops.channel_norm(x, lens, n, *blk.norms["norm_ff_macaron"], l_max)
        blk.feed_forward_macaron_w1(n, lens, hv, l_in_max=l_max, out_act=OUT_RELU)
        blk.feed_forward_macaron_w2(hv, lens, x, l_in_max=l_max, out_alpha=0.5, residual=x)
        # x += MHA(LN(x))
        ops.channel_norm(x, lens, n, *blk.norms["norm_mha"], l_max)
        blk.qkv(n, lens, qkv, l_in_max=l_max)
        pos, cap = self._positions(blk, l_max, x.device)
        if self.precision != "fp32":
            ops.relpos_attention_tc(qkv, lens, ctx, blk.pos16[0], blk.pos16[1], blk.bias_u, blk.bias_v, self.attention_heads, l_max)
        else:
            ops.relpos_attention(qkv, lens, ctx, pos, cap - 1, blk.bias_u, blk.bias_v, self.attention_heads, l_max)
        blk.out(ctx, lens, x, l_in_max=l_max, residual=x)
        # x += ConvModule(LN(x))
        ops.channel_norm(x, lens, n, *blk.norms["norm_conv"], l_max)
        g = h[:, :2 * self.attention_dimension]
        blk.pw1(n, lens, g, l_in_max=l_max)
        ops.glu_dwconv(g, lens, ctx, blk.dw_w, blk.dw_b, *blk.bn, l_max)
        blk.pw2(ctx, lens, x, l_in_max=l_max, residual=x)
        # x += 0.5 * FFN(LN(x));  x = LN(x)
        ops.channel_norm(x, lens, n, *blk.norms["norm_ff"], l_max)
        blk.feed_forward_w1(n, lens, hv, l_in_max=l_max, out_act=OUT_RELU)
        blk.feed_forward_w2(hv, lens, x, l_in_max=l_max, out_alpha=0.5, residual=x)
        ops.channel_norm(x, lens, x, *blk.norms["norm_final"], l_max)

    def _conformer_ws(self, b, l_max, units, dev):
        d = self.attention_dimension
        ld = _pad4(l_max)
        z = lambda c: torch.zeros((b, c, ld), dtype=torch.float32, device=dev)  # noqa: E731
        return dict(n=z(d), h=z(max(units, 2 * d)), qkv=z(3 * d), ctx=z(d))

    def _predictor(self, name, enc, lens, t_max, cln, cln_base):
        """VariancePredictor.py:53-80 / DurationPredictor.py:63-77: returns (B, T_ld) pre-activation values."""
        pk = self._packed
        b, _, ld = enc.shape
        ch = self.predictor_chans
        a = torch.zeros((b, ch, ld), dtype=torch.float32, device=enc.device)
        hbuf = torch.zeros((b, ch, ld), dtype=torch.float32, device=enc.device)
        h = enc
        for i, layer in enumerate(pk[f"{name}.convs"]):
            layer(h, lens, a, l_in_max=t_max, out_act=OUT_RELU)
            if cln is not None:
                ops.channel_norm(a, lens, hbuf, cln[cln_base + 2 * i], cln[cln_base + 2 * i + 1], t_max, conditional=True)
            else:
                ops.channel_norm(a, lens, hbuf, *pk[f"{name}.ln{i}"], t_max)
            h = hbuf
        out = torch.zeros((b, 1, ld), dtype=torch.float32, device=enc.device)
        pk[f"{name}.linear"](h, lens, out, l_in_max=t_max)
        return out[:, 0, :]

    # ------------------------------------------------------------------------------------------
    # the batched hot path
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def synthesize_batch(self, text_tensors, text_lengths, gold_durations=None, gold_pitch=None, gold_energy=None,
                         duration_scaling_factor=1.0, utterance_embedding=None, lang_ids=None, pitch_variance_scale=1.0,
                         energy_variance_scale=1.0, pause_duration_scaling_factor=1.0, noise=None, taps=None):
        """text_tensors (B,T,62) fp32 CUDA, text_lengths (B).  gold_* are (B,T) / (B,T,1) padded tensors.
        noise: None -> drawn per utterance like Glow.py:363 (torch.randn((1,80,F)) from the global CPU generator,
        in utterance order); "device" -> drawn on the GPU; or a (B,80,>=Fmax) tensor of standard normals.
        Returns a dict: mel_ncl (B,80,F'max_ld), mel_lengths (B) int32 [= 2*floor(F/2)], frames (B) int32,
        decoded_ncl (B,80,Fld), durations (B,T) int64, pitch (B,T), energy (B,T), log_durations (B,T) or None.
        taps: optional dict that receives clones of the stage outputs (NCL) for per-stage parity checks.

        The path is two device-only segments around its one host sync (the frame counts that size the decoder
        buffers): `_segment_text` (encoder, predictors, prosody edits) and `_segment_frames` (length regulator,
        decoder, PostNet, PostFlow).  With `enable_cuda_graphs()` each segment is captured once per padded shape and
        replayed (see `_graphed`)."""
        if self._packed is None:
            self.store_inverse_all()
        dev = text_tensors.device
        if dev.type != "cuda":
            raise EngineError("toucan_b200 has no CPU path: tensors must live on a CUDA device")
        b, t_max, _ = text_tensors.shape
        odim = self.output_spectrogram_channels
        text = text_tensors.contiguous().float()
        tlen = text_lengths.to(device=dev, dtype=torch.int32).contiguous()
        if self.multispeaker_model and utterance_embedding is None:
            raise EngineError("multispeaker model needs an utterance embedding")
        emb = utterance_embedding.to(dev).reshape(b, -1).float().contiguous() if self.multispeaker_model else None
        lang = lang_ids.to(dev).reshape(-1).long() if (self.multilingual_model and lang_ids is not None) else None
        scal = (float(duration_scaling_factor), float(pitch_variance_scale), float(energy_variance_scale),
                float(pause_duration_scaling_factor))
        graphs = (self._graphs is not None and taps is None and gold_durations is None and gold_pitch is None
                  and gold_energy is None)

        if graphs:
            r1, key1 = self._graphed_text(text, tlen, emb, lang, scal)
        else:
            gold = tuple(None if g is None else g.to(dev) for g in (gold_durations, gold_pitch, gold_energy))
            r1 = self._segment_text(text, tlen, emb, lang, gold, scal, t_max, taps)

        # ---- the one host sync of the path: frame counts size every later buffer
        frames_host = r1["frames"].cpu()
        f_max = int(frames_host.max())
        if f_max <= 0:
            raise EngineError("synthesize_batch: no frames to synthesise")
        if f_max // 2 <= 0:
            raise EngineError("synthesize_batch: utterances shorter than 2 frames cannot pass the PostFlow")
        f_run = (f_max + 63) // 64 * 64 if graphs else f_max      # graphs: one capture per 64-frame bucket
        f_ld = _pad4(f_run)
        if noise is None:
            zn = torch.zeros((b, odim, f_ld), dtype=torch.float32)
            for i in range(b):
                fi = int(frames_host[i])
                if fi > 0:
                    zn[i, :, :fi] = torch.randn((1, odim, fi))[0]
            zn = zn.to(dev)
        elif isinstance(noise, str) and noise == "device":
            zn = torch.randn((b, odim, f_ld), dtype=torch.float32, device=dev)
        else:
            zn = torch.zeros((b, odim, f_ld), dtype=torch.float32, device=dev)
            zn[:, :, :f_max] = noise.to(dev)[:, :, :f_max]

        if graphs:
            r2 = self._graphed_frames(key1, r1, tlen, zn, f_run)
        else:
            r2 = self._segment_frames(r1["enc"], r1["cum"], tlen, r1["frames"], r1["pitch"], r1["energy"], zn, f_run, taps)
        return dict(mel_ncl=r2["mel"], mel_lengths=r2["mel_lengths"], frames=r1["frames"], decoded_ncl=r2["decoded"],
                    durations=r1["dur"], pitch=r1["pitch"], energy=r1["energy"], log_durations=r1["log_d"],
                    frames_host=frames_host)

    def _segment_text(self, text, tlen, emb, lang, gold, scal, t_max, taps=None):
        """Encoder, variance predictors, prosody edits, duration rounding + prefix sums (A2-A8).  Device only.
        text (B,T,62) with T >= t_max; every kernel masks by tlen, so t_max may be padded up."""
        pk = self._packed
        dev = text.device
        b, _, idim = text.shape
        d = self.attention_dimension
        duration_scale, pitch_scale, energy_scale, pause_scale = scal
        gold_durations, gold_pitch, gold_energy = gold
        t_ld = _pad4(t_max)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731

        # ---- utterance embedding: normalised once here (InferenceToucanTTS.py:202), once more in the encoder (Conformer.py:132)
        e1 = e2 = None
        if self.multispeaker_model:
            e1 = ops.l2_normalize(emb)
            e2 = ops.l2_normalize(e1)

        # ---- encoder (Conformer.py:92-134)
        x_in = ops.to_ncl(text, tlen, z(b, idim, t_ld), t_max)
        h100 = pk["embed0"](x_in, tlen, z(b, pk["embed0"].c_out, t_ld), l_in_max=t_max, out_act=OUT_TANH)
        x = pk["embed2"](h100, tlen, z(b, d, t_ld), l_in_max=t_max)
        lang_vec = pk["lang_emb"].index_select(0, lang).contiguous() if lang is not None else None
        ops.rowvec_affine(x, tlen, x, t_max, vec=lang_vec, scale=math.sqrt(d))
        ws = self._conformer_ws(b, t_max, self.encoder_units, dev)
        for blk in pk["enc"]:
            self._conformer_block(blk, x, tlen, t_max, ws, None)
        if self.multispeaker_model:
            cat = z(b, d + self.utt_embed_dim, t_ld)
            ops.channel_norm(x, tlen, cat[:, :d], *pk["out_norm"], t_max)
            ops.rowvec_affine(None, tlen, cat[:, d:], t_max, vec=e2)
            enc = pk["hs_proj"](cat, tlen, z(b, d, t_ld), l_in_max=t_max)
        else:
            enc = ops.channel_norm(x, tlen, z(b, d, t_ld), *pk["out_norm"], t_max)

        # ---- variance predictors (A7) and prosody edits (A8)
        cln = ops.cln_mlp(e1, *pk["cln"]) if self.multispeaker_model else None
        base = {"duration": 0, "pitch": 2 * self.predictor_layers["duration"],
                "energy": 2 * (self.predictor_layers["duration"] + self.predictor_layers["pitch"])}
        if gold_pitch is None:
            pitch = self._predictor("pitch", enc, tlen, t_max, cln, base["pitch"])
        else:
            pitch = z(b, t_ld)
            pitch[:, :t_max] = gold_pitch.reshape(b, t_max).float()
        if gold_energy is None:
            energy = self._predictor("energy", enc, tlen, t_max, cln, base["energy"])
        else:
            energy = z(b, t_ld)
            energy[:, :t_max] = gold_energy.reshape(b, t_max).float()
        log_d = None
        if gold_durations is None:
            log_d = self._predictor("duration", enc, tlen, t_max, cln, base["duration"])
            dur, cum, frames = ops.duration_finalize(text, tlen, log_dur=log_d, pause_scale=pause_scale,
                                                     duration_scale=duration_scale)
        else:
            gd = torch.zeros((b, t_ld), dtype=torch.int64, device=dev)
            gd[:, :t_max] = gold_durations.reshape(b, t_max).long()
            dur, cum, frames = ops.duration_finalize(text, tlen, gold_dur=gd, pause_scale=pause_scale,
                                                     duration_scale=duration_scale)
        if taps is not None:
            taps.update(encoder=enc.clone(), pitch_raw=pitch.clone(), energy_raw=energy.clone())
        pitch = pitch.contiguous()
        energy = energy.contiguous()
        ops.variance_edit(pitch, text, tlen, 0, pitch_scale)
        ops.variance_edit(energy, text, tlen, 1, energy_scale)
        return dict(enc=enc, pitch=pitch, energy=energy, dur=dur, cum=cum, frames=frames, log_d=log_d)

    def _segment_frames(self, enc, cum, tlen, frames, pitch, energy, zn, f_max, taps=None):
        """Length regulator, decoder, feat_out + PostNet, PostFlow (A9-A13).  Device only; f_max >= max(frames) (every
        kernel masks by `frames`, so it may be padded up); zn (B,80,>=f_max) standard normals (scaled by 0.8 here, in
        place)."""
        pk = self._packed
        dev = enc.device
        b = enc.shape[0]
        d = self.attention_dimension
        f_ld = _pad4(f_max)
        odim = self.output_spectrogram_channels
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731

        # ---- length regulator + pitch/energy embedding (A9, A10) straight into the PostFlow conditioning buffer
        cat = z(b, odim + d, f_ld)             # rows [0,80): refined mel, rows [80,272): upsampled enriched encoding
        up = cat[:, odim:]
        ops.length_regulate(enc, cum, tlen, frames, f_max, pitch=pitch, energy=energy, wp=pk["pitch_embed"][0],
                            bp=pk["pitch_embed"][1], we=pk["energy_embed"][0], be=pk["energy_embed"][1], out=up)

        # ---- decoder (A11)
        x = ops.rowvec_affine(up, frames, z(b, d, f_ld), f_max, scale=math.sqrt(d))
        ws = self._conformer_ws(b, f_max, self.decoder_units, dev)
        for blk in pk["dec"]:
            self._conformer_block(blk, x, frames, f_max, ws, None)

        if taps is not None:
            taps.update(upsampled=up.clone(), decoder=x.clone())

        # ---- feat_out + PostNet (A12)
        decoded = pk["feat_out"](x, frames, z(b, odim, f_ld), l_in_max=f_max)
        pch = pk["postnet"][0][0].c_out
        pa, pb = z(b, pch, f_ld), z(b, pch, f_ld)
        h = decoded
        for i, (layer, gamma, beta) in enumerate(pk["postnet"]):
            last = i == len(pk["postnet"]) - 1
            raw = layer(h, frames, pa[:, :layer.c_out], l_in_max=f_max)
            if last:
                ops.group_norm(raw, frames, cat[:, :odim], gamma, beta, 20, f_max, residual=decoded)
            else:
                h = ops.group_norm(raw, frames, pb, gamma, beta, 32, f_max, tanh=True)

        if taps is not None:
            taps.update(decoded=decoded.clone(), refined=cat[:, :odim].clone())

        # ---- PostFlow (A13): Glow.forward(infer=True), blocks in reverse
        hid = self.flow_hidden
        g = pk["g_proj"](cat, frames, z(b, hid, f_ld), l_in_max=f_max)
        len2 = torch.div(frames, 2, rounding_mode="floor").to(torch.int32)
        l2_max = f_max // 2
        l2_ld = _pad4(l2_max)
        ops.rowvec_affine(zn, frames, zn, f_max, scale=0.8)
        xf = ops.squeeze2(zn, frames, z(b, 2 * odim, l2_ld), f_max)
        g2 = ops.squeeze2(g, frames, z(b, 2 * hid, l2_ld), f_max)
        hb, cond = z(b, hid, l2_ld), z(b, 2 * hid * self.flow_layers, l2_ld)
        ab, acts, skip, ml = z(b, 2 * hid, l2_ld), z(b, hid, l2_ld), z(b, hid, l2_ld), z(b, 2 * odim, l2_ld)
        for fl in reversed(pk["flows"]):
            fl.start(xf[:, :odim], len2, hb, l_in_max=l2_max)
            fl.cond(g2, len2, cond, l_in_max=l2_max)
            for n in range(self.flow_layers):
                fl.in_layers[n](hb, len2, ab, l_in_max=l2_max, residual=cond[:, n * 2 * hid:(n + 1) * 2 * hid])
                ops.wn_gate(ab, len2, acts, l2_max)
                if fl.res[n] is not None:
                    fl.res[n](acts, len2, hb, l_in_max=l2_max, residual=hb)
                fl.skip[n](acts, len2, skip, l_in_max=l2_max, accumulate=n > 0)
            fl.end(skip, len2, ml, l_in_max=l2_max)
            ops.flow_close(xf, ml, len2, l2_max, fl.w_inv, fl.an_bias, fl.an_logs)
            if taps is not None:
                taps.setdefault("flow_blocks", []).append(xf.clone())
        mel = ops.squeeze2(xf, len2, z(b, odim, _pad4(2 * l2_max)), l2_max, inverse=True)
        return dict(mel=mel, mel_lengths=(len2 * 2).to(torch.int32), decoded=decoded)

    # ------------------------------------------------------------------------------------------
    # CUDA-graph replay of the two segments (one capture per padded shape)
    # ------------------------------------------------------------------------------------------
    def enable_cuda_graphs(self, max_cached=8):
        """Capture `_segment_text` per (batch, phonemes padded to 8, scaling factors) and `_segment_frames` per (that key,
        frames padded to 64) in CUDA graphs and replay them: the ~550 launches of a batch are enqueued by two graph
        launches instead of ~25 ms of host work, which is what bounds batches of fewer than ~100 utterances.  Inputs
        are copied into the graph's static buffers; results are views of graph-owned memory, valid until the next
        call with the same shape (TextToWave consumes them immediately).  Padding is free: every kernel masks by the
        utterance lengths.  `max_cached` bounds the captured text-segment shapes (LRU; their frame-segment graphs go
        with them).  Eager execution stays the default and is what gold prosody inputs and `taps` use."""
        self._graphs = {}
        self._graphs_max = int(max_cached)

    def disable_cuda_graphs(self):
        self._graphs = None

    @staticmethod
    def _capture(fn):
        """Warm `fn` up on a side stream (planner caches, position tables, allocator pool), then capture it."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = fn()
        return graph, out

    def _graphed_text(self, text, tlen, emb, lang, scal):
        b, t_max, idim = text.shape
        t_run = (t_max + 7) // 8 * 8
        key = (b, t_run, emb is not None, lang is not None, scal, str(text.device))
        ent = self._graphs.pop(key, None)
        if ent is None:
            while len(self._graphs) >= self._graphs_max:
                self._graphs.pop(next(iter(self._graphs)))          # least recently used first
            st = dict(text=torch.zeros((b, t_run, idim), dtype=torch.float32, device=text.device),
                      tlen=torch.zeros_like(tlen), emb=None if emb is None else torch.zeros_like(emb),
                      lang=None if lang is None else torch.zeros_like(lang))
            ent = dict(static=st, frames={})
        self._graphs[key] = ent                                       # most recently used last
        st = ent["static"]
        st["text"][:, :t_max].copy_(text)
        st["tlen"].copy_(tlen)
        if emb is not None:
            st["emb"].copy_(emb)
        if lang is not None:
            st["lang"].copy_(lang)
        if "graph" not in ent:
            ent["graph"], ent["out"] = self._capture(
                lambda: self._segment_text(st["text"], st["tlen"], st["emb"], st["lang"], (None, None, None), scal, t_run))
        ent["graph"].replay()
        return ent["out"], key

    def _graphed_frames(self, key1, r1, tlen, zn, f_run):
        ent1 = self._graphs[key1]
        ent = ent1["frames"].get(f_run)
        st = ent1["static"]
        if ent is None:
            if len(ent1["frames"]) >= 4:
                ent1["frames"].pop(next(iter(ent1["frames"])))
            zs = torch.zeros_like(zn)
            ent = dict(zn=zs)
            zs.copy_(zn)
            ent["graph"], ent["out"] = self._capture(
                lambda: self._segment_frames(r1["enc"], r1["cum"], st["tlen"], r1["frames"], r1["pitch"], r1["energy"], zs, f_run))
            ent1["frames"][f_run] = ent
        ent["zn"].copy_(zn)
        ent["graph"].replay()
        return ent["out"]

    # ------------------------------------------------------------------------------------------
    # reference-shaped entry points
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _forward(self, text_tensors, text_lengths, gold_durations=None, gold_pitch=None, gold_energy=None,
                 duration_scaling_factor=1.0, utterance_embedding=None, lang_ids=None, pitch_variance_scale=1.0,
                 energy_variance_scale=1.0, pause_duration_scaling_factor=1.0, noise=None):
        """InferenceToucanTTS.py:183-250.  Returns (before_outs, after_outs, durations, pitch, energy), each
        `.squeeze()`d like the reference; for B > 1 the tensors are padded to the longest utterance."""
        if not self.multilingual_model:
            lang_ids = None
        if not self.multispeaker_model:
            utterance_embedding = None
        r = self.synthesize_batch(text_tensors, text_lengths, gold_durations, gold_pitch, gold_energy,
                                  duration_scaling_factor, utterance_embedding, lang_ids, pitch_variance_scale,
                                  energy_variance_scale, pause_duration_scaling_factor, noise=noise)
        b, t_max = text_tensors.shape[0], text_tensors.shape[1]
        dev = text_tensors.device
        odim = self.output_spectrogram_channels
        f_max = int(r["frames_host"].max())
        m_max = 2 * (f_max // 2)
        before = torch.zeros((b, f_max, odim), dtype=torch.float32, device=dev)
        after = torch.zeros((b, m_max, odim), dtype=torch.float32, device=dev)
        ops.from_ncl(r["decoded_ncl"], r["frames"], before, f_max)
        ops.from_ncl(r["mel_ncl"], r["mel_lengths"], after, m_max)
        return (before.squeeze(), after.squeeze(), r["durations"][:, :t_max].squeeze(), r["pitch"][:, :t_max].squeeze(),
                r["energy"][:, :t_max].squeeze())

    @torch.inference_mode()
    def forward(self, text, durations=None, pitch=None, energy=None, utterance_embedding=None,
                return_duration_pitch_energy=False, lang_id=None, duration_scaling_factor=1.0, pitch_variance_scale=1.0,
                energy_variance_scale=1.0, pause_duration_scaling_factor=1.0):
        """InferenceToucanTTS.py:252-319: one utterance, text (T,62) -> mel (F',80)."""
        text_length = torch.tensor([text.shape[0]], dtype=torch.long, device=text.device)
        if durations is not None:
            durations = durations.unsqueeze(0).to(text.device)
        if pitch is not None:
            pitch = pitch.unsqueeze(0).to(text.device)
        if energy is not None:
            energy = energy.unsqueeze(0).to(text.device)
        if lang_id is not None:
            lang_id = lang_id.unsqueeze(0).to(text.device)
        _, after, dur, p, e = self._forward(
            text.unsqueeze(0), text_length, gold_durations=durations, gold_pitch=pitch, gold_energy=energy,
            utterance_embedding=utterance_embedding.unsqueeze(0) if utterance_embedding is not None else None,
            lang_ids=lang_id, duration_scaling_factor=duration_scaling_factor, pitch_variance_scale=pitch_variance_scale,
            energy_variance_scale=energy_variance_scale, pause_duration_scaling_factor=pause_duration_scaling_factor)
        if return_duration_pitch_energy:
            return after, dur, p, e
        return after
