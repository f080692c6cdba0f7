"""Drop-in vocoder generators: HiFiGAN ("Avocodo") and BigVGAN on the B200 engine.

Same constructor and forward signatures and the same state_dict layout as
InferenceInterfaces/InferenceArchitectures/InferenceAvocodo.py:6-96 and InferenceBigVGAN.py:19-121 of
the reference; `forward(c)` keeps the reference's batch-1 contract ((80,T) -> (T*384,)), and
`forward_batch` is the additive batched entry ((B,80,T) + per-utterance lengths), whose result for
each utterance equals a batch-1 call (zero / replicate padding at the utterance's own ends).

Residual pairs (act -> conv_d -> act -> conv_1 -> +x: AMP.py:51-60, ResidualBlock.py:93-97) of the stages
with <= 128 channels are ONE tb200_respair launch each: the tensor between the two convolutions never
leaves the SM, and the 1/3 multi-receptive-field mean is folded into the last pair of a block.  Every
other convolution (input / output convs, up-samplers, the 256-channel stage) is one tb200_conv1d launch
with its activation fused in the prologue and bias / residual / mean in the epilogue.
fuse_pairs=False keeps the two-launch form of every pair (fp16 tensor between the convs through HBM).
"""
import torch

from . import layouts, ops
from ._lib import ACT_AA_SNAKEBETA, ACT_LEAKY_RELU, ACT_NONE, OUT_NONE, OUT_TANH


def _pad4(n):
    return (n + 3) // 4 * 4


def _pad8(n):
    return (n + 7) // 8 * 8


class _GeneratorBase(torch.nn.Module):
    """Shared engine of the two generators; subclasses provide the key naming and activations."""

    upsample_scales = (8, 6, 4, 2)

    def _init_engine(self, precision, activation_dtype="f32", fuse_pairs=True):
        if activation_dtype not in ("f32", "f16") or (activation_dtype == "f16" and precision != "f16"):
            raise ops._lib.EngineError("activation_dtype must be 'f32', or 'f16' together with precision='f16'")
        self.precision = precision
        self.activation_dtype = activation_dtype   # storage type of the residual stream in HBM
        self.fuse_pairs = bool(fuse_pairs)
        self._packed = None
        self._buffers_cache = {}

    # -- to be provided by subclasses --------------------------------------------------------
    def _names(self):
        raise NotImplementedError

    def _stage_channels(self, i):
        return self.channels // 2 ** (i + 1)

    # -- load-time packing -------------------------------------------------------------------
    @torch.no_grad()
    def remove_weight_norm(self):
        """Reference: folds weight_g * v/||v|| into .weight (InferenceAvocodo.py:82-89,
        InferenceBigVGAN.py:97-105).  Here: fold on the parameters' device and pack the tensor-core
        operand images.  The weight_g / weight_v parameters stay as they are (state_dict unchanged)."""
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise ops._lib.EngineError("toucan_b200 generators run on CUDA only: call .to('cuda') before "
                                       "remove_weight_norm()/forward()")
        # folded on the host (load time; a few hundred small tensors), so that the only device work of loading a model is
        # the packing kernels of this library and plain copies
        sd = {k: v.to(dev) for k, v in layouts.fold_weight_norm({k: v.detach().cpu() for k, v in self.state_dict().items()}).items()}
        n = self._names()
        prec = self.precision
        pk = {}
        pk["pre"] = ops.ConvLayer(sd[n["pre"] + ".weight"], sd[n["pre"] + ".bias"], padding=(self.kernel_size - 1) // 2,
                                  precision=prec)
        for i, (u, k) in enumerate(zip(self.upsample_scales, self.upsample_kernel_sizes)):
            if k != 2 * u or u % 2:
                raise ops._lib.EngineError("the engine supports upsampling layers with kernel == 2*stride, stride even")
            pk[f"up{i}"] = ops.ConvLayer(sd[n["up"].format(i) + ".weight"], sd[n["up"].format(i) + ".bias"],
                                         transposed_stride=u, precision=prec)
            for j, kr in enumerate(self.resblock_kernel_sizes):
                blk = i * len(self.resblock_kernel_sizes) + j
                for m, d in enumerate(self.resblock_dilations[j]):
                    c1, c2 = n["c1"].format(blk, m), n["c2"].format(blk, m)
                    pk[f"b{blk}.c1.{m}"] = ops.ConvLayer(sd[c1 + ".weight"], sd[c1 + ".bias"], dilation=d,
                                                         padding=(kr - 1) // 2 * d, precision=prec)
                    pk[f"b{blk}.c2.{m}"] = ops.ConvLayer(sd[c2 + ".weight"], sd[c2 + ".bias"], dilation=1,
                                                         padding=(kr - 1) // 2, precision=prec)
                    if n["act"] is not None:
                        for which, idx in (("a1", 2 * m), ("a2", 2 * m + 1)):
                            a = n["act"].format(blk, idx)
                            pk[f"b{blk}.{which}.{m}"] = (sd[a + ".alpha"].float().contiguous(), sd[a + ".beta"].float().contiguous())
                    if self._pair_fused(self._stage_channels(i), kr):
                        pk[f"b{blk}.pair.{m}"] = ops.ResPair(pk[f"b{blk}.c1.{m}"], pk[f"b{blk}.c2.{m}"],
                                                             pk.get(f"b{blk}.a1.{m}"), pk.get(f"b{blk}.a2.{m}"))
        pk["post"] = ops.ConvLayer(sd[n["post"] + ".weight"], sd[n["post"] + ".bias"], padding=(self.kernel_size - 1) // 2,
                                   precision=prec)
        if n["act_post"] is not None:
            pk["post_act"] = (sd[n["act_post"] + ".alpha"].float().contiguous(), sd[n["act_post"] + ".beta"].float().contiguous())
        self._packed = pk
        self._buffers_cache = {}

    def _pair_fused(self, channels, kernel):
        """Which residual pairs run as one tb200_respair launch (measured per shape on B200 at config-2 sizes,
        profiles/r2_pair_fused_vs_two_launch.txt).  The fused kernel keeps an X tile, two operand tiles (and for
        BigVGAN the tile between the convs) in shared memory: long kernels (K = 11: 50 halo rows per tile, weights
        streamed from L2 per tile) and the 128-channel pairs with K >= 7 are faster as two launches."""
        if not (self.fuse_pairs and ops.ResPair.supported(channels, self.precision)):
            return False
        if self.activation_dtype != "f16":   # fp32 streams double the X tile: smaller tiles, no gain over two launches
            return False
        if channels == 128 and kernel > 3:
            return False
        snake = self._names()["act"] is not None
        if snake:
            return (kernel <= 7 and channels != 64) or (kernel == 3 and channels == 64)
        return kernel <= 7 or channels == 64

    # -- workspace ---------------------------------------------------------------------------
    def _workspace(self, b, frames, dev):
        """Stage buffers for a (b, frames) batch, carved out of flat per-name allocations that only ever grow: a
        different bucket shape re-uses the same memory without a fill (every kernel masks by the utterance lengths,
        nothing is read unmasked past them -- tests/test_vocoder_gpu.py::test_workspace_garbage_is_never_read)."""
        key = (b, frames, str(dev))
        ws = self._buffers_cache.get(key)
        if ws is None:
            self._buffers_cache = {k: v for k, v in self._buffers_cache.items() if k == "flat"}
            flat = self._buffers_cache.setdefault("flat", {})

            def carve(name, shape, dtype):
                need = shape[0] * shape[1] * shape[2]
                buf = flat.get(name)
                if buf is None or buf.numel() < need or buf.dtype != dtype or buf.device != dev:
                    flat.pop(name, None)
                    buf = flat[name] = torch.zeros(need, dtype=dtype, device=dev)
                return buf[:need].view(shape)

            # residual stream: fp32 (default) or fp16 (activation_dtype="f16": halves the HBM bytes per element; every
            # value is rounded to 11 bits once more per layer -- measured SNR in tests/test_vocoder_gpu.py)
            sdt = torch.float16 if self.activation_dtype == "f16" else torch.float32
            pad = _pad8   # 16-byte aligned rows for either type (vector loads, bulk copies of tb200_respair)
            ws = {"h": carve("h", (b, self.channels, pad(frames)), sdt)}
            length = frames
            for i, u in enumerate(self.upsample_scales):
                length *= u
                c = self._stage_channels(i)
                for name in ("up", "r0", "r1", "sum"):
                    ws[f"{name}{i}"] = carve(f"{name}{i}", (b, c, pad(length)), sdt)
                # value between the two convs of a residual pair: an MMA operand only -> fp16 in the
                # fp16-operand mode, fp32 otherwise
                tdt = torch.float16 if self.precision == "f16" else torch.float32
                # row pitch a multiple of 8 elements: the lane=channel snake staging reads 16-byte groups
                ws[f"t{i}"] = carve(f"t{i}", (b, c, _pad8(length) + 8), tdt)
            self._buffers_cache[key] = ws
        return ws

    # -- the generator -----------------------------------------------------------------------
    # workspace bytes a single forward_batch call may take; larger batches run in chunks
    max_workspace_bytes = 64 << 30

    def _max_chunk(self, frames):
        """Largest batch one kernel pass takes: every stage tensor must stay below 2^31 elements (the per-layer
        epilogue's fast path uses 32-bit element offsets) and the workspace inside `max_workspace_bytes`."""
        esz = 2 if self.activation_dtype == "f16" else 4
        length, per_utt_elems, per_utt_bytes = frames, self.channels * _pad8(frames), self.channels * _pad8(frames) * esz
        for i, u in enumerate(self.upsample_scales):
            length *= u
            n = self._stage_channels(i) * (_pad8(length) + 8)
            per_utt_elems = max(per_utt_elems, n)
            per_utt_bytes += n * (4 * esz + (2 if self.precision == "f16" else 4))
        return max(1, min(((1 << 31) - 1) // per_utt_elems, self.max_workspace_bytes // per_utt_bytes))

    @torch.no_grad()
    def forward_batch(self, c, lengths=None):
        """c (B,80,F) fp32 CUDA; lengths (B) frames per utterance (int tensor) or None.
        Returns a new tensor wave (B, F*prod(scales)) fp32; samples past lengths[b]*384 are zero."""
        b, _, frames = c.shape
        cap = self._max_chunk(frames)
        if b <= cap:
            return self._forward_chunk(c, lengths)
        parts = [self._forward_chunk(c[i:i + cap], None if lengths is None else lengths[i:i + cap]) for i in range(0, b, cap)]
        return torch.cat(parts, dim=0)

    def _forward_chunk(self, c, lengths=None):
        if self._packed is None:
            self.remove_weight_norm()
        pk = self._packed
        b, _, frames = c.shape
        dev = c.device
        c = c.contiguous().float()
        ws = self._workspace(b, frames, dev)
        len_t = lengths.to(device=dev, dtype=torch.int32).contiguous() if lengths is not None else None
        res_act = ACT_AA_SNAKEBETA if self._names()["act"] is not None else ACT_LEAKY_RELU

        h = pk["pre"](c, len_t, ws["h"], l_in_max=frames)
        length = frames
        for i, u in enumerate(self.upsample_scales):
            up = pk[f"up{i}"](h, len_t, ws[f"up{i}"], l_in_max=length,
                              act=self.up_act, slope=0.1)
            length *= u
            if len_t is not None:
                len_t = len_t * u
            total = ws[f"sum{i}"]
            nblk = len(self.resblock_kernel_sizes)
            for j in range(nblk):
                blk = i * nblk + j
                cur = up
                ndil = len(self.resblock_dilations[j])
                for m in range(ndil):
                    last = m == ndil - 1
                    pair = pk.get(f"b{blk}.pair.{m}")
                    if pair is not None:
                        if last:  # fold the branch into the multi-receptive-field mean
                            pair(cur, len_t, total, l_max=length, slope=0.1, out_alpha=1.0 / nblk, res_beta=1.0 / nblk,
                                 accumulate=j > 0)
                        else:
                            cur = pair(cur, len_t, ws[f"r{m % 2}{i}"], l_max=length, slope=0.1)
                        continue
                    a1 = pk.get(f"b{blk}.a1.{m}", (None, None))
                    a2 = pk.get(f"b{blk}.a2.{m}", (None, None))
                    xt = pk[f"b{blk}.c1.{m}"](cur, len_t, ws[f"t{i}"], l_in_max=length, act=res_act, slope=0.1,
                                              alpha=a1[0], beta=a1[1])
                    if last:  # fold the branch into the multi-receptive-field mean
                        pk[f"b{blk}.c2.{m}"](xt, len_t, total, l_in_max=length, act=res_act, slope=0.1, alpha=a2[0],
                                             beta=a2[1], out_alpha=1.0 / nblk, residual=cur, res_beta=1.0 / nblk,
                                             accumulate=j > 0)
                    else:
                        dst = ws[f"r{m % 2}{i}"]
                        cur = pk[f"b{blk}.c2.{m}"](xt, len_t, dst, l_in_max=length, act=res_act, slope=0.1, alpha=a2[0],
                                                   beta=a2[1], residual=cur)
            h = total
        pa = pk.get("post_act", (None, None))
        # the result is a fresh tensor (the intermediates live in the cached workspace and are overwritten by the next call)
        wave = torch.zeros((b, 1, _pad4(length)), dtype=torch.float32, device=dev)
        pk["post"](h, len_t, wave, l_in_max=length, act=self.post_act, slope=0.01, alpha=pa[0], beta=pa[1], out_act=OUT_TANH)
        return wave[:, 0, :length]

    def forward(self, c, *args, **kwargs):
        raise NotImplementedError


class HiFiGANGenerator(_GeneratorBase):
    """InferenceAvocodo.py:6-96."""

    up_act = ACT_LEAKY_RELU
    post_act = ACT_LEAKY_RELU

    def __init__(self, path_to_weights, in_channels=80, out_channels=1, channels=512, kernel_size=7,
                 upsample_scales=(8, 6, 4, 2), upsample_kernel_sizes=(16, 12, 8, 4), resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilations=((1, 3, 5), (1, 3, 5), (1, 3, 5)), use_additional_convs=True, bias=True,
                 nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1},
                 use_weight_norm=True, precision="f16", activation_dtype="f32", fuse_pairs=True):
        super().__init__()
        if not (use_additional_convs and bias and use_weight_norm and nonlinear_activation == "LeakyReLU"
                and nonlinear_activation_params.get("negative_slope", 0.1) == 0.1 and out_channels == 1):
            raise ops._lib.EngineError("unsupported HiFiGAN variant (engine covers the reference's inference configuration)")
        self.in_channels, self.channels, self.kernel_size = in_channels, channels, kernel_size
        self.upsample_scales, self.upsample_kernel_sizes = tuple(upsample_scales), tuple(upsample_kernel_sizes)
        self.resblock_kernel_sizes, self.resblock_dilations = tuple(resblock_kernel_sizes), tuple(map(tuple, resblock_dilations))
        lay, alias = layouts.hifigan_layout(in_channels, out_channels, channels, kernel_size, upsample_scales,
                                            upsample_kernel_sizes, resblock_kernel_sizes, resblock_dilations)
        layouts.attach(self, lay, alias)
        self._init_engine(precision, activation_dtype, fuse_pairs)
        if path_to_weights is not None:
            self.load_state_dict(torch.load(path_to_weights, map_location="cpu")["generator"])

    def _names(self):
        return dict(pre="input_conv", up="upsamples.{}.1", c1="blocks.{}.convs1.{}.1", c2="blocks.{}.convs2.{}.1",
                    act=None, post="output_conv.1", act_post=None)

    def forward(self, c, normalize_before=False):
        """c (80,T) -> (T*384,)  (InferenceAvocodo.py:69-80)."""
        if normalize_before:
            c = (c - self.mean) / self.scale
        return self.forward_batch(c.unsqueeze(0)).squeeze()


class BigVGAN(_GeneratorBase):
    """InferenceBigVGAN.py:19-121."""

    up_act = ACT_NONE
    post_act = ACT_AA_SNAKEBETA

    def __init__(self, path_to_weights, num_mels=80, upsample_initial_channel=512, upsample_rates=(8, 6, 4, 2),
                 upsample_kernel_sizes=(16, 12, 8, 4), resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), precision="f16", activation_dtype="f32",
                 fuse_pairs=True):
        super().__init__()
        self.in_channels, self.channels, self.kernel_size = num_mels, upsample_initial_channel, 7
        self.upsample_scales, self.upsample_kernel_sizes = tuple(upsample_rates), tuple(upsample_kernel_sizes)
        self.resblock_kernel_sizes, self.resblock_dilations = tuple(resblock_kernel_sizes), tuple(map(tuple, resblock_dilation_sizes))
        self.num_kernels, self.num_upsamples = len(resblock_kernel_sizes), len(upsample_rates)
        sd = torch.load(path_to_weights, map_location="cpu")["generator"] if path_to_weights is not None else None
        # alias_free_torch may or may not have stored its filters as persistent buffers: accept both
        has_filters = sd is None or any(k.endswith(".filter") for k in sd)
        lay, alias = layouts.bigvgan_layout(num_mels, upsample_initial_channel, upsample_rates, upsample_kernel_sizes,
                                            resblock_kernel_sizes, resblock_dilation_sizes, filter_buffers=has_filters)
        layouts.attach(self, lay, alias)
        self._init_engine(precision, activation_dtype, fuse_pairs)
        if sd is not None:
            self.load_state_dict(sd)

    def _names(self):
        return dict(pre="conv_pre", up="ups.{}.0", c1="resblocks.{}.convs1.{}", c2="resblocks.{}.convs2.{}",
                    act="resblocks.{}.activations.{}.act", post="conv_post", act_post="activation_post.act")

    def forward(self, x):
        """x (80,T) -> (T*384,)  (InferenceBigVGAN.py:72-95)."""
        return self.forward_batch(x.unsqueeze(0)).squeeze()
