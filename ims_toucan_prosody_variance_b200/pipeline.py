"""Batched text -> 24 kHz waveform engine: ToucanTTS acoustic model + vocoder generator in one call.

This is the additive batched entry the reference lacks (its `read_to_file` loops sentence by sentence,
ToucanTTSInterface.py:269-280): phoneme tensors in, waveforms out, everything in between stays on the GPU
(the acoustic model hands the mel to the vocoder in the NCL layout it consumes).  One host sync per batch
(the frame counts that size the decoder buffers).
"""
import torch

from . import sharding
from ._lib import EngineError

SAMPLES_PER_FRAME = 384
SAMPLE_RATE = 24_000


class TextToWave:
    def __init__(self, phone2mel, mel2wav, max_batch=128, max_padding_ratio=None):
        self.phone2mel, self.mel2wav = phone2mel, mel2wav
        self.max_batch, self.max_padding_ratio = max_batch, max_padding_ratio

    @torch.no_grad()
    def synthesize_padded(self, text_tensors, text_lengths, utterance_embedding, lang_ids=None, noise=None, **prosody):
        """One ragged batch: text_tensors (B,T,62), text_lengths (B), utterance_embedding (B,E) ->
        (wave (B, Lmax) fp32 CUDA, wave_lengths (B) int32, acoustic result dict)."""
        r = self.phone2mel.synthesize_batch(text_tensors, text_lengths, utterance_embedding=utterance_embedding,
                                            lang_ids=lang_ids, noise=noise, **prosody)
        m_max = 2 * (int(r["frames_host"].max()) // 2)
        mel = r["mel_ncl"][:, :, :m_max] if r["mel_ncl"].shape[2] != m_max else r["mel_ncl"]
        wave = self.mel2wav.forward_batch(mel, r["mel_lengths"])
        return wave, r["mel_lengths"] * SAMPLES_PER_FRAME, r

    @torch.no_grad()
    def synthesize(self, texts, utterance_embeddings, lang_ids=None, noise="device", device=None, **prosody):
        """texts: list of (T_i,62) tensors; utterance_embeddings: (N,E) or a single (E,) vector; lang_ids: (N,) or
        int or None.  Utterances are sorted by length and synthesised in batches of at most `max_batch`.
        Returns a list of N 1-D fp32 CUDA waveforms in input order."""
        n = len(texts)
        if n == 0:
            return []
        device = device or next(self.phone2mel.parameters()).device
        if device.type != "cuda":
            raise EngineError("toucan_b200 has no CPU path")
        emb = torch.as_tensor(utterance_embeddings)
        if emb.dim() == 1:
            emb = emb.unsqueeze(0).expand(n, -1)
        if lang_ids is not None and not torch.is_tensor(lang_ids):
            lang_ids = torch.full((n,), int(lang_ids), dtype=torch.int64)
        lengths = [int(t.shape[0]) for t in texts]
        out = [None] * n
        for batch in sharding.bucket_by_length(range(n), lengths, self.max_batch, self.max_padding_ratio):
            t_max = max(lengths[i] for i in batch)
            # padded batch staged in pinned memory: the host-to-device copy below is asynchronous
            x = torch.zeros((len(batch), t_max, texts[batch[0]].shape[1]), dtype=torch.float32, pin_memory=True)
            for row, i in enumerate(batch):
                x[row, :lengths[i]] = texts[i]
            idx = torch.as_tensor(batch, dtype=torch.int64)
            wave, wlen, _ = self.synthesize_padded(
                x.to(device, non_blocking=True), torch.tensor([lengths[i] for i in batch], dtype=torch.int32),
                emb[idx].to(device), lang_ids=lang_ids[idx] if lang_ids is not None else None, noise=noise, **prosody)
            wl = wlen.cpu()
            for row, i in enumerate(batch):
                out[i] = wave[row, :int(wl[row])]
        return out
