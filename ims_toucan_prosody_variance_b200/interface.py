"""Drop-in `ToucanTTSInterface` / `UtteranceCloner` entry points over the B200 engine.

Same constructor and method signatures as InferenceInterfaces/ToucanTTSInterface.py:20-310 and
InferenceInterfaces/UtteranceCloner.py:20-200 of the reference.  The acoustic model and the vocoder are the
engine's modules; the components that sit BEFORE the hot path -- the articulatory text frontend
(Preprocessing/TextFrontend.py), the style-embedding network and the audio preprocessor -- are outside
the engine's scope and are injected: pass `text2phone=` / `style_embedding_function=` /
`audio_preprocessor=`, or have the reference's IMS-Toucan checkout importable (its own classes are then used
unchanged).  Plotting (`view=True`) is not part of the engine.
"""
import itertools
import os
import wave as _wave

import torch

from ._lib import EngineError
from .frontend import PhoneTensoriser
from .pipeline import SAMPLE_RATE, TextToWave
from .toucantts import ToucanTTS
from .vocoder import BigVGAN, HiFiGANGenerator

# Preprocessing/TextFrontend.py:490-524 (get_language_id)
LANGUAGE_IDS = {"de": 1, "el": 2, "es": 3, "fi": 4, "ru": 5, "hu": 6, "nl": 7, "fr": 8, "pt": 9, "pl": 10, "it": 11,
                "en": 12, "cmn": 13, "vi": 14, "uk": 15, "fa": 16, "pt-br": 17}


def get_language_id(language):
    return torch.LongTensor([LANGUAGE_IDS[language]])


def float2pcm(sig):
    """Utility/utils.py:20-33 for int16: (sig * 32768).clip(-32768, 32767).astype(int16), on the tensor's device."""
    return (sig * 32768.0).clamp(-32768, 32767).to(torch.int16)


def _reference_frontend(language):
    try:
        from Preprocessing.TextFrontend import ArticulatoryCombinedTextFrontend  # the user's IMS-Toucan checkout
    except Exception as exc:  # pragma: no cover - depends on the deployment
        raise EngineError("no text frontend: pass text2phone= (an object with string_to_tensor(text, input_phonemes=...)) "
                          "or make IMS-Toucan's Preprocessing.TextFrontend importable") from exc
    return ArticulatoryCombinedTextFrontend(language=language, add_silence_to_end=True)


class ToucanTTSInterface(torch.nn.Module):
    """ToucanTTSInterface.py:20-310."""

    def __init__(self, device="cuda", tts_model_path=None, embedding_model_path=None, vocoder_model_path=None,
                 faster_vocoder=True, language="en", text2phone=None, style_embedding_function=None,
                 audio_preprocessor=None, precision="tf32", vocoder_precision="f16"):
        super().__init__()
        self.device = device
        if tts_model_path is None or vocoder_model_path is None:
            raise EngineError("tts_model_path and vocoder_model_path must name checkpoint files "
                              "(the reference's MODELS_DIR shorthands need its model downloader)")
        self._frontend_factory = (lambda lang: text2phone) if text2phone is not None else _reference_frontend
        self._text2phone = text2phone
        self._tensorisers = {}
        self._language = language
        checkpoint = torch.load(tts_model_path, map_location="cpu")
        self.use_lang_id = True
        try:  # same fall-through as ToucanTTSInterface.py:53-62
            self.phone2mel = ToucanTTS(weights=checkpoint["model"], precision=precision)
        except RuntimeError:
            try:
                self.use_lang_id = False
                self.phone2mel = ToucanTTS(weights=checkpoint["model"], lang_embs=None, precision=precision)
            except RuntimeError:
                self.phone2mel = ToucanTTS(weights=checkpoint["model"], lang_embs=None, utt_embed_dim=None, precision=precision)
        self.phone2mel = self.phone2mel.to(torch.device(device))
        self.phone2mel.store_inverse_all()
        self.style_embedding_function = style_embedding_function
        self.audio_preprocessor = audio_preprocessor
        if style_embedding_function is not None and embedding_model_path is not None:
            self.style_embedding_function.load_state_dict(torch.load(embedding_model_path, map_location="cpu")["style_emb_func"])
            self.style_embedding_function.to(device)
        cls = HiFiGANGenerator if faster_vocoder else BigVGAN
        self.mel2wav = cls(path_to_weights=vocoder_model_path, precision=vocoder_precision).to(torch.device(device))
        self.mel2wav.remove_weight_norm()
        self.default_utterance_embedding = checkpoint["default_emb"].to(device)
        self.lang_id = get_language_id(language) if self.use_lang_id else None
        self.engine = TextToWave(self.phone2mel, self.mel2wav)
        self.eval()

    # -- frontend ---------------------------------------------------------------------------
    @property
    def text2phone(self):
        if self._text2phone is None:
            self._text2phone = self._frontend_factory(self._language)
        return self._text2phone

    def set_utterance_embedding(self, path_to_reference_audio="", embedding=None):
        """ToucanTTSInterface.py:103-115."""
        if embedding is not None:
            self.default_utterance_embedding = embedding.squeeze().to(self.device)
            return
        if self.style_embedding_function is None or self.audio_preprocessor is None:
            raise EngineError("embedding from audio needs style_embedding_function= and audio_preprocessor= (outside the engine)")
        import soundfile
        wave, sr = soundfile.read(path_to_reference_audio)
        spec = self.audio_preprocessor.audio_to_mel_spec_tensor(wave).transpose(0, 1)
        spec_len = torch.LongTensor([len(spec)])
        self.default_utterance_embedding = self.style_embedding_function(spec.unsqueeze(0).to(self.device),
                                                                         spec_len.unsqueeze(0).to(self.device)).squeeze()

    def set_language(self, lang_id):
        self.set_phonemizer_language(lang_id=lang_id)
        self.set_accent_language(lang_id=lang_id)

    def set_phonemizer_language(self, lang_id):
        self._language = lang_id
        self._text2phone = self._frontend_factory(lang_id)

    def set_accent_language(self, lang_id):
        self.lang_id = get_language_id(lang_id).to(self.device) if self.use_lang_id else None

    # -- synthesis --------------------------------------------------------------------------
    def forward(self, text, view=False, duration_scaling_factor=1.0, pitch_variance_scale=1.0, energy_variance_scale=1.0,
                pause_duration_scaling_factor=1.0, durations=None, pitch=None, energy=None, input_is_phones=False,
                return_plot_as_filepath=False):
        """ToucanTTSInterface.py:132-229: one sentence -> 1-D waveform on `device`."""
        if view or return_plot_as_filepath:
            raise EngineError("plotting is not part of the engine")
        with torch.inference_mode():
            phones = self.text2phone.string_to_tensor(text, input_phonemes=input_is_phones).to(torch.device(self.device))
            mel = self.phone2mel(phones, utterance_embedding=self.default_utterance_embedding, durations=durations,
                                 pitch=pitch, energy=energy, lang_id=self.lang_id,
                                 duration_scaling_factor=duration_scaling_factor, pitch_variance_scale=pitch_variance_scale,
                                 energy_variance_scale=energy_variance_scale,
                                 pause_duration_scaling_factor=pause_duration_scaling_factor)
            return self.mel2wav(mel.transpose(0, 1))

    def _tensorise(self, text_list, input_is_phones):
        """Sentences -> list of (T_i, 62) feature tensors.  With the reference's frontend (anything that exposes its
        `phone_to_vector` table) the character loop of string_to_tensor (TextFrontend.py:213-288) is replaced by the
        vectorised lookup of frontend.PhoneTensoriser over all sentences at once, staged in pinned memory; the
        grapheme-to-phoneme step stays the frontend's `get_phone_string`.  Other frontends: their own string_to_tensor."""
        fe = self.text2phone
        if hasattr(fe, "phone_to_vector") and (input_is_phones or hasattr(fe, "get_phone_string")):
            tz = self._tensorisers.get(id(fe))
            if tz is None:
                try:
                    tz = PhoneTensoriser.from_frontend(fe)
                except ImportError:
                    tz = False                     # no feature index available: keep the frontend's own loop
                self._tensorisers = {id(fe): tz}
            if tz:
                strings = list(text_list) if input_is_phones else [
                    fe.get_phone_string(text=t, include_eos_symbol=True, for_feature_extraction=True) for t in text_list]
                return tz.batch(strings)[2]
        return [fe.string_to_tensor(t, input_phonemes=input_is_phones) for t in text_list]

    def forward_batch(self, text_list, duration_scaling_factor=1.0, pitch_variance_scale=1.0, energy_variance_scale=1.0,
                      pause_duration_scaling_factor=1.0, input_is_phones=False, noise=None):
        """Additive batched entry: all sentences in one ragged engine call; returns a list of waveforms."""
        phones = self._tensorise(text_list, input_is_phones)
        return self.engine.synthesize(phones, self.default_utterance_embedding.cpu(),
                                      lang_ids=int(self.lang_id) if self.lang_id is not None else None, noise=noise,
                                      duration_scaling_factor=duration_scaling_factor, pitch_variance_scale=pitch_variance_scale,
                                      energy_variance_scale=energy_variance_scale,
                                      pause_duration_scaling_factor=pause_duration_scaling_factor)

    def read_to_file(self, text_list, file_location, duration_scaling_factor=1.0, pitch_variance_scale=1.0,
                     energy_variance_scale=1.0, silent=False, dur_list=None, pitch_list=None, energy_list=None,
                     increased_compatibility_mode=False):
        """ToucanTTSInterface.py:231-285.  Sentences without external prosody are synthesised in ONE batched engine
        call; concatenation, silence insertion and the PCM16 conversion happen on the GPU; one D2H copy at the end."""
        dur_list, pitch_list, energy_list = dur_list or [], pitch_list or [], energy_list or []
        items = [it for it in itertools.zip_longest(text_list, dur_list, pitch_list, energy_list) if it[0] is not None and it[0].strip() != ""]
        waves = [None] * len(items)
        plain = [k for k, it in enumerate(items) if it[1] is None and it[2] is None and it[3] is None]
        if plain:
            if not silent:
                for k in plain:
                    print("Now synthesizing: {}".format(items[k][0]))
            for k, w in zip(plain, self.forward_batch([items[k][0] for k in plain], duration_scaling_factor, pitch_variance_scale,
                                                      energy_variance_scale)):
                waves[k] = w
        for k, (text, d, p, e) in enumerate(items):
            if waves[k] is None:
                if not silent:
                    print("Now synthesizing: {}".format(text))
                waves[k] = self(text, durations=d.to(self.device) if d is not None else None,
                                pitch=p.to(self.device) if p is not None else None,
                                energy=e.to(self.device) if e is not None else None,
                                duration_scaling_factor=duration_scaling_factor, pitch_variance_scale=pitch_variance_scale,
                                energy_variance_scale=energy_variance_scale)
        silence = torch.zeros([10600], device=self.device)
        parts = [silence]
        for w in waves:
            parts += [w.reshape(-1), silence]
        wav = torch.cat(parts, 0)
        if increased_compatibility_mode:  # 16-bit 48 kHz: every sample twice (ToucanTTSInterface.py:282-283)
            pcm = float2pcm(wav.repeat_interleave(2)).cpu().numpy()
            _write_pcm16(file_location, pcm, 48000)
        else:
            try:
                import soundfile
                soundfile.write(file=file_location, data=wav.cpu().numpy(), samplerate=SAMPLE_RATE)
            except ImportError:
                _write_pcm16(file_location, float2pcm(wav).cpu().numpy(), SAMPLE_RATE)
        return wav

    def read_aloud(self, text, view=False, duration_scaling_factor=1.0, pitch_variance_scale=1.0, energy_variance_scale=1.0,
                   blocking=False, increased_compatibility_mode=False):
        """ToucanTTSInterface.py:287-310 (needs the optional `sounddevice` package)."""
        if text.strip() == "":
            return
        import sounddevice
        wav = self(text, view, duration_scaling_factor=duration_scaling_factor, pitch_variance_scale=pitch_variance_scale,
                   energy_variance_scale=energy_variance_scale).cpu()
        wav = torch.cat((wav, torch.zeros([12000])), 0).numpy()
        if increased_compatibility_mode:
            wav = float2pcm(torch.from_numpy(wav).repeat_interleave(2)).numpy()
            sounddevice.play(wav, samplerate=48000)
        else:
            sounddevice.play(wav, samplerate=SAMPLE_RATE)
        if blocking:
            sounddevice.wait()


def _write_pcm16(path, pcm, rate):
    with _wave.open(path, "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(rate)
        f.writeframes(pcm.tobytes())


class UtteranceCloner:
    """UtteranceCloner.py:20-200.  `extract_prosody` (aligner fine-tuning, pitch / energy extraction) sits before the
    hot path and is outside the engine: supply `prosody_extractor(transcript, audio_path, lang) ->
    (durations, pitch, energy, start_silence_frames, end_silence_frames)`; the synthesis with the overridden prosody
    runs on the engine."""

    def __init__(self, model_id, device, language="en", speed_over_quality=False, tts=None, prosody_extractor=None, **tts_kwargs):
        self.tts = tts if tts is not None else ToucanTTSInterface(device=device, tts_model_path=model_id,
                                                                  faster_vocoder=speed_over_quality, language=language, **tts_kwargs)
        self.device = device
        self.prosody_extractor = prosody_extractor

    def extract_prosody(self, transcript, ref_audio_path, lang="de", on_line_fine_tune=True):
        if self.prosody_extractor is None:
            raise EngineError("extract_prosody needs prosody_extractor= (aligner / pitch / energy extraction is outside the engine)")
        return self.prosody_extractor(transcript, ref_audio_path, lang)

    def clone_utterance(self, path_to_reference_audio_for_intonation, path_to_reference_audio_for_voice,
                        transcription_of_intonation_reference, filename_of_result=None, lang="de"):
        """UtteranceCloner.py:147-167."""
        self.tts.set_utterance_embedding(path_to_reference_audio=path_to_reference_audio_for_voice)
        duration, pitch, energy, sil_start, sil_end = self.extract_prosody(transcription_of_intonation_reference,
                                                                           path_to_reference_audio_for_intonation, lang=lang)
        self.tts.set_language(lang)
        return self._speak(transcription_of_intonation_reference, duration, pitch, energy, sil_start, sil_end, filename_of_result)

    def biblical_accurate_angel_mode(self, path_to_reference_audio_for_intonation, transcription_of_intonation_reference,
                                     list_of_speaker_references_for_ensemble, filename_of_result=None, lang="de"):
        """UtteranceCloner.py:169-200: several voices, one intonation, averaged."""
        prev = self.tts.default_utterance_embedding.clone().detach()
        duration, pitch, energy, sil_start, sil_end = self.extract_prosody(transcription_of_intonation_reference,
                                                                           path_to_reference_audio_for_intonation, lang=lang)
        self.tts.set_language(lang)
        speeches = []
        for path in list_of_speaker_references_for_ensemble:
            self.tts.set_utterance_embedding(path_to_reference_audio=path)
            speeches.append(self.tts(transcription_of_intonation_reference, durations=duration.clone(), pitch=pitch.clone(),
                                     energy=energy.clone()))
        out = self._finish(torch.stack(speeches).mean(dim=0), sil_start, sil_end, filename_of_result)
        self.tts.default_utterance_embedding = prev.to(self.device)
        return out

    def _speak(self, text, duration, pitch, energy, sil_start, sil_end, filename):
        speech = self.tts(text, view=False, durations=duration, pitch=pitch, energy=energy)
        return self._finish(speech, sil_start, sil_end, filename)

    def _finish(self, speech, sil_start, sil_end, filename):
        start = torch.zeros([sil_start * 3], device=speech.device)
        end = torch.zeros([sil_end * 3], device=speech.device)
        cloned = torch.cat((start, speech, end), dim=0)
        if filename is not None:   # UtteranceCloner.py:165-166 writes with soundfile; PCM16 WAV when it is not installed
            try:
                import soundfile
                soundfile.write(file=filename, data=cloned.cpu().numpy(), samplerate=SAMPLE_RATE)
            except ImportError:
                _write_pcm16(filename, float2pcm(cloned).cpu().numpy(), SAMPLE_RATE)
        return cloned.cpu().numpy()
