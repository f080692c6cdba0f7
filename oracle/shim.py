"""Import shim for the LIVE reference (TEST INFRASTRUCTURE, authoring container only).

Makes ``/root/reference`` importable without its audio / plotting / phonemizer
dependencies: the reference imports librosa, matplotlib,
soundfile, ... at module top (Utility/utils.py:8-17,
Preprocessing/TextFrontend.py:8-10, InferenceInterfaces/ToucanTTSInterface.py:4-7)
but calls none of them on the hot path.  Nothing here runs on the GPU box
(``/root/reference`` does not exist there); ``available()`` says so.
"""
import importlib.machinery
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("TOUCAN_REFERENCE_ROOT", "/root/reference")

_STUBS = ["librosa", "librosa.display", "librosa.core", "matplotlib", "matplotlib.pyplot", "matplotlib.lines",
          "dragonmapper", "dragonmapper.transcriptions", "phonemizer", "phonemizer.backend", "pypinyin",
          "sounddevice", "soundfile", "pyloudnorm"]


class _Stub(types.ModuleType):
    """Module whose every attribute is another callable stub."""

    def __init__(self, name):
        super().__init__(name)
        self.__spec__ = importlib.machinery.ModuleSpec(name, None)
        self.__path__ = []

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        child = _Stub(self.__name__ + "." + item)
        setattr(self, item, child)
        return child

    def __call__(self, *args, **kwargs):
        return _Stub(self.__name__ + "()")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "InferenceInterfaces"))


_installed = set()
FRONTEND_STUBS = ["dragonmapper", "dragonmapper.transcriptions", "phonemizer", "phonemizer.backend", "pypinyin"]


def install(stubs=None):
    """Put the reference and the restated alias_free_torch on sys.path, register stubs (all of _STUBS by default; a
    subset for callers that only need part of the reference -- a stubbed librosa in sys.modules changes what
    third-party packages such as transformers believe is installed, so tests that do not need it do not register it)."""
    if not available():
        raise RuntimeError(f"live reference not found at {REFERENCE_ROOT} (it never exists on the GPU box)")
    for name in (_STUBS if stubs is None else stubs):
        if name not in _installed and name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
        _installed.add(name)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)  # oracle/alias_free_torch wins over any installed one
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    warnings.filterwarnings("ignore", message=".*weight_norm.*")


def reference_classes():
    """Return the reference's inference + train-side classes for the hot path."""
    install()
    from InferenceInterfaces.InferenceArchitectures.InferenceAvocodo import HiFiGANGenerator as InfHiFiGAN
    from InferenceInterfaces.InferenceArchitectures.InferenceBigVGAN import BigVGAN as InfBigVGAN
    from InferenceInterfaces.InferenceArchitectures.InferenceToucanTTS import ToucanTTS as InfToucanTTS
    from TrainingInterfaces.Spectrogram_to_Wave.BigVGAN.BigVGAN import BigVGAN as TrainBigVGAN
    from TrainingInterfaces.Spectrogram_to_Wave.HiFiGAN.HiFiGAN import HiFiGANGenerator as TrainHiFiGAN
    from TrainingInterfaces.Text_to_Spectrogram.ToucanTTS.ToucanTTS import ToucanTTS as TrainToucanTTS
    return dict(InfHiFiGAN=InfHiFiGAN, InfBigVGAN=InfBigVGAN, InfToucanTTS=InfToucanTTS,
                TrainBigVGAN=TrainBigVGAN, TrainHiFiGAN=TrainHiFiGAN, TrainToucanTTS=TrainToucanTTS)
