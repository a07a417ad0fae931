"""B200-native spectral front/back end of audio-visual speech enhancement.

Hot path only: mix at SNR -> STFT -> mel -> dB -> AV-aligned slices, and
mel -> linear -> ISTFT at predict time, as hand-written sm_100a CUDA behind a C ABI
(include/avse_b200.h), mirrored here with the reference's own Python signatures.

The directory name carries a hyphen (it is the reference's repository name); import it with
`importlib.import_module("audio-visual-speech-enhancement_b200")` or through the top-level
alias module `avse_b200`.
"""
from . import build  # noqa: F401
from . import _native  # noqa: F401
from .mediaio_compat import AudioSignal, AudioMixer  # noqa: F401

__all__ = ["build", "_native", "AudioSignal", "AudioMixer", "engine", "data_processor"]


def __getattr__(name):
    # engine / data_processor import torch; keep the package importable for build-only use
    if name in ("engine", "data_processor"):
        import importlib
        return importlib.import_module(__name__ + "." + name)
    raise AttributeError(name)
