"""Stand-ins for the two mediaio classes the reference's audio path uses (mediaio is a separate,
un-vendored package: github.com/avivga/mediaio).  Only the surface data_processor.py touches is
provided (dp:36-42, dp:78, dp:116, dp:122-133; speech_enhancer.py:176-177).  Host-side numpy;
the arithmetic that matters (variance, mix) runs on the GPU in engine.SpectralEngine."""
from __future__ import annotations

import numpy as np


class AudioSignal(object):

    def __init__(self, data, sample_rate):
        self._data = np.array(data, copy=True)
        self._sample_rate = int(sample_rate)

    @staticmethod
    def from_wav_file(wave_file_path):
        from scipy.io import wavfile
        sample_rate, data = wavfile.read(wave_file_path)
        return AudioSignal(data, sample_rate)

    def save_to_wav_file(self, wave_file_path, sample_type=np.int16):
        from scipy.io import wavfile
        info = np.iinfo(sample_type)
        wavfile.write(wave_file_path, self._sample_rate, np.clip(self._data, info.min, info.max).astype(sample_type))

    def get_data(self, channel_index=None):
        if channel_index is None or self._data.ndim == 1:
            return self._data
        return self._data[:, channel_index]

    def get_number_of_samples(self):
        return self._data.shape[0]

    def get_number_of_channels(self):
        return 1 if self._data.ndim == 1 else self._data.shape[1]

    def get_sample_rate(self):
        return self._sample_rate

    def get_sample_type(self):
        return self._data.dtype

    def set_sample_type(self, sample_type):
        self._data = self._data.astype(sample_type)

    def pad_with_zeros(self, new_length):
        if self.get_number_of_samples() > new_length:
            raise Exception("cannot pad for shorter signal length")
        pad = [(0, new_length - self.get_number_of_samples())] + [(0, 0)] * (self._data.ndim - 1)
        self._data = np.pad(self._data, pad, mode="constant")

    def truncate(self, new_length):
        if self.get_number_of_samples() < new_length:
            raise Exception("cannot truncate for longer signal length")
        self._data = self._data[:new_length]

    def amplify_by_factor(self, factor):
        self._data = self._data.astype(np.float64) * factor

    @staticmethod
    def concat(signals):
        return AudioSignal(np.concatenate([s.get_data() for s in signals]), signals[0].get_sample_rate())


class AudioMixer(object):

    @staticmethod
    def snr_factor(signal, noise, snr_db):
        s = signal.get_data().astype(np.float64)
        n = noise.get_data().astype(np.float64)
        if s.size != n.size:
            raise Exception("signal and noise must have the same length")
        return float(np.sqrt(np.var(s) / np.var(n)) * (10.0 ** (-snr_db / 20.0)))

    @staticmethod
    def mix(audio_signals, mixing_weights=None):
        if mixing_weights is None:
            mixing_weights = [1.0 / len(audio_signals)] * len(audio_signals)
        mixed = np.zeros(audio_signals[0].get_data().shape, dtype=np.float64)
        for sig, w in zip(audio_signals, mixing_weights):
            mixed += float(w) * sig.get_data().astype(np.float64)
        return AudioSignal(mixed, audio_signals[0].get_sample_rate())
