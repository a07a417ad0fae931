"""Drop-in mirror of the audio half of /root/reference/data_processor.py.

Same function names, positional arguments, return shapes and in-place mutation semantics as the
reference (cited as dp:LINE), but every spectrogram / reconstruction is computed by the CUDA
library through `engine.SpectralEngine`.  Batched entry points (`preprocess_audio_pairs`,
`preprocess_data`) replace the reference's `multiprocess.Pool(16)` fan-out (dp:189-198): CUDA
work is batched in the parent process, one launch over many utterances.

Module surface == the reference module's: preprocess_video_sample, preprocess_audio_signal,
reconstruct_speech_signal, signal_to_spectrogram, reconstruct_signal_from_spectrogram, preprocess_audio_pair, Sample,
preprocess_sample, try_preprocess_sample, preprocess_data, VideoNormalizer -- so that the reference's
`speech_enhancer.py` runs unmodified with `sys.modules["data_processor"]` pointing here
(tests/test_dropin_speech_enhancer.py).  The video front end (dp:12-32: video decode + mouth crop through the third-party
`facedetection` / `mediaio.video_io`) is out of scope as arithmetic: `preprocess_video_sample` calls those packages when
they are importable, and `preprocess_data` accepts any other callable with the same contract.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch

from .engine import SpectralEngine, geometry, fit_noise, N_FFT, HOP, N_BINS, N_MELS, SPSS
from .mediaio_compat import AudioSignal, AudioMixer

_engines = {}


def _device_index(device):
    """GPU index of `device`; None or a bare "cuda" mean the current device."""
    if device is None:
        return torch.cuda.current_device()
    idx = torch.device(device).index
    return torch.cuda.current_device() if idx is None else idx


def get_engine(sample_rate=16000, video_frame_rate=25.0, slice_duration_ms=200, device=None):
    """One cached SpectralEngine per (sample_rate, fps, slice_ms, device)."""
    dev = _device_index(device)
    key = (int(sample_rate), float(video_frame_rate), slice_duration_ms, dev)
    if key not in _engines:
        _engines[key] = SpectralEngine(sample_rate, video_frame_rate, slice_duration_ms, device="cuda:%d" % dev)
    return _engines[key]


def _engine_for_fft(sample_rate, n_fft, hop_length, device=None):
    """Engine for an explicit (n_fft, hop_length) pair (the arguments of dp:77 / dp:99)."""
    dev = _device_index(device)
    key = (int(sample_rate), "fft", int(n_fft), int(hop_length), dev)
    if key not in _engines:
        _engines[key] = SpectralEngine(sample_rate, float(sample_rate) / n_fft, 200, device="cuda:%d" % dev, n_fft=n_fft, hop=hop_length)
    return _engines[key]


def _mono_f32(audio_signal):
    data = audio_signal.get_data(channel_index=0)  # dp:78
    return np.ascontiguousarray(data, dtype=np.float32)


def _mono(audio_signal, keep_int16):
    """Channel 0 (dp:78) as float32, or as the raw int16 WAV samples (from_wav_file, dp:122-123) when the kernels
    decode them themselves (halves the host->device bytes)."""
    data = audio_signal.get_data(channel_index=0)
    return np.ascontiguousarray(data, dtype=np.int16 if keep_int16 else np.float32)


def _to_dev(x, eng):
    return torch.from_numpy(x).to(eng.device, non_blocking=True)


# --------------------------------------------------------------------------------------------
# dp:77-96
# --------------------------------------------------------------------------------------------
def signal_to_spectrogram(audio_signal, n_fft, hop_length, mel=True, db=True):
    """dp:77-96.  Returns (magnitude (80|321, T) float32, phase (321, T) complex64)."""
    eng = _engine_for_fft(audio_signal.get_sample_rate(), n_fft, hop_length)
    x = _to_dev(_mono_f32(audio_signal), eng)
    db_mel, D = eng.spectrogram(x, stft=True)
    mag, phase = eng.magphase(D)          # librosa.magphase: 1+0j where D == 0 (dp:80); (321, T) orientation
    mag, phase = mag[0], phase[0]
    if mel and db:
        magnitude = db_mel[0]
    else:
        # non-default flag combinations are not on the reference's call paths (dp:47, dp:64 always
        # pass mel=True, db=True); they are served from the complex STFT with torch glue.
        magnitude = mag
        if mel:
            magnitude = torch.from_numpy(eng.filterbank()).to(mag.device, torch.float32) @ magnitude
        if db:
            magnitude = 20.0 * torch.log10(torch.clamp(magnitude, min=1e-5))
            magnitude = torch.maximum(magnitude, magnitude.max() - 80.0)
    return magnitude.cpu().numpy(), phase.cpu().numpy()


# --------------------------------------------------------------------------------------------
# dp:35-57
# --------------------------------------------------------------------------------------------
def preprocess_audio_signal(audio_signal, slice_duration_ms, n_video_slices, video_frame_rate):
    """dp:35-57.  Mutates audio_signal (pad/truncate, dp:39-42) and returns (n, 80, 20) float32."""
    samples_per_slice = int((float(slice_duration_ms) / 1000) * audio_signal.get_sample_rate())
    signal_length = samples_per_slice * n_video_slices
    if audio_signal.get_number_of_samples() < signal_length:
        audio_signal.pad_with_zeros(signal_length)
    else:
        audio_signal.truncate(signal_length)
    eng = get_engine(audio_signal.get_sample_rate(), video_frame_rate, slice_duration_ms)
    x = _to_dev(_mono_f32(audio_signal), eng)
    return eng.preprocess_signals(x, n_video_slices)[0].cpu().numpy()


# --------------------------------------------------------------------------------------------
# dp:60-74, dp:99-116
# --------------------------------------------------------------------------------------------
def reconstruct_speech_signal(mixed_signal, speech_spectrograms, video_frame_rate):
    """dp:60-74.  speech_spectrograms: (n, 80, 20) dB slices.  Returns AudioSignal (float32 data)."""
    eng = get_engine(mixed_signal.get_sample_rate(), video_frame_rate)
    spec = np.asarray(speech_spectrograms, dtype=np.float32)
    if spec.ndim != 3 or spec.shape[1:] != (eng.n_mels, eng.spss):
        raise ValueError("speech_spectrograms must have shape (n_slices, 80, 20)")
    pcm = _to_dev(_mono_f32(mixed_signal), eng)
    out = eng.reconstruct(pcm, _to_dev(np.ascontiguousarray(spec), eng))
    return AudioSignal(out[0].cpu().numpy(), mixed_signal.get_sample_rate())


def reconstruct_signal_from_spectrogram(magnitude, phase, sample_rate, n_fft, hop_length, mel=True, db=True):
    """dp:99-116 with an explicit phase array (321, T).  Returns AudioSignal."""
    if not (mel and db):
        raise NotImplementedError("only the reference's call form: mel=True, db=True (dp:72-74)")
    eng = _engine_for_fft(sample_rate, n_fft, hop_length)
    mag = np.ascontiguousarray(np.asarray(magnitude, dtype=np.float32))
    ph = np.ascontiguousarray(np.asarray(phase).astype(np.complex64).T)  # frame-major (T, 321)
    out = eng.reconstruct_with_phase(_to_dev(mag, eng).unsqueeze(0), _to_dev(ph, eng).unsqueeze(0))
    return AudioSignal(out[0].cpu().numpy(), sample_rate)


# --------------------------------------------------------------------------------------------
# dp:119-139
# --------------------------------------------------------------------------------------------
def _fit_noise_np(noise, n_speech):
    """dp:125-128 as an explicit host copy: double until long enough, truncate == periodic tiling.  The product path
    does not use it (the kernels address noise[i mod Ln] themselves); kept for callers and tests."""
    n = noise.shape[0]
    if n < n_speech:
        noise = noise[np.arange(n_speech) % n]
    return noise[:n_speech]


def preprocess_audio_pair_signals(speech_signal, noise_signal, slice_duration_ms, n_video_slices, video_frame_rate, snr_db=0):
    """dp:125-139 on in-memory AudioSignal objects.  Mutates speech_signal (pad/truncate, dp:136) like the
    reference; returns (mixed_slices, speech_slices, noise_slices, mixed_signal)."""
    out = preprocess_audio_pairs([speech_signal], [noise_signal], slice_duration_ms, [n_video_slices], video_frame_rate,
                                 snr_db=[snr_db])
    if isinstance(out[0], Exception):
        raise out[0]
    return out[0]


def preprocess_audio_pair(speech_file_path, noise_file_path, slice_duration_ms, n_video_slices, video_frame_rate):
    """dp:119-139 (same positional signature)."""
    print("preprocessing pair: %s, %s" % (speech_file_path, noise_file_path))
    speech_signal = AudioSignal.from_wav_file(speech_file_path)
    noise_signal = AudioSignal.from_wav_file(noise_file_path)
    return preprocess_audio_pair_signals(speech_signal, noise_signal, slice_duration_ms, n_video_slices, video_frame_rate)


def _pair_bucket(eng, sr, speech_signals, noise_signals, idxs, nvs, snr_db):
    """One launch over the utterances `idxs` (equal n_video_slices).  Returns {i: 4-tuple or Exception}."""
    L = eng.samples_per_slice * nvs
    i16 = all(speech_signals[i].get_data().dtype == np.int16 and noise_signals[i].get_data().dtype == np.int16 for i in idxs)
    sp = [_mono(speech_signals[i], i16) for i in idxs]
    nz = [_mono(noise_signals[i], i16) for i in idxs]
    for r, i in enumerate(idxs):
        if len(nz[r]) == 0 and len(sp[r]) > 0:
            raise ValueError("empty noise signal")          # the reference's doubling loop (dp:125-126) would never end
    width = max(max(len(s) for s in sp), 1)
    dt = np.int16 if i16 else np.float32
    S = np.zeros((len(idxs), width), dtype=dt)
    N = np.zeros((len(idxs), width), dtype=dt)
    lens = np.zeros(len(idxs), dtype=np.int32)
    nlens = np.zeros(len(idxs), dtype=np.int32)
    for r in range(len(idxs)):
        m = min(len(nz[r]), len(sp[r]))                     # dp:128 truncate; a shorter noise is tiled in-kernel (dp:125-126)
        S[r, :len(sp[r])] = sp[r]
        N[r, :m] = nz[r][:m]
        lens[r] = len(sp[r])
        nlens[r] = len(nz[r]) if len(nz[r]) < len(sp[r]) else len(sp[r])
    snr = None
    if snr_db is not None:
        snr = _to_dev(np.asarray([snr_db[i] for i in idxs], dtype=np.float32), eng)
    info = {}
    mixed, speech, noise, pcm = eng.preprocess_pairs(_to_dev(S, eng), _to_dev(N, eng), nvs, lengths=_to_dev(lens, eng), snr_db=snr,
                                                     noise_lengths=_to_dev(nlens, eng), info=info)
    factor = info["factor"].cpu().numpy()
    mixed, speech, noise, pcm = mixed.cpu().numpy(), speech.cpu().numpy(), noise.cpu().numpy(), pcm.cpu().numpy()
    res = {}
    for r, i in enumerate(idxs):
        if not np.isfinite(factor[r]):
            # silent noise file: numpy gives an inf / nan factor and librosa.stft refuses the non-finite mixture, so the
            # reference drops this sample in try_preprocess_sample (dp:180-186)
            res[i] = ValueError("SNR factor is not finite (noise variance is zero)")
            continue
        # observable mutation of the reference: speech padded/truncated in place (dp:136 -> dp:39-42)
        if speech_signals[i].get_number_of_samples() < L:
            speech_signals[i].pad_with_zeros(L)
        else:
            speech_signals[i].truncate(L)
        res[i] = (mixed[r], speech[r], noise[r], AudioSignal(pcm[r], sr))
    return res


def preprocess_audio_pairs(speech_signals, noise_signals, slice_duration_ms, n_video_slices, video_frame_rate, snr_db=None):
    """Batched dp:119-139: lists of AudioSignal -> list of (mixed, speech, noise slices, mixed AudioSignal).

    Utterances are bucketed by n_video_slices (equal signal_length) and each bucket is one launch.  Failure isolation is
    per SAMPLE like the reference's try_preprocess_sample (dp:180-186): if a bucket fails as a whole it is retried
    utterance by utterance, and the entry of a failing utterance is its Exception instead of a tuple."""
    sr = speech_signals[0].get_sample_rate()
    eng = get_engine(sr, video_frame_rate, slice_duration_ms)
    results = [None] * len(speech_signals)
    buckets = {}
    for i, nvs in enumerate(n_video_slices):
        buckets.setdefault(int(nvs), []).append(i)
    for nvs, idxs in buckets.items():
        try:
            done = _pair_bucket(eng, sr, speech_signals, noise_signals, idxs, nvs, snr_db)
        except Exception as batch_error:
            done = {}
            for i in idxs:
                if len(idxs) == 1:
                    done[i] = batch_error
                    break
                try:
                    done.update(_pair_bucket(eng, sr, speech_signals, noise_signals, [i], nvs, snr_db))
                except Exception as e:
                    done[i] = e
        for i in idxs:
            results[i] = done[i]
    return results


# --------------------------------------------------------------------------------------------
# dp:142-198 sample assembly and the batch driver
# --------------------------------------------------------------------------------------------
Sample = namedtuple('Sample', [
    'speaker_id',
    'video_file_path',
    'speech_file_path',
    'noise_file_path',
    'video_samples',
    'mixed_spectrograms',
    'speech_spectrograms',
    'noise_spectrograms',
    'mixed_signal',
    'video_frame_rate'
])


def assemble_sample(speech_entry, noise_file_path, video_samples, video_frame_rate, pair_result):
    """dp:164-177: n_slices = min(video, audio); truncate the four arrays."""
    mixed_spectrograms, speech_spectrograms, noise_spectrograms, mixed_signal = pair_result
    n_slices = min(video_samples.shape[0], mixed_spectrograms.shape[0])
    return Sample(
        speaker_id=speech_entry.speaker_id,
        video_file_path=speech_entry.video_path,
        speech_file_path=speech_entry.audio_path,
        noise_file_path=noise_file_path,
        video_samples=video_samples[:n_slices],
        mixed_spectrograms=mixed_spectrograms[:n_slices],
        speech_spectrograms=speech_spectrograms[:n_slices],
        noise_spectrograms=noise_spectrograms[:n_slices],
        mixed_signal=mixed_signal,
        video_frame_rate=video_frame_rate
    )


def preprocess_video_sample(video_file_path, slice_duration_ms, mouth_height=128, mouth_width=128):
    """dp:12-32 contract: (slices (n, height, width, frames_per_slice) float32, frame rate).  The decode and the mouth crop
    belong to the third-party `mediaio.video_io` / `facedetection` packages (dp:7, dp:9; out of scope);
    they are imported here, on first use, so that this module loads without them."""
    from facedetection.face_detection import FaceDetector
    from mediaio.video_io import VideoFileReader
    print("preprocessing %s" % video_file_path)
    detector = FaceDetector()
    with VideoFileReader(video_file_path) as reader:
        frames = reader.read_all_frames(convert_to_gray_scale=True)
        n_frames, fps = reader.get_frame_count(), reader.get_frame_rate()
    crops = np.empty((mouth_height, mouth_width, n_frames), dtype=np.float32)
    for t in range(n_frames):
        crops[:, :, t] = detector.crop_mouth(frames[t], bounding_box_shape=(mouth_width, mouth_height))
    per_slice = int((float(slice_duration_ms) / 1000) * fps)           # dp:24
    n_slices = int(float(n_frames) / per_slice)                        # dp:25
    slices = crops[:, :, :n_slices * per_slice].reshape(mouth_height, mouth_width, n_slices, per_slice)
    return np.ascontiguousarray(np.moveaxis(slices, 2, 0)), fps


def preprocess_sample(speech_entry, noise_file_path, slice_duration_ms=200):
    """dp:156-177 for one sample (same signature)."""
    print("preprocessing sample: %s, %s, %s..." % (speech_entry.video_path, speech_entry.audio_path, noise_file_path))
    video_samples, video_frame_rate = preprocess_video_sample(speech_entry.video_path, slice_duration_ms)
    pair = preprocess_audio_pair(speech_entry.audio_path, noise_file_path, slice_duration_ms, video_samples.shape[0], video_frame_rate)
    return assemble_sample(speech_entry, noise_file_path, video_samples, video_frame_rate, pair)


def try_preprocess_sample(sample_paths):
    """dp:180-186."""
    try:
        return preprocess_sample(*sample_paths)
    except Exception as e:
        print("failed to preprocess %s (%s)" % (sample_paths, e))
        return None


def preprocess_data(speech_entries, noise_file_paths, video_preprocessor=None, slice_duration_ms=200):
    """dp:189-198 (callable exactly like the reference: `preprocess_data(speech_entries, noise_file_paths)`, se:25) with the
    Pool(16) fan-out replaced by one batched GPU pass per (frame rate, slice count) bucket.

    video_preprocessor(video_path, slice_duration_ms) -> (video_samples, fps): defaults to this module's
    preprocess_video_sample (dp:12-32).  A sample whose video, audio or arithmetic fails is dropped and the others are
    kept, like try_preprocess_sample (dp:180-186); the order of the surviving samples is the input order (Pool.map)."""
    print("preprocessing data...")
    if video_preprocessor is None:
        video_preprocessor = preprocess_video_sample
    staged = []
    for entry, noise_path in zip(speech_entries, noise_file_paths):
        try:
            video_samples, fps = video_preprocessor(entry.video_path, slice_duration_ms)
            speech = AudioSignal.from_wav_file(entry.audio_path)
            noise = AudioSignal.from_wav_file(noise_path)
            staged.append((entry, noise_path, video_samples, fps, speech, noise))
        except Exception as e:  # dp:184-186
            print("failed to preprocess %s (%s)" % ((entry, noise_path), e))
    by_key = {}
    for pos, item in enumerate(staged):
        by_key.setdefault((float(item[3]), item[4].get_sample_rate()), []).append(pos)
    done = [None] * len(staged)
    for (fps, _sr), positions in by_key.items():
        items = [staged[p] for p in positions]
        try:
            res = preprocess_audio_pairs([it[4] for it in items], [it[5] for it in items], slice_duration_ms,
                                         [it[2].shape[0] for it in items], fps)
        except Exception as e:
            res = [e] * len(items)
        for p, it, r in zip(positions, items, res):
            if isinstance(r, Exception):
                print("failed to preprocess %s (%s)" % ((it[0], it[1]), r))
            else:
                done[p] = assemble_sample(it[0], it[1], it[2], it[3], r)
    return [d for d in done if d is not None]


class VideoNormalizer(object):
    """dp:201-212 with the reference's constructor and in-place `normalize`: per-pixel mean / population std over
    (slices, frames) and (x - mean) / std, computed by avse_video_stats / avse_video_normalize on the GPU.  The state is
    two numpy images, so the object pickles like the reference's (se:55-56, se:66-67)."""

    def __init__(self, video_samples):
        # video_samples: slices x height x width x frames_per_slice
        from .engine import VideoNormalizer as _DeviceNormalizer
        dev = _DeviceNormalizer(get_engine(), video_samples)
        self.__mean_image = dev.mean_image.cpu().numpy()
        self.__std_image = dev.std_image.cpu().numpy()

    def normalize(self, video_samples):
        from .engine import VideoNormalizer as _DeviceNormalizer
        dev = _DeviceNormalizer.from_images(get_engine(), self.__mean_image, self.__std_image)
        dev.normalize(video_samples)


def make_sample_set(samples, max_samples=None, permutation=None):
    """speech_enhancer.py:241-262 (next row f1), same signature: a random subset of `max_samples` samples, the slices of
    all of them concatenated, ONE shared permutation.  The two spectrogram arrays are concatenated and permuted on the
    GPU (avse_gather_rows); the video slices are only indexed (they never enter the spectral path).
    permutation: optional explicit permutation of the concatenated rows (tests); None: np.random.permutation like se:255."""
    import random as _random
    n_samples = len(samples) if max_samples is None else min(len(samples), max_samples)
    samples = _random.sample(list(samples), n_samples)
    video = np.concatenate([s.video_samples for s in samples], axis=0)
    perm = np.random.permutation(video.shape[0]) if permutation is None else np.asarray(permutation)
    eng = get_engine()
    mixed = _to_dev(np.ascontiguousarray(np.concatenate([s.mixed_spectrograms for s in samples], axis=0), dtype=np.float32), eng)
    speech = _to_dev(np.ascontiguousarray(np.concatenate([s.speech_spectrograms for s in samples], axis=0), dtype=np.float32), eng)
    m, sp, _ = eng.make_sample_set(mixed.unsqueeze(0), speech.unsqueeze(0), permutation=_to_dev(perm.astype(np.int64), eng))
    return video[perm], m.cpu().numpy(), sp.cpu().numpy()


class MelConverter(object):
    """Facade named by BASELINE.json's north_star (the reference snapshot has free functions instead,
    SURVEY.md section 0 item 2).  Signatures are this build's design."""

    def __init__(self, sample_rate=16000, video_frame_rate=25.0, slice_duration_ms=200):
        self.engine = get_engine(sample_rate, video_frame_rate, slice_duration_ms)
        self.sample_rate = sample_rate
        self.video_frame_rate = video_frame_rate
        self.slice_duration_ms = slice_duration_ms

    def signal_to_mel_spectrogram(self, audio_signal):
        return signal_to_spectrogram(audio_signal, self.engine.n_fft, self.engine.hop, mel=True, db=True)[0]

    def signal_to_slices(self, audio_signal, n_video_slices):
        return preprocess_audio_signal(audio_signal, self.slice_duration_ms, n_video_slices, self.video_frame_rate)

    def mix_pair(self, speech_signal, noise_signal, n_video_slices, snr_db=0):
        return preprocess_audio_pair_signals(speech_signal, noise_signal, self.slice_duration_ms, n_video_slices,
                                             self.video_frame_rate, snr_db=snr_db)

    def reconstruct_signal_from_mel_spectrogram(self, mixed_signal, mel_slices):
        return reconstruct_speech_signal(mixed_signal, mel_slices, self.video_frame_rate)
