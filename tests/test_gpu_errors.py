"""Error behaviour at the boundary: the C ABI returns a negative status for bad arguments (never throws, never
launches), the Python binding turns it into an exception with the library's message, and the batch driver keeps the
reference's "bad sample is skipped, the batch continues" semantics (dp:180-186, dp:196)."""
import ctypes
import importlib

import numpy as np
import pytest
import torch

from tests.cases import SR, FPS, SLICE_MS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mod():
    return importlib.import_module("audio-visual-speech-enhancement_b200.engine")


@pytest.fixture(scope="module")
def native():
    return importlib.import_module("audio-visual-speech-enhancement_b200._native")


@pytest.fixture(scope="module")
def eng(mod):
    return mod.SpectralEngine(SR, FPS, SLICE_MS, device="cuda:0")


def test_c_abi_status_codes(eng, native):
    lib = eng._lib
    a = native.ForwardArgs()
    assert lib.avse_forward(eng._ctx, ctypes.byref(a), None) == -1          # AVSE_E_ARG: NULL speech / max_key
    assert b"required" in lib.avse_last_error()
    assert lib.avse_forward(None, ctypes.byref(a), None) == -1
    assert lib.avse_snr_factor(eng._ctx, None, None, 0, 0, None, None, 1, 10, None, None, None, None, None, None) == -1
    h = ctypes.c_void_p()
    assert lib.avse_create(16000, 0.0, 8000.0, 9999, ctypes.byref(h)) == -1 and not h.value   # bad device index
    assert lib.avse_create(16000, 9000.0, 8000.0, 0, ctypes.byref(h)) == -2 and not h.value   # AVSE_E_CONFIG: fmin > fmax
    per = ctypes.c_longlong(0)
    assert lib.avse_inverse_work_elems(0, ctypes.byref(per)) == -1


def test_python_layer_raises_with_the_library_message(eng, native):
    s = torch.zeros((2, 300), device="cuda")            # L <= 320: reflect padding impossible (librosa would raise too)
    with pytest.raises(native.AvseError, match="L > 320"):
        eng.forward_raw(s, s.clone())
    s = torch.zeros((1, 16000), device="cuda")
    with pytest.raises(native.AvseError, match="n_slices"):
        eng.forward_raw(s, s.clone(), n_slices=6)       # more slices than int(T / 20) (dp:50)
    with pytest.raises(native.AvseError, match="int16"):
        _single_int16(eng, native)                      # int16 samples without a noise signal: refused, not mis-read
    # the engine is still usable after failed calls
    ok = eng.preprocess_pairs(torch.randn((1, 16000), device="cuda"), torch.randn((1, 16000), device="cuda"), 5)
    assert torch.isfinite(ok[0]).all()


def _single_int16(eng, native):
    a = native.ForwardArgs()
    x = torch.zeros((1, 16000), dtype=torch.int16, device="cuda")
    out = torch.empty((1, 5, 80, 20), device="cuda")
    key = torch.zeros((1, 3), dtype=torch.int32, device="cuda")
    a.speech, a.in_stride, a.B, a.L, a.n_slices = x.data_ptr(), 16000, 1, 16000, 5
    a.out_speech, a.out_stride, a.max_key, a.sample_format = out.data_ptr(), 8000, key.data_ptr(), 1
    native.check(eng._lib.avse_forward(eng._ctx, ctypes.byref(a), None), "avse_forward")


def test_unsupported_geometry_is_refused(mod, native):
    # every n_fft the reference derives is served (generic kernels); only sizes beyond the tables' limits are refused
    eng30 = mod.SpectralEngine(SR, 30.0, SLICE_MS, device="cuda:0")     # n_fft 533 (odd), hop 133, 24 frames per slice
    assert (eng30.n_fft, eng30.hop, eng30.spss, eng30.n_bins) == (533, 133, 24, 267) and not eng30.specialised
    with pytest.raises(native.AvseError, match="n_fft"):
        mod.SpectralEngine(48000, 10.0, SLICE_MS, device="cuda:0")      # n_fft 4800 > 4096
    with pytest.raises(ValueError):
        mod.SpectralEngine(SR, 9000.0, SLICE_MS, device="cuda:0")       # n_fft 1, hop 0


def test_batch_driver_skips_failed_samples(eng, tmp_path):
    dp = importlib.import_module("audio-visual-speech-enhancement_b200.data_processor")
    from scipy.io import wavfile
    from collections import namedtuple
    Entry = namedtuple("Entry", ["speaker_id", "audio_path", "video_path"])
    rng = np.random.RandomState(0)
    entries, noises = [], []
    for i in range(3):
        sp, nz = tmp_path / ("s%d.wav" % i), tmp_path / ("n%d.wav" % i)
        wavfile.write(str(sp), SR, (rng.randn(16000) * 3000).astype(np.int16))
        wavfile.write(str(nz), SR, (rng.randn(9000) * 1000).astype(np.int16))
        entries.append(Entry("spk", str(sp), "v%d.mp4" % i))
        noises.append(str(nz))
    entries[1] = Entry("spk", str(tmp_path / "missing.wav"), "v1.mp4")   # unreadable audio -> sample dropped, others kept

    def video(path, slice_ms):
        return np.zeros((4 if path == "v2.mp4" else 5, 128, 128, 5), np.uint8), FPS

    samples = dp.preprocess_data(entries, noises, video)
    assert [s.video_file_path for s in samples] == ["v0.mp4", "v2.mp4"]
    assert samples[0].mixed_spectrograms.shape == (5, 80, 20)
    assert samples[1].mixed_spectrograms.shape == (4, 80, 20)             # n_slices = min(video, audio), dp:164
    assert samples[0].mixed_signal.get_number_of_samples() == 16000


def test_kernels_stay_inside_their_buffers(eng):
    # compute-sanitizer is not available on this pool: guard bands instead.  Every output row is followed by a gap filled
    # with a sentinel; after the forward, floor and inverse kernels the gaps (and an unused tail utterance) are untouched.
    B, L, n = 5, 16000, 5
    SENT = -12345.0
    g = torch.Generator(device="cuda").manual_seed(11)
    s = torch.randn((B, L), generator=g, device="cuda") * 0.1
    z = torch.randn((B, L), generator=g, device="cuda") * 0.05
    lens = torch.tensor([16000, 15999, 9000, 16000, 321 + 7], dtype=torch.int32, device="cuda")
    gap = 64
    bufs = {k: torch.full((B + 1, n * 1600 + gap), SENT, device="cuda") for k in ("speech", "noise", "mixed")}
    pcm = torch.full((B + 1, L + gap), SENT, device="cuda")
    out = {k: v[:B, :n * 1600].view(B, n, 80, 20) for k, v in bufs.items()}
    out["mixed_pcm"] = pcm[:B, :L]
    f, keys = eng.snr_factor(s, z, lengths=lens)
    r = eng.forward_raw(s, z, L=L, len_speech=lens, len_noise=lens, factor=f, n_slices=n, max_key=keys, out=out)
    eng.floor3_(r["speech"], r["noise"], r["mixed"], keys)
    assert r["mixed"].data_ptr() == bufs["mixed"].data_ptr()
    for k, v in bufs.items():
        assert torch.all(v[:B, n * 1600:] == SENT) and torch.all(v[B] == SENT), k
        assert torch.isfinite(v[:B, :n * 1600]).all() and not torch.any(v[:B, :n * 1600] == SENT), k
    assert torch.all(pcm[:B, L:] == SENT) and torch.all(pcm[B] == SENT) and not torch.any(pcm[:B, :L] == SENT)
    T_use = 100
    rec_buf = torch.full((B + 1, 160 * (T_use - 1) + gap), SENT, device="cuda")
    rec = eng.reconstruct(pcm[:B, :L], r["mixed"], lengths=lens, out=rec_buf[:B, :160 * (T_use - 1)])
    assert rec.data_ptr() == rec_buf.data_ptr()
    assert torch.all(rec_buf[:B, 160 * (T_use - 1):] == SENT) and torch.all(rec_buf[B] == SENT)
    assert torch.isfinite(rec).all() and not torch.any(rec == SENT)
    rec16_buf = torch.full((B + 1, 160 * (T_use - 1) + gap), 12321, dtype=torch.int16, device="cuda")
    eng.reconstruct(pcm[:B, :L], r["mixed"], lengths=lens, out=rec16_buf[:B, :160 * (T_use - 1)])
    assert torch.all(rec16_buf[:B, 160 * (T_use - 1):] == 12321) and torch.all(rec16_buf[B] == 12321)
