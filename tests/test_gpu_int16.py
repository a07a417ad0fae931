"""GPU tests of the WAV sample format either side of the path: raw int16 samples decoded inside the
forward / SNR kernels (AudioSignal.from_wav_file, dp:122-123) and the clip + int16 cast of AudioSignal.save_to_wav_file
(se:176-177) fused into the inverse kernel's last store."""
import importlib

import numpy as np
import pytest
import torch

from oracle import avse_oracle as O
from tests.cases import SR, FPS, SLICE_MS, fitted_noise

pytestmark = pytest.mark.gpu

TOL_DB = 1e-3
TOL_PCM = 1e-4


@pytest.fixture(scope="module")
def mod():
    return importlib.import_module("audio-visual-speech-enhancement_b200.engine")


@pytest.fixture(scope="module")
def eng(mod):
    return mod.SpectralEngine(SR, FPS, SLICE_MS, device="cuda:0")


def _wav_pair(seed, n_s, n_n, amp=9000.0):
    s = np.round(O.synth_speech(n_s, SR, seed) * amp / 0.3).clip(-32768, 32767).astype(np.int16)
    z = np.round(O.synth_noise(n_n, seed) * amp).clip(-32768, 32767).astype(np.int16)
    return s, z


def test_int16_forward_equals_float_path_and_oracle(eng):
    lens = [48000, 40001, 16000]
    snrs = [0.0, -5.0, 10.0]
    S = np.zeros((3, 48000), np.int16)
    Z = np.zeros((3, 48000), np.int16)
    for i, n in enumerate(lens):
        s, z = _wav_pair(40 + i, n, [48000, 9000, 16000][i])
        S[i, :n], Z[i, :n] = s, fitted_noise(s, z)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    L = d(np.array(lens, np.int32))
    snr = d(np.array(snrs, np.float32))
    got = eng.preprocess_pairs(d(S), d(Z), 15, lengths=L, snr_db=snr)
    ref = eng.preprocess_pairs(d(S.astype(np.float32)), d(Z.astype(np.float32)), 15, lengths=L, snr_db=snr)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)            # int16 -> float32 is exact: same arithmetic, same bits
    for i, n in enumerate(lens):
        sp = O.AudioSignal(S[i, :n].copy(), SR)          # int16 data, as from_wav_file returns it
        nz = O.AudioSignal(Z[i, :n].copy(), SR)
        r_mixed, r_speech, r_noise, r_sig = O.preprocess_audio_pair_signals(sp, nz, SLICE_MS, 15, FPS, snr_db=snrs[i])
        for name, g, r in (("mixed", got[0], r_mixed), ("speech", got[1], r_speech), ("noise", got[2], r_noise)):
            err = np.max(np.abs(g[i].cpu().numpy() - r))
            assert err <= TOL_DB, (i, name, err)
        rp = r_sig.get_data()
        assert np.max(np.abs(got[3][i].cpu().numpy() - rp)) <= TOL_PCM * np.max(np.abs(rp))


def test_int16_output_is_clip_then_truncate(eng):
    s, z = _wav_pair(77, 48000, 48000, amp=15000.0)      # loud: the reconstruction exceeds the int16 range in places
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    mixed, speech, noise, pcm = eng.preprocess_pairs(d(s[None]), d(z[None]), 15)
    loud = mixed + 12.0                                  # +12 dB: forces clipping
    f32 = eng.reconstruct(pcm, loud)
    i16 = eng.reconstruct(pcm, loud, out_dtype=torch.int16)
    assert i16.dtype == torch.int16 and i16.shape == f32.shape
    want = torch.clamp(f32, -32768.0, 32767.0).to(torch.int32).to(torch.int16)    # same float32 values: exact rule check
    assert torch.equal(i16, want)
    assert int(i16.max()) == 32767 and int(i16.min()) == -32768
    # against the float64 oracle + save_to_wav_file semantics: within 1 LSB + the PCM tolerance
    sig = O.AudioSignal(pcm[0].double().cpu().numpy(), SR)
    ref = O.reconstruct_speech_signal(sig, loud[0].double().cpu().numpy(), FPS).get_data()
    ref16 = np.clip(ref, -32768, 32767).astype(np.int16)
    tol = 1 + TOL_PCM * float(pcm.abs().max())
    assert np.max(np.abs(i16[0].cpu().numpy().astype(np.int32) - ref16.astype(np.int32))) <= tol


def test_host_pipeline_int16(eng, mod):
    B, L = 6, 16000
    rng = np.random.RandomState(0)
    hs = torch.from_numpy((rng.randn(B, L) * 3000).astype(np.int16)).pin_memory()
    hn = torch.from_numpy((rng.randn(B, L) * 1000).astype(np.int16)).pin_memory()
    pipe = mod.HostPipeline(eng, L, 5, chunk=4, n_streams=2, sample_dtype=torch.int16)
    outs = [torch.zeros((B, 5, 80, 20)).pin_memory() for _ in range(3)] + [torch.zeros((B, L)).pin_memory()]
    pipe.begin_after(torch.cuda.current_stream())
    pipe.submit(hs, hn, *outs)
    pipe.synchronize()
    ref = eng.preprocess_pairs(hs.cuda().float(), hn.cuda().float(), 5)
    for a, b in zip(outs, ref):
        assert torch.equal(a, b.cpu())


def test_host_pipeline_tiles_short_noise_in_kernel(eng, mod):
    """HostPipeline.submit(noise_lengths=...): noise files shorter than the utterance travel as they are (the rest of the pinned
    row is never read) and are tiled inside the kernels (dp:125-128)."""
    B, L = 5, 16000
    rng = np.random.RandomState(2)
    hs = torch.from_numpy((rng.randn(B, L) * 0.1).astype(np.float32)).pin_memory()
    n_noise = [16000, 5000, 333, 9999, 640]
    hn = torch.full((B, L), float("nan")).pin_memory()
    full = torch.zeros((B, L))
    for i, n in enumerate(n_noise):
        z = torch.from_numpy((rng.randn(n) * 0.03).astype(np.float32))
        hn[i, :n] = z
        full[i] = z[torch.arange(L) % n]
    nl = torch.tensor(n_noise, dtype=torch.int32).pin_memory()
    pipe = mod.HostPipeline(eng, L, 5, chunk=2, n_streams=2)
    outs = [torch.zeros((B, 5, 80, 20)).pin_memory() for _ in range(3)] + [torch.zeros((B, L)).pin_memory()]
    pipe.begin_after(torch.cuda.current_stream())
    pipe.submit(hs, hn, *outs, noise_lengths=nl)
    pipe.synchronize()
    ref = eng.preprocess_pairs(hs.cuda(), full.cuda(), 5)
    for a, b in zip(outs, ref):
        assert torch.isfinite(a).all() and torch.allclose(a, b.cpu(), rtol=0, atol=2e-5)


def test_int16_single_signal_is_converted_by_the_host_layer(eng):
    x = (np.random.RandomState(1).randn(20000) * 2000).astype(np.int16)
    a = eng.preprocess_signals(torch.from_numpy(x).cuda(), 5)
    b = eng.preprocess_signals(torch.from_numpy(x.astype(np.float32)).cuda(), 5)
    assert torch.equal(a, b)


def test_data_processor_mirror_takes_wav_int16(eng):
    dp = importlib.import_module("audio-visual-speech-enhancement_b200.data_processor")
    s, z = _wav_pair(5, 16000, 7000)
    got = dp.preprocess_audio_pair_signals(dp.AudioSignal(s.copy(), SR), dp.AudioSignal(z.copy(), SR), SLICE_MS, 5, FPS, snr_db=0)
    ref = O.preprocess_audio_pair_signals(O.AudioSignal(s.copy(), SR), O.AudioSignal(z.copy(), SR), SLICE_MS, 5, FPS, snr_db=0)
    for g, r in zip(got[:3], ref[:3]):
        assert np.max(np.abs(g - r)) <= TOL_DB
    assert np.max(np.abs(got[3].get_data() - ref[3].get_data())) <= TOL_PCM * np.max(np.abs(ref[3].get_data()))
