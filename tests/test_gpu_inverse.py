"""GPU parity tests of the inverse path (avse_inverse through the C ABI) vs the float64 oracle and golden vectors."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import avse_oracle as O
from tests.cases import GOLDEN_CASES, SR, FPS, oracle_pair

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_PCM = 1e-4   # of full scale (BASELINE.json north_star)


@pytest.fixture(scope="module")
def eng():
    mod = importlib.import_module("audio-visual-speech-enhancement_b200.engine")
    return mod.SpectralEngine(SR, FPS, 200, device="cuda:0")


@pytest.fixture(scope="module")
def dp():
    return importlib.import_module("audio-visual-speech-enhancement_b200.data_processor")


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c["name"] for c in GOLDEN_CASES])
def test_reconstruct_matches_oracle_and_golden(eng, case):
    gold = np.load(os.path.join(GOLD, case["name"] + ".npz"))
    out = eng.reconstruct(_dev(gold["mixed_pcm"][None]), _dev(gold["speech"][None]))[0].cpu().numpy()
    want = O.reconstruct_speech_signal(O.AudioSignal(gold["mixed_pcm"].astype(np.float64), SR), gold["speech"].astype(np.float64), FPS).get_data()
    scale = np.max(np.abs(gold["mixed_pcm"]))
    assert out.shape == want.shape
    assert np.max(np.abs(out - want)) <= TOL_PCM * scale
    assert np.max(np.abs(out - gold["recon"])) <= TOL_PCM * scale


def test_batch_every_utterance_matches_its_own_oracle(eng):
    golds = [np.load(os.path.join(GOLD, c["name"] + ".npz")) for c in GOLDEN_CASES[:3]]
    pcm = np.stack([g["mixed_pcm"] for g in golds])
    mel = np.stack([g["speech"] for g in golds])
    out = eng.reconstruct(_dev(pcm), _dev(mel)).cpu().numpy()
    for i, g in enumerate(golds):
        assert np.max(np.abs(out[i] - g["recon"])) <= TOL_PCM * np.max(np.abs(g["mixed_pcm"]))


def test_forward_then_inverse_round_trip_full_size(eng):
    # BASELINE config 4 scale (1000 x 3 s): the chain forward -> inverse must equal, utterance by utterance, the same
    # chain run on a 4-utterance sub-batch (batch-size / chunking independence) and stay finite everywhere.
    B, L = 1000, 48000
    g = torch.Generator(device="cuda").manual_seed(7)
    s = torch.randn((B, L), generator=g, device="cuda") * 0.1
    n = torch.randn((B, L), generator=g, device="cuda") * 0.05
    mixed, speech, noise, pcm = eng.preprocess_pairs(s, n, 15)
    rec = eng.reconstruct(pcm, mixed)
    assert rec.shape == (B, 47840) and torch.isfinite(rec).all()
    idx = torch.tensor([0, 1, 500, 999], device="cuda")
    m2, s2, n2, p2 = eng.preprocess_pairs(s[idx].contiguous(), n[idx].contiguous(), 15)
    rec2 = eng.reconstruct(p2, m2)
    assert torch.equal(m2, mixed[idx]) and torch.equal(rec2, rec[idx])
    # reconstructing from the mixture's own log-mel + its own phase must resemble the mixture (lossy mel round trip)
    err = (rec - pcm[:, :47840]).pow(2).mean().sqrt() / pcm.pow(2).mean().sqrt()
    assert err.item() < 0.6


def test_explicit_phase_and_data_processor_signatures(eng, dp):
    gold = np.load(os.path.join(GOLD, GOLDEN_CASES[0]["name"] + ".npz"))
    pcm = gold["mixed_pcm"]
    sig = dp.AudioSignal(pcm.copy(), SR)
    rec = dp.reconstruct_speech_signal(sig, gold["speech"], FPS)
    assert rec.get_number_of_samples() == 15840 and rec.get_sample_rate() == SR
    assert np.max(np.abs(rec.get_data() - gold["recon"])) <= TOL_PCM * np.max(np.abs(pcm))
    # dp:99-116 form: explicit magnitude (80, T) + phase (321, T)
    osig = O.AudioSignal(pcm.astype(np.float64), SR)
    mag, phase = O.signal_to_spectrogram(osig, 640, 160)
    want = O.reconstruct_signal_from_spectrogram(mag, phase, SR, 640, 160).get_data()
    got = dp.reconstruct_signal_from_spectrogram(mag, phase, SR, 640, 160).get_data()
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= TOL_PCM * np.max(np.abs(pcm))


def test_inverse_is_independent_of_the_work_partition(eng):
    """The I8 kernel cuts the launch's (utterance, group) sequence into one contiguous range per warp; a range that starts inside
    an utterance rebuilds the overlap-add carry by recomputing the previous group.  Results must therefore be bit-identical
    whatever the batch size (i.e. wherever the range boundaries fall) -- also for int16 output and ragged mixture lengths."""
    g = torch.Generator(device="cuda").manual_seed(4)
    B, L, n = 37, 48000, 15
    pcm = torch.randn((B, L), generator=g, device="cuda") * 0.1
    mel = torch.rand((B, n, 80, 20), generator=g, device="cuda") * 60.0 - 70.0
    lens = torch.randint(20000, L + 1, (B,), generator=g, device="cuda", dtype=torch.int32)
    lens[0] = L
    whole = eng.reconstruct(pcm, mel, lengths=lens)
    whole16 = eng.reconstruct(pcm * 30000.0, mel + 90.0, lengths=lens, out_dtype=torch.int16)
    for lo, hi in ((0, 1), (5, 9), (20, 37), (36, 37)):
        part = eng.reconstruct(pcm[lo:hi].contiguous(), mel[lo:hi].contiguous(), lengths=lens[lo:hi].contiguous())
        assert torch.equal(part, whole[lo:hi]), (lo, hi)
        part16 = eng.reconstruct((pcm[lo:hi] * 30000.0).contiguous(), (mel[lo:hi] + 90.0).contiguous(), lengths=lens[lo:hi].contiguous(),
                                 out_dtype=torch.int16)
        assert torch.equal(part16, whole16[lo:hi]), (lo, hi)


@pytest.mark.parametrize("jump_db,tol", [(60.0, 1e-4), (70.0, 1e-4), (80.0, 2e-4)])
def test_inverse_level_jump_between_packed_frames(eng, jump_db, tol):
    """The GPU counterpart of tests/test_emul_inverse.py::test_inverse_level_jump_between_packed_frames: a hard digital step of
    60 / 70 / 80 dB on a frame boundary between two mixture frames that share one packed FFT (see there for the float32 limit)."""
    L = 16000
    step = 160 * 41 + 320
    for seed in (3, 4):
        rng = np.random.RandomState(seed)
        pcm = (0.3 * rng.randn(L)).astype(np.float32)
        pcm[step:] *= np.float32(10.0 ** (-jump_db / 20.0))
        mel = (rng.rand(5, 80, 20) * 30.0 - 50.0).astype(np.float32)
        want = O.reconstruct_speech_signal(O.AudioSignal(pcm.astype(np.float64), SR), mel.astype(np.float64), 25.0).get_data()
        got = eng.reconstruct(torch.from_numpy(pcm[None]).cuda(), torch.from_numpy(mel[None]).cuda())[0].cpu().numpy()
        assert np.max(np.abs(got - want)) <= tol * np.max(np.abs(want)), (jump_db, seed)
