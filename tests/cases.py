"""Seeded synthetic cases shared by the golden generator, the emulation tests and the GPU parity tests."""
import numpy as np

from oracle import avse_oracle as O

SR, FPS, SLICE_MS = 16000, 25.0, 200

GOLDEN_CASES = [
    dict(name="pair_1s_snr0", n_s=16000, n_n=16000, nvs=5, snr=0.0, seed=101, scale=1.0),
    dict(name="pair_1s_snrm10_shortnoise", n_s=17000, n_n=5000, nvs=5, snr=-10.0, seed=102, scale=1.0),
    dict(name="pair_1s_zeropad_int16scale", n_s=12000, n_n=12000, nvs=5, snr=5.0, seed=103, scale=32767.0),
    dict(name="pair_3s_snr0", n_s=48000, n_n=48000, nvs=15, snr=0.0, seed=104, scale=1.0),
]


def make_inputs(case):
    """float32 speech / noise exactly as the GPU path receives them."""
    s = (O.synth_speech(case["n_s"], SR, case["seed"]) * case["scale"]).astype(np.float32)
    n = (O.synth_noise(case["n_n"], case["seed"]) * case["scale"]).astype(np.float32)
    return s, n


def oracle_pair(case, s=None, n=None):
    """float64 oracle of preprocess_audio_pair (dp:119-139) + reconstruct_speech_signal (dp:60-74) on those inputs."""
    if s is None:
        s, n = make_inputs(case)
    sp = O.AudioSignal(s.astype(np.float64), SR)
    nz = O.AudioSignal(n.astype(np.float64), SR)
    mixed, speech, noise, mixed_sig = O.preprocess_audio_pair_signals(sp, nz, SLICE_MS, case["nvs"], FPS, snr_db=case["snr"])
    recon = O.reconstruct_speech_signal(O.AudioSignal(mixed_sig.get_data().copy(), SR), speech, FPS)
    return dict(mixed=mixed, speech=speech, noise=noise, mixed_pcm=mixed_sig.get_data(), recon=recon.get_data())


def fitted_noise(s, n):
    """dp:125-128 on the host: periodic tiling of the noise up to the speech length."""
    if len(n) < len(s):
        n = n[np.arange(len(s)) % len(n)]
    return np.ascontiguousarray(n[:len(s)])
