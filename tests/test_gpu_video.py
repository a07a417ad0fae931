"""GPU tests of the video-side reductions: VideoNormalizer (dp:201-212) and the MSE that
network.evaluate reports (network.py:214-220), against their numpy restatements."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mod():
    return importlib.import_module("audio-visual-speech-enhancement_b200.engine")


@pytest.fixture(scope="module")
def eng(mod):
    return mod.SpectralEngine(16000, 25.0, 200, device="cuda:0")


@pytest.mark.parametrize("shape", [(37, 128, 128, 5), (1, 16, 24, 5), (300, 32, 32, 3), (9, 5, 7, 5), (11, 6, 6, 1)])
def test_video_normalizer_matches_numpy(mod, eng, shape):
    rng = np.random.RandomState(sum(shape))
    video = (rng.rand(*shape) * 255.0).astype(np.float32)           # grey-scale mouth crops (dp:20-22)
    if shape[0] > 1:
        video[:, 3, 5, :] = 17.0                                    # a constant pixel: std 0 -> the reference divides by zero
    mean = np.mean(video.astype(np.float64), axis=(0, 3))           # dp:204-205
    std = np.std(video.astype(np.float64), axis=(0, 3))
    vn = mod.VideoNormalizer(eng, video)
    assert np.max(np.abs(vn.mean_image.cpu().numpy() - mean)) <= 1e-4
    assert np.max(np.abs(vn.std_image.cpu().numpy() - std)) <= 1e-4
    other = (rng.rand(5, *shape[1:]) * 255.0).astype(np.float32)
    m32, s32 = vn.mean_image.cpu().numpy(), vn.std_image.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        want = (other - m32[None, :, :, None]) / s32[None, :, :, None]     # dp:207-212
    got = other.copy()
    assert vn.normalize(got) is got                                  # numpy array overwritten in place, like the reference
    ok = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), ok)
    assert np.max(np.abs(got[ok] - want[ok])) <= 1e-5 * max(1.0, np.max(np.abs(want[ok])))
    dev = torch.from_numpy(other).cuda()
    vn.normalize(dev)                                                # CUDA tensor normalised in place
    assert np.array_equal(np.nan_to_num(dev.cpu().numpy(), nan=0.0, posinf=1e30, neginf=-1e30),
                          np.nan_to_num(got, nan=0.0, posinf=1e30, neginf=-1e30))


def test_mse_matches_numpy(mod, eng):
    rng = np.random.RandomState(0)
    a = (rng.randn(15, 80, 20) * 20 - 30).astype(np.float32)
    b = (a + rng.randn(15, 80, 20)).astype(np.float32)
    got = float(mod.mse(eng, torch.from_numpy(a), torch.from_numpy(b)))
    want = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    assert abs(got - want) <= 1e-6 * want
    big = torch.randn(3_000_001, device="cuda")
    assert abs(float(mod.mse(eng, big, torch.zeros_like(big))) - float((big.double() ** 2).mean())) < 1e-5
