"""GPU test of the device-side make_sample_set (speech_enhancer.py:241-262): np.concatenate over samples + ONE shared
permutation, checked bit-exactly against the numpy restatement of the reference lines (a pure gather: no arithmetic)."""
import importlib

import numpy as np
import pytest
import torch

from tests.cases import SR, FPS, SLICE_MS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    mod = importlib.import_module("audio-visual-speech-enhancement_b200.engine")
    return mod.SpectralEngine(SR, FPS, SLICE_MS, device="cuda:0")


def _reference(mixed, speech, n_slices, perm):
    # se:249-257 on host arrays: concatenate each sample's kept slices, then index with the shared permutation
    m = np.concatenate([mixed[i, :n] for i, n in enumerate(n_slices)], axis=0)
    s = np.concatenate([speech[i, :n] for i, n in enumerate(n_slices)], axis=0)
    return m[perm], s[perm]


@pytest.mark.parametrize("B,n_max,ragged", [(7, 15, True), (64, 15, False), (3, 300, True), (1, 1, False)])
def test_make_sample_set_matches_numpy(eng, B, n_max, ragged):
    rng = np.random.RandomState(B * 1000 + n_max)
    mixed = rng.randn(B, n_max, 80, 20).astype(np.float32)
    speech = rng.randn(B, n_max, 80, 20).astype(np.float32)
    n_slices = rng.randint(1, n_max + 1, size=B) if ragged else np.full(B, n_max)
    N = int(n_slices.sum())
    perm = rng.permutation(N)
    gm, gs, gp = eng.make_sample_set(torch.from_numpy(mixed).cuda(), torch.from_numpy(speech).cuda(),
                                     n_slices=n_slices if ragged else None, permutation=perm)
    rm, rs = _reference(mixed, speech, n_slices, perm)
    assert gm.shape == (N, 80, 20) and gs.shape == (N, 80, 20)
    assert np.array_equal(gm.cpu().numpy(), rm) and np.array_equal(gs.cpu().numpy(), rs)
    assert np.array_equal(gp.cpu().numpy(), perm)


def test_random_permutation_is_shared_and_complete(eng):
    B, n_max = 50, 15
    g = torch.Generator(device="cuda").manual_seed(5)
    mixed = torch.arange(B * n_max, device="cuda", dtype=torch.float32).view(B, n_max, 1, 1).expand(B, n_max, 80, 20).contiguous()
    speech = -mixed
    noise = mixed + 0.5
    gm, gs, perm, gn = eng.make_sample_set(mixed, speech, generator=g, extra=noise)
    assert torch.equal(gm, -gs) and torch.equal(gn, gm + 0.5)          # one permutation for every array
    ids = gm[:, 0, 0].long()
    assert torch.equal(ids, perm)                                       # row i came from concatenated row perm[i]
    assert torch.equal(torch.sort(ids).values, torch.arange(B * n_max, device="cuda"))   # a permutation: nothing lost
    assert not torch.equal(ids, torch.arange(B * n_max, device="cuda"))


def test_bad_permutation_is_rejected(eng):
    mixed = torch.zeros((2, 3, 80, 20), device="cuda")
    with pytest.raises(IndexError):
        eng.make_sample_set(mixed, mixed.clone(), permutation=[0, 1, 2, 3, 4, 6])
    with pytest.raises(IndexError):
        eng.make_sample_set(mixed, mixed.clone(), permutation=[0, 1, 2])
