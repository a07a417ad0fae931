"""an engine built for a GPU that is not the current device must work (every engine method runs under a device guard,
and a bare "cuda" device means the current device, not GPU 0)."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mod():
    return importlib.import_module("audio-visual-speech-enhancement_b200.engine")


def test_bare_cuda_device_is_the_current_device(mod):
    eng = mod.SpectralEngine(16000, 25.0, 200, device="cuda")
    assert eng.device == torch.device("cuda", torch.cuda.current_device())
    with pytest.raises(RuntimeError):
        mod.SpectralEngine(16000, 25.0, 200, device="cpu")          # no CPU path


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_on_a_non_current_device(mod):
    assert torch.cuda.current_device() == 0
    eng0 = mod.SpectralEngine(16000, 25.0, 200, device="cuda:0")
    eng1 = mod.SpectralEngine(16000, 25.0, 200, device="cuda:1")
    g = torch.Generator().manual_seed(3)
    s = torch.randn((3, 16000), generator=g) * 0.1
    z = torch.randn((3, 16000), generator=g) * 0.05
    out0 = eng0.preprocess_pairs(s.to("cuda:0"), z.to("cuda:0"), 5)
    out1 = eng1.preprocess_pairs(s.to("cuda:1"), z.to("cuda:1"), 5)       # current device is still 0
    assert torch.cuda.current_device() == 0
    for a, b in zip(out0, out1):
        assert b.device == torch.device("cuda", 1) and torch.equal(a.cpu(), b.cpu())
    rec0 = eng0.reconstruct(out0[3], out0[1])
    rec1 = eng1.reconstruct(out1[3], out1[1])
    assert torch.equal(rec0.cpu(), rec1.cpu())
    dp = importlib.import_module("audio-visual-speech-enhancement_b200.data_processor")
    assert dp.get_engine(device="cuda:1").device == torch.device("cuda", 1)
