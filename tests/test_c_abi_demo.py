"""examples/c_abi_demo.c: a plain-C client of include/avse_b200.h (cudaMalloc buffers, no Python in the loop).
CPU: it compiles and links against libavse_b200.so.  GPU: its checksums equal the Python engine's on the same inputs."""
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "audio-visual-speech-enhancement_b200")
EXE = os.path.join(ROOT, "build", "c_abi_demo")


def _build():
    importlib.import_module("audio-visual-speech-enhancement_b200._native")        # builds libavse_b200.so if needed
    importlib.import_module("audio-visual-speech-enhancement_b200.build").build_library()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.check_call(["gcc", "-O1", "-std=c11", os.path.join(ROOT, "examples", "c_abi_demo.c"), "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(cuda, "include"), "-L" + PKG, "-lavse_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lm",
                           "-Wl,-rpath," + PKG, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", EXE])
    return EXE


def test_c_client_compiles_and_links():
    exe = _build()
    assert os.path.exists(exe)
    needed = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libavse_b200.so" in needed and "libtorch" not in needed and "libpython" not in needed


@pytest.mark.gpu
def test_c_client_matches_python_engine():
    import torch
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"factor0=(\S+) factor1=(\S+) sum_mixed_db=(\S+) sum_abs_recon=(\S+)", r.stdout)
    assert m, r.stdout
    f0, f1, s_db, s_rec = (float(x) for x in m.groups())
    # the same inputs in numpy (LCG + tones, as in the C source)
    B, L = 2, 16000
    S = np.zeros((B, L), np.float32)
    N = np.zeros((B, L), np.float32)
    state = 12345
    for u in range(B):
        for i in range(L):
            state = (state * 1664525 + 1013904223) & 0xffffffff
            r_ = ((state >> 8) & 0xffff) / 65536.0 - 0.5
            S[u, i] = np.float32(0.3 * np.sin(2.0 * 3.14159265358979 * (220.0 + 110.0 * u) * i / 16000.0) * (0.5 + 0.5 * np.sin(i / 1500.0)))
            N[u, i] = np.float32(0.1 * r_)
    mod = importlib.import_module("audio-visual-speech-enhancement_b200.engine")
    eng = mod.SpectralEngine(16000, 25.0, 200, device="cuda:0")
    s, n = torch.from_numpy(S).cuda(), torch.from_numpy(N).cuda()
    factor, _ = eng.snr_factor(s, n)
    mixed, speech, noise, pcm = eng.preprocess_pairs(s, n, 5)
    rec = eng.reconstruct(pcm, speech)
    assert abs(f0 - float(factor[0])) <= 1e-6 * f0 and abs(f1 - float(factor[1])) <= 1e-6 * f1
    assert abs(s_db - float(mixed.double().sum())) <= 1e-3
    assert abs(s_rec - float(rec.double().abs().sum())) <= 1e-4 * s_rec
