"""Probe (not product): forward + inverse time of a ragged batch (utterance lengths uniform in [lo * L, L], rows zero-padded to L)
next to the full-length batch of the same shape.  Usage on the GPU box: python tools/ragged_probe.py [lo]"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

lo = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
bench.set_workload(3.0, False)
dev = torch.device("cuda", 0)
mod = importlib.import_module("audio-visual-speech-enhancement_b200.engine")
eng = mod.SpectralEngine(16000, 25.0, 200, device=dev)
B, L = 1000, bench.L
speech, noise = bench.synth_batch(torch, B, dev, seed=0)
g = torch.Generator(device="cpu").manual_seed(1)
lens = (torch.rand(B, generator=g) * (1.0 - lo) + lo).mul(L).to(torch.int32).to(dev)
mask = torch.arange(L, device=dev)[None, :] < lens[:, None]
sp_r, nz_r = speech * mask, noise * mask


def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out_f = eng.preprocess_pairs(speech, noise, 15)
out_r = eng.preprocess_pairs(sp_r, nz_r, 15, lengths=lens)
print("forward step  full %.4f ms   ragged %.4f ms" % (timed(lambda: eng.preprocess_pairs(speech, noise, 15)),
                                                       timed(lambda: eng.preprocess_pairs(sp_r, nz_r, 15, lengths=lens))))
print("inverse       full %.4f ms   ragged %.4f ms" % (timed(lambda: eng.reconstruct(out_f[3], out_f[1])),
                                                       timed(lambda: eng.reconstruct(out_r[3], out_r[1], lengths=lens))))
