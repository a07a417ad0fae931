"""Probe (not product): does capturing the three launches of a forward step in a CUDA graph shrink the ~25 us of
launch gaps per step?  Usage on the GPU box: python tools/graph_step_probe.py"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

bench.set_workload(3.0, False)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
mod = importlib.import_module("audio-visual-speech-enhancement_b200.engine")
eng = mod.SpectralEngine(16000, 25.0, 200, device=dev)
B = 1000
speech, noise = bench.synth_batch(torch, B, dev, seed=0)
out = {}
L = bench.L

def step():
    factor, max_key = eng.snr_factor(speech, noise, max_key=out.get("max_key"))
    res = eng.forward_raw(speech, noise, L=L, factor=factor, n_slices=15, max_key=max_key, out=out)
    eng.floor3_(res["speech"], res["noise"], res["mixed"], max_key)
    return res

def timed(fn, n=400):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for _ in range(5): step()
print("eager  ms/step", timed(step))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
torch.cuda.current_stream().wait_stream(s)
print("graph  ms/step", timed(g.replay))
g10 = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    with torch.cuda.graph(g10, stream=s):
        for _ in range(10): step()
print("graph10 ms/step", timed(g10.replay, 40) / 10)
print("eager  ms/step", timed(step))
