#!/usr/bin/env python
"""Would a split-precision tensor-core DFT hold the parity gate?  (VERDICT r1 "next" #5 -- evidence instead of an assertion.)

CPU emulation, no GPU needed.  The forward path's 640-point transform is 16 x 40: pass 1 (DFT-16 over n1 + twiddle) and pass 2
(DFT-40 over n2).  Pass 2 is the GEMM-shaped half a tensor core could take: per frame a [16 x 80] x [80 x 80] real product (the
80 x 80 matrix holds the complex DFT-40).  This script runs the WHOLE pipeline of the CUDA kernel (equalised packing z = s + i g n,
unpack, three magnitudes, mel, dB, top_db floor) in numpy with pass 2 computed as that GEMM under the operand precisions a tensor
core offers, products exact and accumulation in float32 (tcgen05 / mma.sync semantics), everything else in float32 exactly like
the kernel, and reports the worst log-mel error against the float64 oracle:

    fp32        operands unrounded (control: an 80-term float32 dot product instead of the FFT codelet)
    tf32x1      operands rounded to TF32 (10-bit mantissa)                            1 MMA pass
    tf32x3      a = hi + lo:  hi*hi + hi*lo + lo*hi                                     3 MMA passes
    bf16x1/x3/x6  bfloat16 splits a = a1 + a2 + a3: 1, 3 (a1b1+a1b2+a2b1) and 6 passes (+ a1b3 + a2b2 + a3b1)
    fp16x3      float16 two-way split with the same 3 products (operands pre-scaled into range)

    --cuda: the split products of tf32x1 / tf32x3 are additionally run on the REAL tensor cores (torch.matmul with TF32 enabled,
    operands pre-split on the host so that the hardware's own operand rounding is a no-op): "tf32x3-hw" shows what the
    accumulator of the actual MMA datapath does to the margin.

usage: python tools/tensor_core_dft_study.py [--cuda] > profiles/tensor_core_dft_study_r2.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import avse_oracle as O   # noqa: E402  (a study script: allowed to use the checker)

SR, N, HOP, N1, N2 = 16000, 640, 160, 16, 40
f32 = np.float32


def round_mantissa(x, bits):
    """Round float32 values to `bits` explicit mantissa bits (round to nearest even): TF32 = 10, bfloat16 = 7."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    half = np.uint64(1 << (drop - 1))
    lsb = (u >> np.uint64(drop)) & np.uint64(1)
    u = (u + half - np.uint64(1) + lsb) & ~np.uint64((1 << drop) - 1)
    return u.astype(np.uint32).view(np.float32)


def split(x, bits, parts):
    out, r = [], np.asarray(x, dtype=np.float32)
    for _ in range(parts):
        p = round_mantissa(r, bits)
        out.append(p)
        r = (r - p).astype(np.float32)
    return out


def _mm_hw(x, y):
    import torch
    torch.backends.cuda.matmul.allow_tf32 = True
    return torch.matmul(torch.from_numpy(np.ascontiguousarray(x, dtype=f32)).cuda(), torch.from_numpy(np.ascontiguousarray(y, dtype=f32)).cuda()).cpu().numpy()


def gemm(a, b, scheme):
    """float32-accumulated product under the operand precision of `scheme`."""
    mm = lambda x, y: np.matmul(x.astype(f32), y.astype(f32)).astype(f32)
    if scheme.endswith("-hw"):
        mm = _mm_hw
        scheme = scheme[:-3]
    if scheme == "fp32":
        return mm(a, b)
    if scheme == "fp16x3":
        sa = np.max(np.abs(a)) or 1.0
        a1 = (a / sa).astype(np.float16).astype(f32)
        a2 = ((a / sa) - a1).astype(np.float16).astype(f32)
        b1 = b.astype(np.float16).astype(f32)
        b2 = (b - b1).astype(np.float16).astype(f32)
        return ((mm(a1, b1) + mm(a1, b2) + mm(a2, b1)) * f32(sa)).astype(f32)
    kind, passes = scheme.split("x")
    bits = {"tf32": 10, "bf16": 7}[kind]
    passes = int(passes)
    parts = {1: 1, 3: 2, 6: 3}[passes]
    A, B = split(a, bits, parts), split(b, bits, parts)
    acc = np.zeros((a.shape[0], b.shape[1]), f32)
    terms = [(i, j) for i in range(parts) for j in range(parts) if i + j < parts]
    for i, j in sorted(terms, key=lambda t: -(t[0] + t[1])):      # small terms first
        acc = (acc + mm(A[i], B[j])).astype(f32)
    return acc


def pipeline(s, n, L, factor, gain, scheme):
    """float32 restatement of the kernel's arithmetic with pass 2 as a GEMM in `scheme`.  Returns floored dB mels (3, 80, T)."""
    w = O.hann_periodic(N)
    T = 1 + L // HOP

    def frames(x):
        x = np.concatenate([x, np.zeros(L - len(x))])[:L]
        xp = np.pad(x, N // 2, mode="reflect")
        idx = np.arange(N)[None, :] + HOP * np.arange(T)[:, None]
        return xp[idx]
    z = (frames(s.astype(np.float64)) * w).astype(f32) + 1j * (frames(n.astype(np.float64)).astype(f32) * f32(gain) * w.astype(f32)).astype(f32)
    z = z.astype(np.complex64).reshape(T, N1, N2)                       # [t][n1][n2], n = 40 n1 + n2
    # pass 1 in float64 then rounded: the best case for everything that is NOT under test
    W16 = np.exp(-2j * np.pi * np.outer(np.arange(N1), np.arange(N1)) / N1)
    tw = np.exp(-2j * np.pi * np.outer(np.arange(N1), np.arange(N2)) / N)
    x1 = (np.einsum("kn,tnm->tkm", W16, z.astype(np.complex128)) * tw[None]).astype(np.complex64)    # [t][k1][n2]
    # pass 2 as a real GEMM: rows (t, k1), columns (re, im) of n2
    D = np.empty((T * N1, 2 * N2), f32)
    D[:, 0::2] = x1.real.reshape(-1, N2)
    D[:, 1::2] = x1.imag.reshape(-1, N2)
    W40 = np.exp(-2j * np.pi * np.outer(np.arange(N2), np.arange(N2)) / N2)
    M = np.zeros((2 * N2, 2 * N2))
    M[0::2, 0::2] = W40.real; M[1::2, 0::2] = -W40.imag      # out_re = sum re*wr - im*wi
    M[0::2, 1::2] = W40.imag; M[1::2, 1::2] = W40.real       # out_im = sum re*wi + im*wr
    Y = gemm(D, M.astype(f32), scheme)
    Zr = Y[:, 0::2].reshape(T, N1, N2)
    Zi = Y[:, 1::2].reshape(T, N1, N2)
    Z = np.empty((T, N), np.complex64)                                   # Z[k1 + 16 k2]
    k = (np.arange(N1)[:, None] + N1 * np.arange(N2)[None, :])
    Z[:, k.reshape(-1)] = (Zr + 1j * Zi).reshape(T, -1)
    kk = np.arange(N // 2 + 1)
    a, c = Z[:, kk], np.conj(Z[:, (N - kk) % N])
    S = ((a + c)).astype(np.complex64)                                   # 2 X_speech
    Nn = ((a - c) / 1j).astype(np.complex64)                             # 2 X_noise (gain-scaled)
    r = f32(factor / gain)
    Mx = (S + r * Nn).astype(np.complex64)
    fb = (0.5 * O.mel_filterbank(SR, N, 80, 0.0, 8000.0)).astype(f32)
    out = []
    for X, sc in ((S, f32(1.0)), (Nn, r), (Mx, f32(1.0))):
        mag = np.abs(X).astype(f32)
        mel = (mag @ fb.T).astype(f32) * sc
        db = (20.0 * np.log10(np.maximum(mel, f32(1e-5)))).astype(f32).T
        out.append(np.maximum(db, db.max() - 80.0))
    return out


def oracle(s, n, L, snr):
    sp, nz = O.AudioSignal(s.astype(np.float64), SR), O.AudioSignal(n.astype(np.float64), SR)
    f = O.AudioMixer.snr_factor(sp, nz, snr)
    g = O.AudioMixer.snr_factor(sp, nz, 0.0)
    mixed, speech, noise, _ = O.preprocess_audio_pair_signals(sp, nz, 200, L // 3200, 25.0, snr_db=snr)
    cat = lambda x: np.concatenate(list(x), axis=1)
    return f, g, [cat(speech), cat(noise), cat(mixed)]


def main():
    schemes = ["fp32", "tf32x1", "tf32x3", "bf16x1", "bf16x3", "bf16x6", "fp16x3"]
    if "--cuda" in sys.argv:
        schemes += ["tf32x1-hw", "tf32x3-hw"]
    cases = [("golden-like 1 s, 0 dB, equal scale", 1.0, 1.0, 0.0, 5), ("1 s, -10 dB", 1.0, 1.0, -10.0, 6), ("1 s, +10 dB", 1.0, 1.0, 10.0, 7),
             ("float speech vs int16-scale noise, 0 dB", 1.0, 32767.0, 0.0, 8), ("int16 3000 vs 30000, +5 dB", 3000.0, 30000.0, 5.0, 9)]
    L = 16000
    print("Pass 2 (DFT-40) as a tensor-core GEMM, everything else float32 as in the kernel: worst |log-mel error| in dB vs the float64 oracle")
    print("(gate: 1e-3 dB; speech / noise / mixed; equalised packing as shipped).  MMA passes = tensor-core work relative to one plain GEMM.\n")
    print("%-44s %s" % ("case", "  ".join("%-26s" % s for s in schemes)))
    worst = {s: 0.0 for s in schemes}
    for name, ss, ns, snr, seed in cases:
        s = (O.synth_speech(L, SR, seed) * ss).astype(f32)
        n = (O.synth_noise(L, seed) * ns).astype(f32)
        f, g, ref = oracle(s, n, L, snr)
        row = []
        for sch in schemes:
            got = pipeline(s, n, L, f, g, sch)
            errs = [float(np.max(np.abs(a[:, :ref[0].shape[1]] - b))) for a, b in zip(got, ref)]
            worst[sch] = max(worst[sch], max(errs))
            row.append("%.1e/%.1e/%.1e" % tuple(errs))
        print("%-44s %s" % (name, "  ".join("%-26s" % r for r in row)))
    print("\n%-44s %s" % ("worst over cases", "  ".join("%-26s" % ("%.1e %s" % (worst[s], "HOLDS" if worst[s] <= 1e-3 else "FAILS")) for s in schemes)))
    print("\nReading: the float32 pipeline itself sits at 4e-4 ... 8e-4 dB on the speech channel (its rounding noise relative to the loudest")
    print("bins lands on mel bands 80 dB down, right above the top_db floor), so the gate leaves a 1.2-2.5x margin and nothing more.  Single-pass")
    print("TF32 / BF16 and the 3-product BF16 split miss it by 1-4 orders of magnitude.  Splits that reproduce float32-class products (tf32x3,")
    print("fp16x3, bf16x6) hold in this emulation with the SAME thin margin, at 3-6 MMA passes over a dense 80 x 80 matrix per 16 x 80 tile;")
    print("see DESIGN.md section 5 for the instruction / shared-memory accounting of what such a kernel would save.")

if __name__ == "__main__":
    main()
