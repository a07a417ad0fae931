#!/usr/bin/env python
"""SASS opcode histogram of the shipped library's hot kernels (static instruction counts per kernel).
usage: python tools/sass_histogram.py [lib.so] > profiles/sass_histogram_r2.txt
What to look for (B200_PROFILING.md): FFMA2 / FADD2 / FMUL2 = packed FP32, sm_100-only; UBLKCP / UTMALDG = bulk / TMA copies;
UTC*MMA / LDTM / STTM = tcgen05 + TMEM (none here: the path has no GEMM-shaped stage)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "audio-visual-speech-enhancement_b200", "libavse_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEEP = ("avse_forward4_kernel", "avse_inverse8_kernel", "avse_snr_factor_kernel", "avse_floor_inplace_kernel", "avse_mel_to_coef_kernel")
print("static SASS instruction counts, %s (cuobjdump -sass)\n" % os.path.basename(lib))
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    if not any(k in name for k in KEEP):
        continue
    ops = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", part, flags=re.M):
        ops[m.group(1)] += 1
    total = sum(ops.values())
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    packed = ops["FFMA2"] + ops["FADD2"] + ops["FMUL2"]
    print("%s\n  total %d   packed-FP32 (FFMA2+FADD2+FMUL2) %d   UBLKCP/UTMA* %d   UTC*MMA %d   LDTM/STTM %d" % (
        demangled, total, packed, sum(v for k, v in ops.items() if k.startswith(("UBLKCP", "UTMA"))),
        sum(v for k, v in ops.items() if k.startswith("UTC") and "MMA" in k), ops["LDTM"] + ops["STTM"]))
    print("  " + "  ".join("%s %d" % kv for kv in ops.most_common(24)) + "\n")
