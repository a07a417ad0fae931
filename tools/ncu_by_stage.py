#!/usr/bin/env python
"""Attribute an .ncu-rep's per-SASS-instruction counters of one kernel to source functions ("stages").

usage: python tools/ncu_by_stage.py <rep.ncu-rep> <lib.so> <kernel-mangled-name> [frames_per_launch]
The .so must be the build the profile was taken from (compiled with -lineinfo).  nvdisasm gives file:line per
instruction; lines inside avse_dft.cuh (shared codelets) inherit the stage of the closest preceding stage-specific line."""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

# NCU_KERNEL=<regex> selects one kernel of a multi-kernel report
KFILTER = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []

rep, so, kname = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
F = float(sys.argv[4]) if len(sys.argv) > 4 else 301000.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if (".text." + kname + ":") in txt:
        lines = txt.split(".text." + kname + ":")[1].splitlines()
        break
assert lines is not None, "kernel not found"
# function line ranges per file
def func_ranges(path):
    out = []
    cur = None
    for i, l in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*>\s*)?AVSE_HD\s+\S+\s+(\w+)\(", l) or re.match(r"^__global__.*\s(\w+)\(", l)
        if m:
            cur = m.group(1)
        out.append(cur)
    return out
ranges = {}
def stage_of(f, ln):
    if f not in ranges:
        ranges[f] = func_ranges(f) if os.path.exists(f) else []
    r = ranges[f]
    return (r[ln - 1] if 0 < ln <= len(r) else None) or os.path.basename(f)
insts = []   # (offset, file, line)
cf, cl = None, 0
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cf, cl = m.group(1), int(m.group(2)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*);", l)
    if m:
        insts.append((int(m.group(1), 16), cf, cl, m.group(2)))
    if l.startswith("//---") or ".section" in l:
        if insts: break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + KFILTER, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = []
for r in rows[2:]:                       # a report with several launches repeats the header: keep the first launch only
    if len(r) < len(h):
        continue
    if r[ix["Address"]] == "Address":
        break
    data.append(r)
base = int(data[0][ix["Address"]], 16)
byoff = {int(r[ix["Address"]], 16) - base: r for r in data}
agg = collections.defaultdict(lambda: collections.Counter())
label = "prologue"
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for off, f, ln, text in insts:
    r = byoff.get(off)
    if r is None: continue
    if f and not f.endswith("avse_dft.cuh"):
        label = stage_of(f, ln)
    a = agg[label]
    a["inst"] += int(r[ix["Instructions Executed"]]); a["samples"] += int(r[ix["# Samples"]])
    a["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    for c in stall_cols:
        a[c] += int(r[ix[c]] or 0)
tot = sum(a["samples"] for a in agg.values())
print("%-28s %9s %8s %9s   top stalls" % ("stage", "inst/frm", "samp %", "smem wf/frm"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    tops = sorted(((a[c], c) for c in stall_cols), reverse=True)[:4]
    print("%-28s %9.1f %8.1f %9.1f   %s" % (k, a["inst"] / F, 100.0 * a["samples"] / max(tot, 1), a["wf"] / F,
                                          ", ".join("%s %.0f%%" % (c[6:], 100.0 * v / max(a["samples"], 1)) for v, c in tops)))
