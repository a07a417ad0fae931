#!/usr/bin/env python
"""List the shared-memory instructions of a kernel with the most excessive (bank-conflict) wavefronts, with source lines.
usage: python tools/ncu_smem_excess.py <rep.ncu-rep> <lib.so> <kernel-mangled-name> [frames]"""
import csv, glob, io, os, re, subprocess, sys, tempfile

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
        lines = txt.split(".text." + kname + ":")[1].splitlines(); break
assert lines is not None
loc = {}; cf, cl = None, 0
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cf, cl = os.path.basename(m.group(1)), int(m.group(2)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*);", l)
    if m: loc[int(m.group(1), 16)] = (cf, cl)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + KFILTER, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) >= len(h)]
base = int(data[0][ix["Address"]], 16)
out = []
for r in data:
    try:
        w = float(r[ix["L1 Wavefronts Shared"]] or 0); e = float(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    except ValueError:
        continue
    if w > 0:
        off = int(r[ix["Address"]], 16) - base
        out.append((e, w, off, r[ix["Source"]].strip()[:60], loc.get(off)))
tot_e = sum(o[0] for o in out); tot_w = sum(o[1] for o in out)
print("shared wavefronts/frame %.1f, excessive %.1f" % (tot_w / F, tot_e / F))
agg = {}
for e, w, off, s, lc in out:
    k = lc
    a = agg.setdefault(k, [0.0, 0.0, s]); a[0] += e; a[1] += w
for k, (e, w, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print("%-28s excess %6.2f  total %6.2f /frame   %s" % ("%s:%s" % k if k else "?", e / F, w / F, s))
