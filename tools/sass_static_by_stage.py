#!/usr/bin/env python
"""Static SASS instruction count of one kernel per source function (needs -lineinfo).
usage: python tools/sass_static_by_stage.py <lib.so> <kernel-mangled-name>"""
import collections, glob, os, re, subprocess, sys, tempfile
so, kname = os.path.abspath(sys.argv[1]), sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if (".text." + kname + ":") in txt:
        lines = txt.split(".text." + kname + ":")[1].splitlines()
        break
assert lines is not None, "kernel not found"
ranges = {}
def func_ranges(path):
    out, cur = [], None
    for l in open(path):
        m = re.match(r"^(?:template.*>\s*)?AVSE_HD\s+\S+\s+(\w+)\(", l) or re.match(r"^__global__.*\s(\w+)\(", l)
        if m: cur = m.group(1)
        out.append(cur)
    return out
def stage_of(f, ln):
    if f not in ranges: ranges[f] = func_ranges(f) if os.path.exists(f) else []
    r = ranges[f]
    return (r[ln - 1] if 0 < ln <= len(r) else None) or os.path.basename(f)
cnt = collections.Counter(); cf, cl = None, 0; n = 0
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cf, cl = m.group(1), int(m.group(2)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*);", l):
        cnt[stage_of(cf, cl) if cf else "?"] += 1; n += 1
    if (l.startswith("//---") or ".section" in l) and n: break
for k, v in cnt.most_common(): print("%6d  %s" % (v, k))
print("%6d  total" % n)
