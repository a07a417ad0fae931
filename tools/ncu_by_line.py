#!/usr/bin/env python
"""Per-SOURCE-LINE attribution of an .ncu-rep's per-SASS counters for one kernel and one source file: executed instructions per
STFT frame, share of the stall samples, static instruction count and opcode mix of every line.  This is how round 2 found the
kernel prologue's 13 serial L2 round trips, the per-tile pointer arithmetic and the divergent dB branch (profiles/README.md).

usage: python tools/ncu_by_line.py <rep.ncu-rep> <lib.so> <kernel-mangled-name> <source-file-suffix>   (NCU_KERNEL=<regex> as in ncu_by_stage.py)
The .so must be the build the profile was taken from (-lineinfo)."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, so, kname, want = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], sys.argv[4]
F = 301000.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines=None
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    if (".text." + kname + ":") in txt:
        lines = txt.split(".text." + kname + ":")[1].splitlines(); break
insts=[]; cf,cl=None,0
for l in lines:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cf, cl = m.group(1), int(m.group(2)); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*);", l)
    if m: insts.append((int(m.group(1),16), cf, cl, m.group(2)))
    if l.startswith("//---") or ".section" in l:
        if insts: break
KF = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"] + KF, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); h=rows[1]; ix={n:i for i,n in enumerate(h)}
data=[]
for r in rows[2:]:
    if len(r)<len(h): continue
    if r[ix["Address"]]=="Address": break
    data.append(r)
base=int(data[0][ix["Address"]],16)
byoff={int(r[ix["Address"]],16)-base:r for r in data}
agg=collections.defaultdict(collections.Counter)
for off,f,ln,text in insts:
    r=byoff.get(off)
    if r is None or not f or not f.endswith(want): continue
    a=agg[ln]; a["inst"]+=int(r[ix["Instructions Executed"]]); a["samp"]+=int(r[ix["# Samples"]]); a["n"]+=1
    a["ops:"+text.split()[0 if not text.startswith("@") else 1].split(".")[0]]+=1
tot=sum(int(r[ix["# Samples"]]) for r in data)
srcl=open([f for _,f,_,_ in insts if f and f.endswith(want)][0]).read().splitlines()
for ln in sorted(agg):
    a=agg[ln]
    if a["samp"]*100/tot<0.08: continue
    ops=" ".join("%s:%d"%(k[4:],v) for k,v in a.most_common() if k.startswith("ops:"))
    print("%5d %7.1f i/f %5.2f%% n=%3d | %s | %s"%(ln,a["inst"]/F,a["samp"]*100.0/tot,a["n"],srcl[ln-1].strip()[:90],ops[:80]))
