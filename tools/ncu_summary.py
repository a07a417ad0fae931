#!/usr/bin/env python
"""Summarise an .ncu-rep of one kernel: headline raw metrics + executed-instruction mix per opcode
(usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [frames_per_launch])."""
import collections
import csv
import io
import os
import subprocess
import sys

# NCU_KERNEL=<regex> selects one kernel of a multi-kernel report
KFILTER = ["-k", "regex:" + os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else []

rep = sys.argv[1]
F = float(sys.argv[2]) if len(sys.argv) > 2 else 301000.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + KFILTER, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        v = m[k]
        try:
            fv = float(v.replace(",", ""))
            extra = "   (%.1f per frame)" % (fv / F) if fv > 1e6 and "bytes" not in k else ""
        except ValueError:
            extra = ""
        print("%-75s %s%s" % (k, v, extra))
for k in hdr:
    if "issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k.lower():
        print("%-75s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), m[k]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + KFILTER, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
op = collections.Counter(); st = collections.Counter(); wf = collections.Counter()
tot = tots = 0
for r in rows[2:]:
    if len(r) < len(h):
        continue
    if r[ix["Source"]] == "Source":       # a report with several launches repeats the header: keep the first launch only
        break
    toks = r[ix["Source"]].split()
    o = toks[1] if toks[0].startswith("@") else toks[0]
    o = o.split(".")[0]
    n = int(r[ix["Instructions Executed"]]); sm = int(r[ix["# Samples"]])
    op[o] += n; st[o] += sm; tot += n; tots += sm
    wf[o] += int(r[ix["L1 Wavefronts Shared"]] or 0)
print("warp-instructions per frame: %.1f  (stall samples %d)" % (tot / F, tots))
for o, n in op.most_common(28):
    print("  %-10s %8.1f /frame   samples %5.1f %%   smem wavefronts/frame %6.1f" % (o, n / F, 100.0 * st[o] / max(tots, 1), wf[o] / F))
