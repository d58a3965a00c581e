#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.
usage: summarize_ncu.py <rep> <out.md> [title]"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    md = ["# %s" % title, "", "source: `%s` (ncu --set full --clock-control none)" % rep, ""]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        md += ["## kernel `%s`" % name[:110], "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                md.append("| %s | %s | %s |" % (k, r[i], units[i]))
        md += ["", "| stall reason (warps per issue) | value |", "|---|---|"]
        for s in ("wait", "math_pipe_throttle", "dispatch_stall", "not_selected", "no_instruction", "barrier", "short_scoreboard",
                  "long_scoreboard", "branch_resolving", "lg_throttle"):
            k = STALLS % s
            if k in hdr:
                md.append("| %s | %s |" % (s, r[hdr.index(k)]))
        md.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        h = srows[1]
        ix = {n: i for i, n in enumerate(h)}
        agg = collections.defaultdict(collections.Counter)
        cnt = collections.Counter()
        execd = collections.Counter()
        for r in srows[2:]:
            if len(r) < len(h):
                continue
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
            if not m:
                continue
            op = m.group(1)
            cls = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else (".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0])
            cnt[cls] += 1
            try:
                execd[cls] += int(r[ix["Instructions Executed"]])
            except Exception:
                pass
            for k in ("stall_dispatch", "stall_math", "stall_wait", "stall_not_selected", "stall_selected", "stall_no_inst"):
                try:
                    agg[cls][k] += int(r[ix[k]])
                except Exception:
                    pass
        md += ["## SASS by opcode class (static count, warp-instructions executed, stall samples)", "",
               "| class | static | executed | dispatch | math | wait | not_selected | selected | no_inst |", "|---|---|---|---|---|---|---|---|---|"]
        for c, n in cnt.most_common(16):
            a = agg[c]
            md.append("| %s | %d | %d | %d | %d | %d | %d | %d | %d |" % (c, n, execd[c], a["stall_dispatch"], a["stall_math"], a["stall_wait"],
                                                                   a["stall_not_selected"], a["stall_selected"], a["stall_no_inst"]))
    open(out, "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
