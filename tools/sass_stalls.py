#!/usr/bin/env python3
"""Static issue-latency analysis of a SASS loop (sm_100a): sums the stall counts ptxas encoded
in the control bits of every instruction between two addresses.  The sum is the time ONE warp
needs per trip when it never waits for a pipe or another warp, i.e. the dependency-latency
floor of the loop; compare with (#IMAD.WIDE x 4.1 clk) x warps per sub-partition.

usage: sass_stalls.py file.sass <kernel-substring> [lo hi]   (addresses in hex; default = largest backward branch)
"""
import collections
import re
import sys


def parse(path, kernel):
    ins, on = [], False
    prev = None
    for l in open(path):
        if "Function :" in l:
            on = kernel in l
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", l)
        if m:
            prev = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
            ins.append(prev)
            continue
        m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", l)
        if m and prev is not None and prev[3] is None:
            prev[3] = int(m.group(1), 16)
    return ins


def ctrl(w2):
    c = w2 >> 41
    return {"stall": c & 0xF, "yield": (c >> 4) & 1, "wbar": (c >> 5) & 7, "rbar": (c >> 8) & 7, "wait": (c >> 11) & 0x3F}


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    ins = parse(path, kernel)
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    else:
        best = (0, 0, 0)
        for a, t, _, _ in ins:
            m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > best[0]:
                best = (a - int(m.group(1), 16), int(m.group(1), 16), a)
        lo, hi = best[1], best[2]
    loop = [i for i in ins if lo <= i[0] <= hi]
    tot = 0
    by = collections.defaultdict(lambda: [0, 0])
    for a, t, w1, w2 in loop:
        c = ctrl(w2)
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
        op = op.split(".")[0] + (".WIDE" if ".WIDE" in op else "") + (".X" if op.endswith(".X") or ".X." in op else "")
        by[op][0] += 1
        by[op][1] += c["stall"]
        tot += c["stall"]
    n = len(loop)
    wide = sum(v[0] for k, v in by.items() if "WIDE" in k)
    print("loop 0x%x..0x%x: %d instr, %d IMAD.WIDE, %d others; sum of stall counts %d clk (%.2f/instr)" % (lo, hi, n, wide, n - wide, tot, tot / max(n, 1)))
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print("  %-14s n=%5d stall_sum=%6d avg=%.2f" % (k, v[0], v[1], v[1] / v[0]))


if __name__ == "__main__":
    main()
