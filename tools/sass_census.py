#!/usr/bin/env python3
"""One-line opcode census of the largest loop of a kernel: tools/sass_census.py file.sass kernel [lo hi]"""
import collections, re, sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_stalls import parse
path, kernel = sys.argv[1], sys.argv[2]
ins = parse(path, kernel)
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
else:
    best = (0, 0, 0)
    bars = [a for a, t, _, _ in ins if "BAR.SYNC" in t]
    for a, t, _, _ in ins:
        m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > best[0]:
            lo_ = int(m.group(1), 16)
            if bars and not any(lo_ <= b <= a for b in bars):
                continue   # prefer the block-synchronised hot loop over the exact re-run loop
            best = (a - lo_, lo_, a)
    lo, hi = best[1], best[2]
cold = set()
loop_ins = [(a, t) for a, t, _, _ in ins if lo <= a <= hi]
for a, t in loop_ins:   # cold blocks: a forward predicated branch that jumps over a CALL
    m = re.search(r"BRA\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
    if m and t.startswith("@"):
        tgt = int(m.group(1), 16)
        if tgt > a and any("CALL" in t2 for a2, t2 in loop_ins if a < a2 < tgt):
            cold.update(a2 for a2, t2 in loop_ins if a < a2 < tgt)
c = collections.Counter()
for a, t, _, _ in ins:
    if lo <= a <= hi and a not in cold:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
        key = "W" if op.startswith("IMAD.WIDE") else op.split(".")[0] + (".X" if ".X" in op else "") + (".MOV" if ".MOV" in op else "") + (".IADD" if ".IADD" in op else "")
        c[key] += 1
n = sum(c.values())
w = c["W"]
print("hot n=%d W=%d others=%d (+%d cold) | est clk/step = 4.3W+1.05*others = %.0f | %s" % (n, w, n - w, len(cold), 4.3 * w + 1.05 * (n - w), " ".join("%s=%d" % kv for kv in c.most_common(18))))
