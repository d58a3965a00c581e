#!/usr/bin/env python3
"""Print the pipe class of every instruction of a SASS loop as one character
(W = IMAD.WIDE, f = other FMA-pipe, a = ALU, . = other) to see how well ptxas mixed the pipes."""
import sys, textwrap
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
from sass_sim import classify
from sass_stalls import parse
ins = [i for i in parse(sys.argv[1], sys.argv[2]) if int(sys.argv[3], 16) <= i[0] <= int(sys.argv[4], 16)]
s = ""
for a, txt, w1, w2 in ins:
    pipe, occ = classify(txt)
    s += "W" if occ > 4 else ("f" if pipe == "fma" else ("a" if pipe == "alu" else "."))
print("\n".join(textwrap.wrap(s, 160)))
