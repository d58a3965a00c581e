#!/usr/bin/env python3
"""Toy issue model of one SM sub-partition running W warps over a SASS loop body.

Each warp walks the same straight-line loop body (calls/branches not taken); an instruction can
issue when (a) the warp's previous instruction's encoded stall count has elapsed and (b) the pipe
it needs can accept a warp instruction.  Pipe occupancies (cycles per warp instruction) are the
measured B200 figures of profiles/r1_microbench.md.  One instruction issues per cycle per
sub-partition.  Warps meet at the BAR.SYNC once per trip.

usage: sass_sim.py file.sass <kernel-substring> lo hi [warps] [policy]
"""
import re
import sys

from sass_stalls import ctrl, parse

FMAH = {"IMAD.WIDE": 4.1}


def classify(t):
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
    if op.startswith("IMAD.WIDE"):
        return "fma", 4.1
    if op.startswith("IMAD.HI"):
        return "fma", 5.1
    if op.startswith("IMAD"):
        return "fma", 2.0
    if op.split(".")[0] in ("IADD3", "LOP3", "SHF", "SEL", "VIMNMX3", "VIMNMX", "ISETP", "VIADD", "MOV", "PLOP3", "PRMT", "LEA", "IABS"):
        return "alu", 2.0
    return "none", 0.0


def simulate(body, nw, trips, policy, occ_scale=None):
    n = len(body)
    pc = [0] * nw
    trip = [0] * nw
    ready = [0.0] * nw          # earliest issue time of next instruction
    pipe_free = {"fma": 0.0, "alu": 0.0}
    at_bar = [False] * nw
    t = 0.0
    last = 0
    done = 0
    issued = 0
    busy = {"fma": 0.0, "alu": 0.0}
    while done < nw:
        order = list(range(nw))
        if policy == "rr":
            order = order[last + 1:] + order[:last + 1]
        elif policy == "gto":
            order = [last] + [w for w in range(nw) if w != last]
        elif policy == "hi":
            order = order[::-1]
        pick = None
        for w in order:
            if trip[w] >= trips or at_bar[w] or ready[w] > t:
                continue
            pipe, occ, stall, isbar = body[pc[w]]
            if pipe != "none" and pipe_free[pipe] > t:
                continue
            pick = w
            break
        if pick is not None:
            w = pick
            pipe, occ, stall, isbar = body[pc[w]]
            if pipe != "none":
                pipe_free[pipe] = t + occ
                busy[pipe] += occ
            ready[w] = t + max(stall, 1)
            issued += 1
            last = w
            pc[w] += 1
            if isbar:
                at_bar[w] = True
                if all(at_bar[x] or trip[x] >= trips for x in range(nw)):
                    for x in range(nw):
                        at_bar[x] = False
            if pc[w] == n:
                pc[w] = 0
                trip[w] += 1
                if trip[w] >= trips:
                    done += 1
        t += 1.0
    return t / trips, issued / t, busy["fma"] / t, busy["alu"] / t


def main():
    path, kernel, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    nw = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    policy = sys.argv[6] if len(sys.argv) > 6 else "gto"
    ins = [i for i in parse(path, kernel) if lo <= i[0] <= hi]
    body = []
    for a, txt, w1, w2 in ins:
        pipe, occ = classify(txt)
        body.append((pipe, occ, ctrl(w2)["stall"], "BAR.SYNC" in txt))
    for pol in ([policy] if len(sys.argv) > 6 else ["gto", "rr", "hi"]):
        clk, ipc, f, a = simulate(body, nw, 6, pol)
        print("policy %-3s warps %d: %.0f clk/trip  (%.0f per warp-trip)  ipc %.2f  fma busy %.2f  alu busy %.2f" % (pol, nw, clk, clk / nw, ipc, f, a))


if __name__ == "__main__":
    main()
