#!/usr/bin/env python3
"""Ladder kernel time against batch size and mode (per-lane P / P = G / G with table); development aid."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecsimd_b200
from ecsimd_b200 import device as dev
from tools.quick_perf import timeit

ecsimd_b200.init(0)
nmax = 1 << 22
k = dev.synth_values(dev.empty(nmax, 1), 0xEC51D004, 0, nmax, 0)
r = dev.synth_values(dev.empty(nmax, 1), 0xEC51D003, 0, nmax, 0)
J = dev.scalar_mult_base(dev.empty(nmax, 3), r, nmax)
P = dev.from_affine(dev.empty(nmax, 3), dev.to_affine(dev.empty(nmax, 2), J, nmax), nmax)
O = dev.empty(nmax, 3)
for n in (148 * 512, 148 * 512 * 2, 148 * 512 * 13, 1 << 20, 148 * 512 * 14, 148 * 512 * 27, 1 << 21, 1 << 22):
    kk, PP, OO = dev.empty(n, 1), dev.empty(n, 3), dev.empty(n, 3)
    # SOA planes: take the first n lanes of each plane
    kk.copy_(k[:, :n])
    PP.copy_(P[:, :n])
    res = {"n": n, "waves": n / (148 * 512)}
    for name, fn in (("var", lambda: dev.scalar_mult(OO, kk, PP, n)), ("var_noquirk", lambda: dev.scalar_mult(OO, kk, PP, n, quirk=False)),
                     ("base_plain", lambda: dev.scalar_mult_base(OO, kk, n, table=False)), ("base_table", lambda: dev.scalar_mult_base(OO, kk, n))):
        ms = timeit(fn, 2)
        res[name] = {"ms": round(ms, 3), "M_per_s": round(n / ms / 1e3, 3), "ms_per_wave": round(ms / -(-n // (148 * 512)), 4)}
    print(json.dumps(res), flush=True)
