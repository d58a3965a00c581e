#!/usr/bin/env python3
"""Differential fuzz of the GPU engine against the CPU oracle on EVERY lane (not sampled): random scalars and
points (a share of them out-of-contract bit patterns), all three layouts, variable base / generator / fused
affine.  usage: fuzz_vs_oracle.py [seeds] [lanes]
Use at least 9 seeds (3 layouts x 3 fixed-base layouts) and more than 2 x 148 x 512 lanes: a kernel can be right on a
partly filled GPU and wrong at full occupancy (profiles/r2e_recolor_bug/README.md)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs
import ecsimd_b200
from ecsimd_b200 import host as eng

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
ecsimd_b200.init(0)
orc = _libs.oracle(nt=os.cpu_count() or 1)
G = np.concatenate([_libs.to_words([_libs.GX_INT]), _libs.to_words([_libs.GY_INT])], axis=1)
GJ = orc.from_affine(np.repeat(G, n, axis=0))
bad = 0
t0 = time.time()
for s in range(seeds):
    rnd = np.random.RandomState(1000 + s)
    k = _libs.raw256(0xF00D0000 + s, n)
    P = orc.from_affine(orc.to_affine(orc.scalar_mult(_libs.raw256(0xBEEF0000 + s, n), GJ)))
    # out-of-contract lanes: arbitrary patterns, all-ones words, values >= p
    m = n // 16
    P[:m, :16] = _libs.raw256(0xABCD0000 + s, 2 * m).reshape(m, 16)
    P[m:2 * m, rnd.randint(0, 16, size=m)] = 0xFFFFFFFF
    k[:64] &= rnd.randint(0, 2, size=(64, 8)).astype(np.uint32) * np.uint32(0xFFFFFFFF)       # sparse scalars
    want = orc.scalar_mult(k, P)
    layout = ("lane", "pack4", "soa")[s % 3]
    conv = {"lane": (lambda x, nc: x, lambda x, nc: x), "pack4": (eng.lane_to_pack4, eng.pack4_to_lane), "soa": (eng.lane_to_soa, eng.soa_to_lane)}[layout]
    got = conv[1](eng.scalar_mult(conv[0](k, 1), conv[0](P, 3), layout=layout), 3)
    e1 = int((got != want).any(axis=1).sum())
    wantg = orc.scalar_mult(k, GJ)
    # fixed base: every layout's table instance and plain instance over the seeds (each is its own kernel)
    blayout = ("lane", "pack4", "soa")[(s // 3) % 3]
    bconv = {"lane": (lambda x, nc: x, lambda x, nc: x), "pack4": (eng.lane_to_pack4, eng.pack4_to_lane), "soa": (eng.lane_to_soa, eng.soa_to_lane)}[blayout]
    e2 = int((bconv[1](eng.scalar_mult_base(bconv[0](k, 1), layout=blayout, table=True), 3) != wantg).any(axis=1).sum())
    e2 += int((bconv[1](eng.scalar_mult_base(bconv[0](k, 1), layout=blayout, table=False), 3) != wantg).any(axis=1).sum())
    e3 = int((eng.scalar_mult_affine(k, P) != orc.to_affine(want)).any(axis=1).sum())
    e4 = int((eng.mgry_sqr(P[:, :8]) != orc.mgry_sqr(P[:, :8])).any(axis=1).sum()) + int((eng.mgry_mul(P[:, :8], P[:, 8:16]) != orc.mgry_mul(P[:, :8], P[:, 8:16])).any(axis=1).sum())
    bad += e1 + e2 + e3 + e4
    print({"seed": s, "layout": layout, "base_layout": blayout, "lanes": n, "mismatch_var": e1, "mismatch_base": e2, "mismatch_affine": e3, "mismatch_field": e4}, flush=True)
print({"total_mismatches": bad, "seconds": round(time.time() - t0, 1)})
sys.exit(1 if bad else 0)
