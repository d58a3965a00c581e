#!/usr/bin/env python3
"""Streaming point-op kernel timings at 2^22 points (BASELINE config 2); development aid.  ECB200_LIB selects an A/B build."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecsimd_b200
from ecsimd_b200 import device as dev
from tools.quick_perf import timeit

ecsimd_b200.init(0)
n = 1 << 22
r = dev.synth_values(dev.empty(n, 1), 0xEC51D003, 0, n, 0)
J = dev.scalar_mult_base(dev.empty(n, 3), r, n)
P = dev.from_affine(dev.empty(n, 3), dev.to_affine(dev.empty(n, 2), J, n), n)
Q, R, O = dev.empty(n, 3), dev.empty(n, 3), dev.empty(n, 3)
dev.trplu(Q, R, P, n)
res = {"lib": os.environ.get("ECB200_LIB", "default")}
res["trplu_ms"] = timeit(lambda: dev.trplu(Q, R, P, n), 5)
res["zdau_ms"] = timeit(lambda: dev.zdau(O, J, R, Q, n), 5)
res["dblu_ms"] = timeit(lambda: dev.dblu(O, J, P, n), 5)
res["zaddu_ms"] = timeit(lambda: dev.zaddu(Q, R, O, J, n), 5)
res["add_z2_1_ms"] = timeit(lambda: dev.add_z2_1(O, R, P, n), 5)
print(json.dumps(res))
