#!/usr/bin/env python3
"""tiny driver for profiling the streaming multiply at 2^24 lanes (one launch after warm-up)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ecsimd_b200
from ecsimd_b200 import device as dev
ecsimd_b200.init(0)
n = 1 << 24
a = dev.synth_values(dev.empty(n, 1), 0xEC51D001, 0, n, 1)
b = dev.synth_values(dev.empty(n, 1), 0xEC51D002, 0, n, 1)
o = dev.empty(n, 1)
for _ in range(4):
    dev.mgry_mul(o, a, b, n)
torch.cuda.synchronize()
print("ok")
