#!/usr/bin/env python3
"""One-off soak at BASELINE configs[4] total size on ONE GPU: 2^26 variable-base lanes (15 GB of operands),
sampled oracle parity, determinism of the order-independent checksum, table vs plain generator path."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _libs
import ecsimd_b200
from ecsimd_b200 import device as dev

ecsimd_b200.init(0)
n = 1 << 26
t0 = time.time()
k = dev.synth_values(dev.empty(n, 1), 0xEC51D004, 0, n, 0)
r = dev.synth_values(dev.empty(n, 1), 0xEC51D003, 0, n, 0)
J = dev.scalar_mult_base(dev.empty(n, 3), r, n)
P = dev.from_affine(dev.empty(n, 3), dev.to_affine(dev.empty(n, 2), J, n), n)
del J, r
out = dev.empty(n, 3)
torch.cuda.synchronize(); t1 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dev.scalar_mult(out, k, P, n); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
c1 = dev.checksum(out)
idx = np.concatenate([np.arange(0, 64), np.arange(n - 64, n), np.random.RandomState(3).randint(0, n, 128)])

def lanes(t, nc):
    sel = t[:, torch.as_tensor(idx, device=t.device), :].cpu().numpy().view(np.uint32)
    m = len(idx)
    return np.ascontiguousarray(sel.reshape(nc, 2, m, 4).transpose(2, 0, 1, 3)).reshape(m, 8 * nc)
orc = _libs.oracle(nt=os.cpu_count() or 1)
ok = np.array_equal(lanes(out, 3), orc.scalar_mult(lanes(k, 1), lanes(P, 3)))
out.zero_(); dev.scalar_mult(out, k, P, n); torch.cuda.synchronize()
det = np.array_equal(dev.checksum(out), c1)
print({"lanes": n, "setup_s": round(t1 - t0, 1), "ladder_ms": round(ms, 1), "M_per_s": round(n / ms / 1e3, 2), "oracle_sample_ok": bool(ok), "deterministic": bool(det),
       "mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 1)})
