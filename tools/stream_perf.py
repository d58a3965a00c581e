#!/usr/bin/env python3
"""Streaming field kernels at a size that does not fit L2 (2^24 lanes = 1.5 GiB of traffic):
achieved HBM GB/s per kernel and layout (CUDA events).  Development aid / profiles input."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ecsimd_b200  # noqa: E402
from ecsimd_b200 import capi, device as dev  # noqa: E402


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ecsimd_b200.init(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
    res = {}
    for layout in ("soa", "lane", "pack4"):
        a = dev.synth_values(dev.empty(n, 1, layout), 0xEC51D001, 0, n, 1, layout)
        b = dev.synth_values(dev.empty(n, 1, layout), 0xEC51D002, 0, n, 1, layout)
        o = dev.empty(n, 1, layout)
        for name, fn, nbytes in (("mgry_mul", lambda: dev.mgry_mul(o, a, b, n, layout), 96), ("mgry_add", lambda: dev.mgry_add(o, a, b, n, layout), 96),
                                 ("mgry_sub", lambda: dev.mgry_sub(o, a, b, n, layout), 96), ("mgry_sqr", lambda: dev.mgry_sqr(o, a, n, layout), 64)):
            ms = timeit(fn)
            res["%s/%s" % (name, layout)] = {"ms": round(ms, 4), "GBps": round(n * nbytes / ms * 1e3 / 1e9, 1), "lanes_per_s": n / ms * 1e3}
        del a, b, o
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
