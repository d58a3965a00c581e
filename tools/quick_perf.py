#!/usr/bin/env python3
"""Quick device-resident timings (CUDA events) of the main kernels; development aid."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ecsimd_b200  # noqa: E402
from ecsimd_b200 import device as dev  # noqa: E402


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ecsimd_b200.init(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    k = dev.synth_values(dev.empty(n, 1), 0xEC51D004, 0, n, 0)
    a = dev.synth_values(dev.empty(n, 1), 0xEC51D001, 0, n, 1)
    b = dev.synth_values(dev.empty(n, 1), 0xEC51D002, 0, n, 1)
    out = dev.empty(n, 1)
    J = dev.empty(n, 3)
    # points: r*G on the device
    dev.scalar_mult_base(J, a, n)
    xy = dev.to_affine(dev.empty(n, 2), J, n)
    P = dev.from_affine(dev.empty(n, 3), xy, n)
    O = dev.empty(n, 3)
    res = {}
    for quirk in (True, False):
        ms = timeit(lambda: dev.scalar_mult(O, k, P, n, quirk=quirk))
        res["scalar_mult%s" % ("" if quirk else "_noquirk")] = {"ms": ms, "per_s": n / ms * 1e3, "TMAC32_per_s": n * 211540 / ms * 1e3 / 1e12}
    ms = timeit(lambda: dev.scalar_mult_base(O, k, n))
    res["scalar_mult_base"] = {"ms": ms, "per_s": n / ms * 1e3}
    ms = timeit(lambda: dev.mgry_mul(out, a, b, n), 10)
    res["mgry_mul_stream"] = {"ms": ms, "per_s": n / ms * 1e3, "GBps": n * 96 / ms * 1e3 / 1e9}
    ms = timeit(lambda: dev.mgry_add(out, a, b, n), 10)
    res["mgry_add_stream"] = {"ms": ms, "GBps": n * 96 / ms * 1e3 / 1e9}
    it = 1024
    ms = timeit(lambda: dev.mgry_mul_chain(out, a, b, it, n))
    res["mgry_mul_chain"] = {"ms": ms, "mulmod_per_s": n * it / ms * 1e3, "TMAC32_per_s": n * it * 64 / ms * 1e3 / 1e12}
    Q, R = dev.empty(n, 3), dev.empty(n, 3)
    dev.trplu(Q, R, P, n)
    ms = timeit(lambda: dev.zdau(O, J, R, Q, n))
    res["zdau_stream"] = {"ms": ms, "per_s": n / ms * 1e3, "GBps": n * 384 / ms * 1e3 / 1e9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
