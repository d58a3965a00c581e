#!/usr/bin/env python3
"""Pipe-mix probes: cycles per step per warp per SM sub-partition for mixes of IMAD.WIDE / IMAD /
IMAD.HI / IADD3 (development aid; results summarised in DESIGN.md)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ecsimd_b200  # noqa: E402
from ecsimd_b200 import capi  # noqa: E402


def main():
    ecsimd_b200.init(0)
    lib = capi.load()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    clk = 1.965e9
    for wps in (2, 4, 8, 16):          # warps per SM sub-partition
        for combo in range(lib.ecb200_microbench_mix_count()):
            threads = 256
            blocks = sms * wps * 4 * 32 // threads
            cnt = (C.c_int * 4)()
            ms = C.c_float()
            iters = 1000
            capi.check(lib.ecb200_microbench_mix(combo, blocks, threads, iters, cnt, C.byref(ms), None))
            # one wave: every SMSP holds `wps` warps for the whole kernel
            cyc_per_step_per_warp = ms.value * 1e-3 * clk / (iters * 8) / wps
            print(json.dumps({"warps_per_smsp": wps, "W": cnt[0], "L": cnt[1], "H": cnt[2], "Apairs": cnt[3],
                              "clk_per_step_per_warp": round(cyc_per_step_per_warp, 2), "ms": round(ms.value, 3)}), flush=True)


if __name__ == "__main__":
    main()
