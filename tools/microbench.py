#!/usr/bin/env python3
"""Integer-pipe micro-benchmarks on one B200 (roofline denominators for DESIGN.md).
Prints one JSON line per probe: instructions/s and per-SM-per-clock rates."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ecsimd_b200  # noqa: E402
from ecsimd_b200 import device  # noqa: E402

NAMES = {0: "IMAD.WIDE.U32 (8 independent chains)", 1: "IMAD + IMAD.HI.U32 pairs", 2: "IADD3",
         3: "IMAD.WIDE.U32 + IADD3 1:1", 4: "IMAD.WIDE.U32.X carry chains (2x4)", 5: "VIMNMX3",
         6: "IMAD.WIDE.U32 + IADD3 1:2"}


def sm_clock():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        return float(out[0]), float(out[1])
    except Exception:
        return None, None


def main():
    ecsimd_b200.init(0)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    res = []
    for which in range(7):
        best = 0.0
        for warps in (4, 8, 16):
            threads = 256
            blocks = sms * warps * 32 // threads * 4
            rate, ms = device.microbench(which, blocks, threads, 2000)
            best = max(best, rate)
        clk, clkmax = sm_clock()
        rec = {"probe": NAMES[which], "inst_per_s": best, "thread_inst_per_clk_per_sm_at_max_clock":
               best / sms / (clkmax * 1e6) if clkmax else None, "sm_mhz_after": clk, "sm_max_mhz": clkmax}
        print(json.dumps(rec), flush=True)
        res.append(rec)
    return res


if __name__ == "__main__":
    main()
