#!/usr/bin/env python3
"""Single-process multi-GPU host batches (ecb200_init_devices): one process, host buffers in the reference's pack layout,
2^23 (k, P) pairs; the batch on one device against the batch cut over all visible devices.  Prints one JSON line.
usage: tools/multi_gpu_host.py [log2 lanes] [pinned|pageable]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ecsimd_b200
from ecsimd_b200 import capi, device as dev, host


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    kind = sys.argv[2] if len(sys.argv) > 2 else "pinned"
    n = 1 << log2n
    ndev = torch.cuda.device_count()
    ecsimd_b200.init(0)
    torch.cuda.set_device(0)
    # inputs generated on device 0, moved to host pack4 buffers
    k = dev.synth_values(dev.empty(n, 1), 0xEC51D004, 0, n, 0)
    r = dev.synth_values(dev.empty(n, 1), 0xEC51D003, 0, n, 0)
    J = dev.scalar_mult_base(dev.empty(n, 3), r, n)
    P = dev.from_affine(dev.empty(n, 3), dev.to_affine(dev.empty(n, 2), J, n), n)
    torch.cuda.synchronize()
    kp = host.lane_to_pack4(host.soa_to_lane(k.cpu().numpy().view(np.uint32), 1), 1)
    Pp = host.lane_to_pack4(host.soa_to_lane(P.cpu().numpy().view(np.uint32), 3), 3)
    del k, r, J, P
    torch.cuda.empty_cache()
    if kind == "pinned":
        hk = torch.from_numpy(kp.view(np.int32)).pin_memory(); hP = torch.from_numpy(Pp.view(np.int32)).pin_memory()
        hout = torch.zeros((n // 4, 96), dtype=torch.int32).pin_memory()
        ptr = lambda t: t.data_ptr()
        as_np = lambda t: t.numpy().view(np.uint32)
    else:
        hk, hP, hout = kp, Pp, np.zeros((n // 4, 96), np.uint32)
        ptr = capi._p
        as_np = lambda t: t
    flags = capi.LAYOUT_PACK4 | capi.MEM_HOST
    call = lambda: capi.call("ecb200_scalar_mult_p256", ptr(hout), ptr(hk), ptr(hP), n, flags, None)

    def timed(reps=2):
        call()
        t0 = time.perf_counter()
        for _ in range(reps):
            call()
        return (time.perf_counter() - t0) / reps
    t1 = timed()
    single = as_np(hout).copy()
    ecsimd_b200.init_devices(list(range(ndev)))
    tn = timed()
    same = bool(np.array_equal(single, as_np(hout)))
    ecsimd_b200.init_devices([])
    print(json.dumps({"lanes": n, "host_buffers": kind + " pack4", "devices": ndev, "one_device_per_s": n / t1, "all_devices_per_s": n / tn,
                      "speedup": t1 / tn, "bit_identical": same}), flush=True)


if __name__ == "__main__":
    main()
