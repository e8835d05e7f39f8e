"""Steady-state target for ncu: the C2 fill launch (64 x FP32[3,224,224] = 38.5 MB) rotating over
four region sets (154 MB > the 126 MB L2), launched back to back exactly as a chain of device
passes launches it (tb200_fill_async, stream epoch in the kernel parameters).  By the time ncu's
--launch-skip has passed, every launch evicts the dirty lines an earlier launch left in L2, so
dram__bytes_write.sum per launch is the steady-state HBM write traffic, not a cold-cache artefact.

    ncu --replay-mode application --cache-control none --clock-control none \\
        -k regex:fill_uniform_kernel --launch-skip 240 --launch-count 8 ... python scripts/fill_steady.py
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from client_b200 import _native  # noqa: E402
from client_b200.device import DeviceBuffer, DeviceOps, make_fill_job  # noqa: E402


def main():
    launches = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dtype = sys.argv[2] if len(sys.argv) > 2 else "FP32"
    ctx = _native.Context(0)
    ops = DeviceOps(ctx)
    slot, slots, sets = 602112, 64, 4
    big = DeviceBuffer(0, sets * slots * slot)
    if dtype == "FP16":  # C3: one FP16[128,3,224,224] tensor per launch
        jobs = [(_native.FillJob * 1)(make_fill_job(big.ptr + s * slots * slot, slots * slot, "FP16", stream_id=s)) for s in range(sets)]
    else:
        jobs = [(_native.FillJob * slots)(*[make_fill_job(big.ptr + (s * slots + k) * slot, slot, "FP32", stream_id=k) for k in range(slots)])
                for s in range(sets)]
    for i in range(launches):
        ops.fill(jobs[i % sets], seed=1, epoch=i * slots)
    ops.sync()
    print("launched %d fills of %d bytes" % (launches, slots * slot))


if __name__ == "__main__":
    main()
