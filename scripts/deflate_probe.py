"""Device deflate on three kinds of 38.5 MB bodies (run under ncu for per-kernel times)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from client_b200 import _native  # noqa: E402
from client_b200.device import DeviceBuffer, DeviceOps, HostBuffer, make_fill_job  # noqa: E402

ops = DeviceOps(_native.Context(0))
n = 38535168
src = DeviceBuffer(0, n)
cap = int(_native.load().tb200_deflate_bound(n))
dst = DeviceBuffer(0, cap)
size = HostBuffer(64)
for label, job in (("ids", make_fill_job(src.ptr, n, "INT64", stream_id=1, low=0, high=30522)),
                   ("zero", make_fill_job(src.ptr, n, "FP32", mode="zero")),
                   ("fp32", make_fill_job(src.ptr, n, "FP32", stream_id=2))):
    ops.fill([job], seed=1)
    ops.sync()
    for _ in range(2):
        ops.deflate_async(dst.ptr, cap, src.ptr, n, size.device_ptr, "gzip")
    ops.sync()
    print(label, int(size.array(np.uint64, 1)[0]))
