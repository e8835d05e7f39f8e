"""One deflate launch of 38.5 MB of token ids (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from client_b200 import _native
from client_b200.device import DeviceBuffer, DeviceOps, HostBuffer, make_fill_job
ctx = _native.Context(0); ops = DeviceOps(ctx)
n = 64 * 602112
src = DeviceBuffer(0, n); cap = int(_native.load().tb200_deflate_bound(n)); dst = DeviceBuffer(0, cap); res = HostBuffer(4096)
kind = sys.argv[1] if len(sys.argv) > 1 else "tokens"
job = make_fill_job(src.ptr, n, "INT64", stream_id=1, low=0, high=30522) if kind == "tokens" else make_fill_job(src.ptr, n, "FP32", mode="zero")
ops.fill([job], seed=5); ops.sync()
for _ in range(3):
    ops.deflate_async(dst.ptr, cap, src.ptr, n, res.device_ptr, "deflate")
ops.sync()
