"""resize_pack_kernel alone: 64 decoded 375x500 RGB images -> Pillow BILINEAR 224x224 -> FP32 CHW INCEPTION,
four rotating source / destination sets (> L2), CUDA-event time per launch.  Run it under ncu for the profile."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from client_b200 import _native  # noqa: E402
from client_b200.device import DeviceBuffer, DeviceOps, make_fill_job  # noqa: E402

ctx = _native.Context(0)
ops = DeviceOps(ctx)
N, H, W, SETS = 64, 375, 500, 4
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
src = [DeviceBuffer(0, N * H * W * 3) for _ in range(SETS)]
dst = [DeviceBuffer(0, N * 3 * 224 * 224 * 4) for _ in range(SETS)]
ops.fill([make_fill_job(b.ptr, N * H * W * 3, "UINT8", stream_id=10 + i) for i, b in enumerate(src)], seed=3)
ops.sync()
for dtype, layout, scaling in (("FP32", "NCHW", "INCEPTION"), ("FP16", "NCHW", "INCEPTION"), ("UINT8", "NHWC", "NONE")):
    ops.graph_begin()
    for s in range(SETS):
        ops.resize_pack_image(dst[s].ptr, dtype, layout, src[s].ptr, N, H, W, 3, 224, 224, scaling)
    g = ops.graph_end()
    for _ in range(3):
        g.launch()
    ops.sync()
    timer = _native.Timer(ctx)
    timer.start()
    for _ in range(reps):
        g.launch()
    timer.stop()
    ops.sync()
    us = timer.elapsed_ms() * 1e3 / (reps * SETS)
    es = _native.DTYPE_SIZES[dtype]
    nbytes = N * (H * W * 3 + 224 * 224 * 3 * es)
    print("%-5s %-4s %-9s %7.2f us/launch  %7.1f GB/s  frac %.3f" % (dtype, layout, scaling, us, nbytes / us / 1e3, nbytes / us / 1e3 / 6574.8))
