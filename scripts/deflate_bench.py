"""Device deflate encoder: throughput and ratio per kind of data, every stream inflated by zlib.
   python scripts/deflate_bench.py   (GPU box)"""
import os
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from client_b200 import _native  # noqa: E402
from client_b200.device import DeviceBuffer, DeviceOps, HostBuffer, make_fill_job  # noqa: E402


def main():
    ctx = _native.Context(0)
    ops = DeviceOps(ctx)
    n = 64 * 602112
    src = DeviceBuffer(0, n)
    cap = int(_native.load().tb200_deflate_bound(n))
    dst = DeviceBuffer(0, cap)
    res = HostBuffer(4096)
    timer = _native.Timer(ctx)
    rng = np.random.default_rng(0)
    zipf = (rng.zipf(1.3, n // 8) % 30522).astype(np.int64)
    text = np.frombuffer((b'the quick brown fox jumps over the lazy dog {"name":"INPUT0","shape":[1,16],"datatype":"INT32"} ' * (n // 97 + 1))[:n], np.uint8)
    cases = [("token ids INT64 [0,30522)", make_fill_job(src.ptr, n, "INT64", stream_id=1, low=0, high=30522)),
             ("token ids INT32 [0,128256)", make_fill_job(src.ptr, n, "INT32", stream_id=2, low=0, high=128256)),
             ("attention mask INT64 0/1", make_fill_job(src.ptr, n, "INT64", stream_id=3, low=0, high=2)),
             ("zero data", make_fill_job(src.ptr, n, "FP32", mode="zero")),
             ("FP32 unit interval", make_fill_job(src.ptr, n, "FP32", stream_id=4)),
             ("zipf token ids INT64 (host data)", zipf), ("repeated text (host data)", text)]
    print("%-36s %9s %8s %8s %10s %10s" % ("data", "us", "GB/s in", "ratio", "zlib6", "zlib6 MB/s"))
    for label, job in cases:
        if isinstance(job, np.ndarray):
            raw = job.view(np.uint8)
            ops.h2d(src.ptr, raw.ctypes.data, raw.size)
            ops.sync()
        else:
            ops.fill([job], seed=5)
            ops.sync()
        for _ in range(2):
            ops.deflate_async(dst.ptr, cap, src.ptr, n, res.device_ptr, "deflate")
        ops.sync()
        timer.start()
        for _ in range(5):
            ops.deflate_async(dst.ptr, cap, src.ptr, n, res.device_ptr, "deflate")
        timer.stop()
        ops.sync()
        ms = timer.elapsed_ms() / 5
        out_bytes = int(res.array(np.uint64, 1)[0])
        stream = ops.download(dst.ptr, out_bytes).tobytes()
        data = ops.download(src.ptr, n).tobytes()
        assert zlib.decompress(stream) == data, label
        import time
        sample = data[:4 << 20]
        t0 = time.perf_counter()
        z = zlib.compress(sample, 6)
        dt = time.perf_counter() - t0
        print("%-36s %9.1f %8.1f %8.4f %10.4f %10.1f" % (label, ms * 1e3, n / ms / 1e6, out_bytes / n, len(z) / len(sample), len(sample) / dt / 1e6), flush=True)


if __name__ == "__main__":
    main()
