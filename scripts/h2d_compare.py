"""set_shared_memory_region (602,112 B, the C2 tensor) under contention: the drop-in against the
restated reference flow (oracle/ref_client.py), P processes each looping for a fixed time, with
and without a larger direct-copy threshold.   python scripts/h2d_compare.py [P ...]"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(impl, direct_kb, seconds, barrier, q):
    x = np.random.default_rng(os.getpid()).random((3, 224, 224), dtype=np.float32)
    if impl == "ref":
        from oracle import ref_client

        r = ref_client.Region("x", 602112, 0)
        from cuda.bindings import runtime as cudart

        stream = ref_client._cuda(cudart.cudaStreamCreate)

        def put():
            flat = np.ascontiguousarray(x).flatten()
            ref_client._cuda(cudart.cudaMemcpyAsync, r.ptr, flat.ctypes.data, flat.size * 4, cudart.cudaMemcpyKind.cudaMemcpyDefault, stream)
            ref_client._cuda(cudart.cudaStreamSynchronize, stream)
    else:
        from client_b200 import _native
        import client_b200.utils.cuda_shared_memory as cudashm

        if direct_kb:
            _native.check(_native.load().tb200_tune(b"h2d_direct_kb", direct_kb))
        h = cudashm.create_shared_memory_region("x%d" % os.getpid(), 602112, 0)

        def put():
            cudashm.set_shared_memory_region(h, [x])
    for _ in range(20):
        put()
    barrier.wait()
    n, t_end = 0, time.perf_counter() + seconds
    t0 = time.perf_counter()
    while time.perf_counter() < t_end:
        put()
        n += 1
    q.put((n, time.perf_counter() - t0))


def main():
    procs = [int(a) for a in sys.argv[1:]] or [1, 8, 32]
    ctx = mp.get_context("spawn")
    for p in procs:
        for impl, kb in (("ref", 0), ("b200", 0), ("b200", 2048)):
            barrier, q = ctx.Barrier(p), ctx.Queue()
            ps = [ctx.Process(target=worker, args=(impl, kb, 2.0, barrier, q)) for _ in range(p)]
            for x in ps:
                x.start()
            parts = [q.get(timeout=120) for _ in ps]
            for x in ps:
                x.join()
            n = sum(a for a, _ in parts)
            dt = max(b for _, b in parts)
            print("P=%-3d %-5s direct_kb=%-5d %9.0f calls/s  %7.1f us/call/process  %6.2f GB/s" % (p, impl, kb, n / dt, dt / (n / p) * 1e6, n * 602112 / dt / 1e9), flush=True)


if __name__ == "__main__":
    main()
