"""set_shared_memory_region / get_contents_as_numpy: the drop-in (libtb200 staging) next
to the reference's flow restated with cuda-python (oracle/cudashm_ref.py), same GPU,
same host arrays.  python scripts/cudashm_compare.py > gpurun_out/cudashm_compare.txt"""

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import client_b200.utils.cuda_shared_memory as cudashm  # noqa: E402
from oracle.cudashm_ref import RefRegion  # noqa: E402

CASES = [("C1 int32[1,16]", np.int32, (1, 16)), ("C5 int32[1,4096]", np.int32, (1, 4096)),
         ("C2 fp32[3,224,224]", np.float32, (3, 224, 224)), ("C3 fp16[128,3,224,224]", np.float16, (128, 3, 224, 224))]


def bench(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    rng = np.random.default_rng(0)
    print("%-26s %12s | %14s %14s %8s | %14s %14s %8s" % ("tensor", "bytes", "ref set us", "ours set us", "speedup", "ref get us", "ours get us", "speedup"))
    for name, dt, shape in CASES:
        x = (rng.random(shape) * 100).astype(dt)
        n = x.nbytes
        reps = 2000 if n < 1 << 20 else 30
        ref = RefRegion(n)
        ours = cudashm.create_shared_memory_region("cmp", n, 0)
        t_ref_set = bench(lambda: ref.set([x]), reps)
        t_our_set = bench(lambda: cudashm.set_shared_memory_region(ours, [x]), reps)
        a = ref.get(dt, shape)
        b = cudashm.get_contents_as_numpy(ours, dt, shape)
        assert np.array_equal(a, x) and np.array_equal(b, x)
        t_ref_get = bench(lambda: ref.get(dt, shape), reps)
        t_our_get = bench(lambda: cudashm.get_contents_as_numpy(ours, dt, shape), reps)
        print("%-26s %12d | %14.1f %14.1f %7.2fx | %14.1f %14.1f %7.2fx   (set: %.1f vs %.1f GB/s)" % (
            name, n, t_ref_set * 1e6, t_our_set * 1e6, t_ref_set / t_our_set, t_ref_get * 1e6, t_our_get * 1e6,
            t_ref_get / t_our_get, n / t_ref_set / 1e9, n / t_our_set / 1e9))
        ref.close()
        cudashm.destroy_shared_memory_region(ours)
    # output region of 4,000 bytes requested from a 256 KB region: the reference copies the whole region
    big = 64 * 4000
    ref = RefRegion(big)
    ours = cudashm.create_shared_memory_region("cmp2", big, 0)
    t_ref = bench(lambda: ref.get(np.float32, (1000,)), 2000)
    t_our = bench(lambda: cudashm.get_contents_as_numpy(ours, np.float32, (1000,)), 2000)
    print("get 1000 x fp32 out of a 256,000-byte region: ref %.1f us, ours %.1f us (%.2fx)" % (t_ref * 1e6, t_our * 1e6, t_ref / t_our))


if __name__ == "__main__":
    main()
