"""Why does the loopback rate per GPU drop when two instances run side by side?  On a 2-GPU box:
both generators at once vs one alone, and the same for 32 drop-in API clients per GPU (bench.py's e2e loop),
under bench.LoopbackBox (one MPS daemon per GPU).  Findings of the first version (one shared daemon): the
native generators do not disturb each other (2 x 452 k), but an MPS server takes 48 clients in all -- 17 of
2 x 32 API clients were refused ("device(s) busy or unavailable").
"""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run_pair(boxes, devs, label, seconds=1.0):
    out = {}

    def one(box, dev, arg):
        g = box.generator(arg, 200, 10, min_seconds=seconds)
        out[dev] = round(g["infer_per_s"] / 1e3, 1)

    ths = [threading.Thread(target=one, args=(boxes[d], d, a)) for d, a in devs]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    print(label, out, "sum", round(sum(out.values()), 1), flush=True)


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    nproc, secs = (4, 0.5) if quick else (32, 2.0)
    with bench.LoopbackBox([0, 1]) as box:  # one MPS daemon per GPU
        print("mps", box.mps, "ordinals", box.ordinal, flush=True)
        run_pair({0: box, 1: box}, [(0, 0), (1, 1)], "warm", 0.3)
        run_pair({0: box, 1: box}, [(0, 0), (1, 1)], "generators, both GPUs ", 0.3 if quick else 1.0)
        if not quick:
            run_pair({0: box}, [(0, 0)], "generator, GPU0 only   ")
        hosts(box, [0, 1], "drop-in API device loops, both GPUs", nproc, secs)
        if not quick:
            hosts(box, [0], "drop-in API device loops, GPU0 only", nproc, secs)


def hosts(box, devs, label, nproc=32, seconds=2.0):
    """`nproc` drop-in API clients per GPU in device mode (bench.py's e2e), all GPUs at once"""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    total = nproc * len(devs)
    ready, go, q = ctx.Barrier(total + 1), ctx.Event(), ctx.Queue()
    procs = []
    saved = dict(os.environ)
    try:
        for d in devs:
            os.environ.clear()
            os.environ.update(box.envs[d])
            for i in range(nproc):
                p = ctx.Process(target=bench._host_loop_worker, args=("b200", box.urls[d], box.ordinal[d], "d%d_%d_%d" % (d, os.getpid(), i), seconds, "device", ready, go, q))
                p.start()
                procs.append(p)
    finally:
        os.environ.clear()
        os.environ.update(saved)
    try:
        ready.wait(timeout=240)
    except Exception:
        pass
    go.set()
    parts = [q.get(timeout=seconds + 240) for _ in procs]
    for p in procs:
        p.join(30)
    errs = [e for _, _, e in parts if e]
    print(label, round(sum(n for n, _, _ in parts) / seconds / 1e3, 1), "k infer/s", ("errors: %d, first: %s" % (len(errs), errs[0][:200])) if errs else "", flush=True)


if __name__ == "__main__":
    main()
