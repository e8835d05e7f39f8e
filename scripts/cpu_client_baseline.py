"""Reference-style CPU client loop against a running server, for comparison with the
device-side load generator: what a tritonclient user does per request with
--shared-memory=cuda semantics (src/python/examples/simple_http_cudashm_client.py flow):
numpy generates the tensor, set_shared_memory_region copies it host->device (+sync), the
request names the region, get_contents_as_numpy reads the output back.  N worker threads.

    python scripts/cpu_client_baseline.py -u 127.0.0.1:8000 -m densenet_onnx --concurrency 8 --seconds 3
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(args, idx, stop, out, ready, go):
    import client_b200.http as httpclient
    import client_b200.utils.cuda_shared_memory as cudashm

    rng = np.random.default_rng(idx)
    client = httpclient.InferenceServerClient(args.url)
    in_h = cudashm.create_shared_memory_region("cpu_in%d" % idx, 602112, 0)
    out_h = cudashm.create_shared_memory_region("cpu_out%d" % idx, 4000, 0)
    client.register_cuda_shared_memory("cpu_in%d" % idx, cudashm.get_raw_handle(in_h), 0, 602112)
    client.register_cuda_shared_memory("cpu_out%d" % idx, cudashm.get_raw_handle(out_h), 0, 4000)
    inp = httpclient.InferInput("data_0", [3, 224, 224], "FP32").set_shared_memory("cpu_in%d" % idx, 602112)
    o = httpclient.InferRequestedOutput("fc6_1")
    o.set_shared_memory("cpu_out%d" % idx, 4000)
    lat = []
    ready.wait()
    go.wait()
    while not stop.is_set():
        t0 = time.perf_counter_ns()
        x = rng.random((3, 224, 224), dtype=np.float32)          # synthetic input on the host
        cudashm.set_shared_memory_region(in_h, [x])               # H2D + sync
        client.infer(args.model, [inp], outputs=[o])
        y = cudashm.get_contents_as_numpy(out_h, np.float32, [1000])  # D2H
        assert np.isfinite(y).all()
        lat.append(time.perf_counter_ns() - t0)
    client.unregister_cuda_shared_memory("cpu_in%d" % idx)
    client.unregister_cuda_shared_memory("cpu_out%d" % idx)
    cudashm.destroy_shared_memory_region(in_h)
    cudashm.destroy_shared_memory_region(out_h)
    client.close()
    out[idx] = lat


def run_threads(args, base, q=None, all_ready=None):
    stop, go = threading.Event(), threading.Event()
    ready = threading.Barrier(args.concurrency + 1)
    out = {}
    threads = [threading.Thread(target=worker, args=(args, base + i, stop, out, ready, go)) for i in range(args.concurrency)]
    for t in threads:
        t.start()
    ready.wait()           # regions created and registered, CUDA context up
    if all_ready is not None:
        all_ready.wait()   # ... in every client process
    t0 = time.perf_counter()
    go.set()
    time.sleep(args.seconds)
    stop.set()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    lat = np.concatenate([np.array(v, dtype=np.float64) for v in out.values()]) / 1e3
    if q is not None:
        q.put((lat, dt))
    return lat, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-u", "--url", default="127.0.0.1:8000")
    ap.add_argument("-m", "--model", default="densenet_onnx")
    ap.add_argument("--concurrency", type=int, default=8, help="worker threads per process")
    ap.add_argument("--processes", type=int, default=1, help="client processes (each its own interpreter and CUDA context)")
    ap.add_argument("--seconds", type=float, default=3.0)
    args = ap.parse_args()
    if args.processes <= 1:
        lat, dt = run_threads(args, 0)
    else:
        import multiprocessing as mp

        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        all_ready = ctx.Barrier(args.processes)
        procs = [ctx.Process(target=run_threads, args=(args, 1000 * (p + 1), q, all_ready)) for p in range(args.processes)]
        for p in procs:
            p.start()
        parts = [q.get() for _ in procs]
        for p in procs:
            p.join()
        lat = np.concatenate([a for a, _ in parts])
        dt = max(d for _, d in parts)
    print(json.dumps({"client": "reference-style CPU loop (numpy + set_shared_memory_region + get_contents_as_numpy)",
                      "processes": args.processes, "threads_per_process": args.concurrency,
                      "concurrency": args.concurrency * max(1, args.processes), "throughput": lat.size / dt,
                      "p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "count": int(lat.size)}))


if __name__ == "__main__":
    main()
