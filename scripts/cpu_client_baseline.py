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


def worker_grpc(args, idx, stop, out, ready, go):
    """The reference flow for tensors on the wire (simple_grpc_infer_client.py): numpy generates
    the token ids, set_data_from_numpy copies them into the request, grpcio sends it, as_numpy
    reads the logits.  BASELINE configs[3] (bert_large) / configs[4] prompt shape (llama, unary
    is not possible for a decoupled model: --mode grpc-llama-stream uses one stream per worker)."""
    import queue

    import client_b200.grpc as grpcclient

    rng = np.random.default_rng(idx)
    client = grpcclient.InferenceServerClient(args.url)
    lat, first = [], []
    stream = args.mode == "grpc-llama-stream"
    if stream:
        done = queue.Queue()
        state = {}

        def on_response(result, error):
            now = time.perf_counter_ns()
            if error is not None:
                done.put(now)
                return
            state.setdefault("first", now)
            params = result.get_response().parameters
            if "triton_final_response" not in params or params["triton_final_response"].bool_param:
                done.put(now)

        client.start_stream(callback=on_response)
    ready.wait()
    go.wait()
    while not stop.is_set():
        t0 = time.perf_counter_ns()
        if stream:
            ids = rng.integers(0, 128256, (1, 4096), dtype=np.int32)
            state.clear()
            client.async_stream_infer("llama3_8b", [grpcclient.InferInput("input_ids", [1, 4096], "INT32").set_data_from_numpy(ids)],
                                      parameters={"max_tokens": args.max_tokens})
            t1 = done.get(timeout=60)
            first.append(state.get("first", t1) - t0)
            lat.append(t1 - t0)
            continue
        ids = rng.integers(0, 30522, (1, 384), dtype=np.int64)
        mask = rng.integers(0, 2, (1, 384), dtype=np.int64)
        ins = [grpcclient.InferInput("input_ids", [1, 384], "INT64").set_data_from_numpy(ids),
               grpcclient.InferInput("attention_mask", [1, 384], "INT64").set_data_from_numpy(mask)]
        y = client.infer("bert_large", ins).as_numpy("logits")
        assert np.isfinite(y).all()
        lat.append(time.perf_counter_ns() - t0)
    if stream:
        client.stop_stream()
    client.close()
    out[idx] = lat
    out[("first", idx)] = first


def run_threads(args, base, q=None, all_ready=None):
    stop, go = threading.Event(), threading.Event()
    ready = threading.Barrier(args.concurrency + 1)
    out = {}
    fn = worker if args.mode == "cudashm" else worker_grpc
    threads = [threading.Thread(target=fn, args=(args, base + i, stop, out, ready, go)) for i in range(args.concurrency)]
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
    lat = np.concatenate([np.array(v, dtype=np.float64) for k, v in out.items() if not isinstance(k, tuple)]) / 1e3
    first = [np.array(v, dtype=np.float64) for k, v in out.items() if isinstance(k, tuple)]
    first = (np.concatenate(first) / 1e3) if first else np.zeros(0)
    if q is not None:
        q.put((lat, dt, first))
    return lat, dt, first


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-u", "--url", default="127.0.0.1:8000")
    ap.add_argument("-m", "--model", default="densenet_onnx")
    ap.add_argument("--concurrency", type=int, default=8, help="worker threads per process")
    ap.add_argument("--processes", type=int, default=1, help="client processes (each its own interpreter and CUDA context)")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--mode", default="cudashm", choices=["cudashm", "grpc-bert", "grpc-llama-stream"],
                    help="cudashm: C2 flow over HTTP; grpc-bert: C4 tensors in the message; grpc-llama-stream: C5 prompts on a stream per worker")
    ap.add_argument("--max-tokens", type=int, default=16)
    args = ap.parse_args()
    if args.processes <= 1:
        lat, dt, first = run_threads(args, 0)
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
        lat = np.concatenate([a for a, _, _ in parts])
        dt = max(d for _, d, _ in parts)
        first = np.concatenate([f for _, _, f in parts])
    label = {"cudashm": "reference-style CPU loop (numpy + set_shared_memory_region + get_contents_as_numpy)",
             "grpc-bert": "reference-style CPU loop (numpy + set_data_from_numpy + grpcio infer + as_numpy), bert_large",
             "grpc-llama-stream": "reference-style CPU loop (numpy + set_data_from_numpy + async_stream_infer, one stream per worker), llama3_8b"}[args.mode]
    extra = {"ttft_p50_us": float(np.percentile(first, 50)), "tokens_per_s": lat.size * args.max_tokens / dt} if first.size else {}
    print(json.dumps({"client": label, **extra,
                      "processes": args.processes, "threads_per_process": args.concurrency,
                      "concurrency": args.concurrency * max(1, args.processes), "throughput": lat.size / dt,
                      "p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "count": int(lat.size)}))


if __name__ == "__main__":
    main()
