"""perf_analyzer-style command line for the device-side load generator.

perf_analyzer is not part of the reference: the flags below
restate its publicly documented surface; behaviour parity with the real tool is
unpinned.  Example (BASELINE config C2):

    python -m client_b200.perf -m densenet_onnx -u 127.0.0.1:8000 -i http \\
        --shared-memory cuda --concurrency-range 1:64:2x -p 1000

Extra flags: ``--input-data-mode per-request|once`` (regenerate every request's input on
the device, or fill once like perf_analyzer does), ``--gpus N`` (one instance per GPU,
independent replicas), ``--json`` (one JSON line per level).
"""

import argparse
import json
import multiprocessing as mp
import sys

from .loadgen import ConcurrencyManager, SlotSet, TensorSpec, measure


def parse_range(text):
    """'start[:end[:step]]'; a step ending in 'x' multiplies (1:64:2x -> 1,2,4,...,64)."""
    parts = text.split(":")
    start = int(parts[0])
    end = int(parts[1]) if len(parts) > 1 else start
    step = parts[2] if len(parts) > 2 else "1"
    levels, c = [], start
    while c <= end:
        levels.append(c)
        c = c * int(step[:-1]) if step.endswith("x") else c + int(step)
    return levels


def build_parser():
    ap = argparse.ArgumentParser(prog="client_b200.perf", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-m", "--model-name", required=True)
    ap.add_argument("-x", "--model-version", default="")
    ap.add_argument("-u", "--url", default=None, help="host:port (default localhost:8000 http / :8001 grpc)")
    ap.add_argument("-i", "--protocol", default="http", choices=["http", "grpc", "HTTP", "gRPC", "GRPC"])
    ap.add_argument("-b", "--batch-size", type=int, default=1)
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--concurrency-range", default="1", help="start:end:step")
    ap.add_argument("--shared-memory", default="none", choices=["none", "system", "cuda"])
    ap.add_argument("--output-shared-memory-size", type=int, default=0, help="bytes per output when the shape is dynamic")
    ap.add_argument("--input-data", default="random", choices=["random", "zero"])
    ap.add_argument("--input-data-mode", default="per-request", choices=["per-request", "once"])
    ap.add_argument("--shape", action="append", default=[], help="name:d1,d2,... for dynamic inputs")
    ap.add_argument("-p", "--measurement-interval", type=int, default=1000, help="window in ms")
    ap.add_argument("-s", "--stability-percentage", type=float, default=10.0)
    ap.add_argument("-r", "--max-trials", type=int, default=10)
    ap.add_argument("--percentile", type=int, default=None)
    ap.add_argument("-f", "--filename", default=None, help="CSV report")
    ap.add_argument("--streaming", action="store_true", help="gRPC bidirectional stream (decoupled models)")
    ap.add_argument("-a", "--async", dest="async_mode", action="store_true")
    ap.add_argument("--sync", action="store_true")
    ap.add_argument("--request-parameter", action="append", default=[], help="name:value:type")
    ap.add_argument("--random-seed", type=int, default=0)
    ap.add_argument("--device-id", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-validate", action="store_true")
    ap.add_argument("--engine", default="python", choices=["python", "native"],
                    help="python: threads over the drop-in clients; native: libtb200's C++ workers (HTTP, sync)")
    ap.add_argument("--device-window-us", type=int, default=0,
                    help="native engine: wait up to this long for more returned slots before a device pass "
                         "(~150 helps when client and server are time-sliced CUDA contexts, i.e. no MPS)")
    ap.add_argument("--json", action="store_true")
    return ap


def _client_factory(protocol, url, verbose):
    if protocol == "grpc":
        from .. import grpc as mod
    else:
        from .. import http as mod
    return lambda: mod.InferenceServerClient(url, verbose=verbose)


def _specs(args, client, protocol):
    shapes = {}
    for s in args.shape:
        name, dims = s.rsplit(":", 1)
        shapes[name] = [int(d) for d in dims.split(",")]
    if protocol == "grpc":
        md = client.get_model_metadata(args.model_name, args.model_version, as_json=True)
    else:
        md = client.get_model_metadata(args.model_name, args.model_version)

    def fix(t, is_output):
        shape = [int(d) for d in t.get("shape", [])]
        if t["name"] in shapes:
            shape = shapes[t["name"]]
        if any(d < 0 for d in shape):
            if is_output and args.output_shared_memory_size:
                es = TensorSpec(t["name"], t["datatype"], [1]).nbytes
                shape = [args.output_shared_memory_size // es]
            elif is_output:
                shape = [1 if d < 0 else d for d in shape]
            else:
                raise SystemExit("input '%s' has a dynamic shape %s: pass --shape %s:d1,d2,..." % (t["name"], shape, t["name"]))
        return TensorSpec(t["name"], t["datatype"], shape)

    return [fix(t, False) for t in md["inputs"]], [fix(t, True) for t in md["outputs"]]


def run_instance(args, device_id, prefix, out_queue=None, staging_factory=None):
    protocol = args.protocol.lower()
    url = args.url or ("localhost:8001" if protocol == "grpc" else "localhost:8000")
    if "," in url:  # one server per GPU instance: -u host:p0,host:p1,...
        urls = [u.strip() for u in url.split(",") if u.strip()]
        url = urls[device_id % len(urls)]
    make_client = _client_factory(protocol, url, args.verbose)
    control = make_client()
    inputs, outputs = _specs(args, control, protocol)
    params = {}
    for p in args.request_parameter:
        name, value, typ = p.split(":")
        params[name] = {"int": int, "bool": lambda v: v.lower() == "true", "string": str, "float": float}[typ](value)
    token_range = {"input_ids": (0, 30522) if "bert" in args.model_name else (0, 128256), "attention_mask": (0, 2)}
    rows = []
    for level in parse_range(args.concurrency_range):
        slotset = SlotSet(inputs, outputs, level, args.shared_memory, device_id, args.input_data,
                          args.random_seed + 1000003 * device_id, token_range, name_prefix="%s_c%d" % (prefix, level),
                          staging=staging_factory() if staging_factory else None)
        slotset.register(control)
        if args.engine == "native":
            if protocol != "http" or args.streaming:
                raise SystemExit("--engine native drives HTTP synchronous requests only")
            from .native import NativeLoadGenerator

            gen = NativeLoadGenerator(url, args.model_name, args.model_version, slotset, level,
                                      regenerate=args.input_data_mode == "per-request", validate=not args.no_validate,
                                      device_window_us=args.device_window_us)
            gen.start()
            try:
                res = measure_native(gen, args.measurement_interval, args.stability_percentage, args.max_trials, args.percentile)
            finally:
                gen.stop()
                slotset.unregister(control)
                slotset.close()
            res.update(concurrency=level, device=device_id, errors=[], input_bytes=slotset.in_bytes, engine="native",
                       launches_per_request=(2.0 * res["device_batches"] / max(1, res["device_slots"])) if res.get("device_batches") else 0.0)
            rows.append(res)
            if out_queue is None:
                _report(args, res)
            continue
        mgr = ConcurrencyManager(make_client, protocol, args.model_name, args.model_version, slotset, level,
                                 per_request_data=args.input_data_mode == "per-request", validate=not args.no_validate,
                                 streaming=args.streaming, request_parameters=params or None)
        mgr.start()
        try:
            res = measure(mgr, args.measurement_interval, args.stability_percentage, args.max_trials, args.percentile)
        finally:
            mgr.stop()
            slotset.unregister(control)
            slotset.close()
        res.update(concurrency=level, device=device_id, nonfinite=mgr.nonfinite, errors=mgr.errors,
                   launches_per_request=(2.0 * len(mgr.device_batches) / max(1, sum(mgr.device_batches))) if mgr.device_batches else 0.0,
                   input_bytes=slotset.in_bytes)
        rows.append(res)
        if out_queue is None:
            _report(args, res)
    control.close()
    if out_queue is not None:
        out_queue.put(rows)
    return rows


def measure_native(gen, interval_ms, stability_pct, max_trials, percentile, min_windows=3):
    """Same window / stability rule as loadgen.measure, over the native generator."""
    gen.window(0.2)  # discard the ramp-up
    windows = []
    for _ in range(max_trials):
        w = gen.window(interval_ms / 1e3)
        w["latency_us"] = w["p%d_us" % percentile] if percentile in (50, 90, 95, 99) else w["avg_us"]
        windows.append(w)
        if len(windows) >= min_windows:
            last = windows[-min_windows:]
            thr = [x["throughput"] for x in last]
            lat = [x["latency_us"] for x in last]
            if min(thr) > 0 and (max(thr) - min(thr)) / max(thr) <= stability_pct / 100.0 and \
                    (max(lat) == 0 or (max(lat) - min(lat)) / max(lat) <= stability_pct / 100.0):
                break
    last = windows[-min_windows:] if len(windows) >= min_windows else windows
    merged = dict(last[-1])
    merged["throughput"] = sum(x["throughput"] for x in last) / len(last)
    merged["count"] = sum(x["count"] for x in last)
    merged["failed"] = sum(x["failed"] for x in last)
    merged["nonfinite"] = sum(x["nonfinite"] for x in last)
    merged["device_batches"] = sum(x["device_batches"] for x in last)
    merged["device_slots"] = sum(x["device_slots"] for x in last)
    merged["windows"] = len(windows)
    return merged


def _report(args, res):
    if args.json:
        print(json.dumps(res), flush=True)
        return
    lat = res.get("latency_us", 0.0)
    print("Concurrency: %d, throughput: %.2f infer/sec, latency %.0f usec" % (res["concurrency"], res["throughput"], lat))
    if "p50_us" in res:
        print("    p50: %.0f usec  p90: %.0f  p95: %.0f  p99: %.0f  (windows %d, failed %d)" % (
            res["p50_us"], res["p90_us"], res["p95_us"], res["p99_us"], res["windows"], res["failed"]))
    if "ttft_p50_us" in res:
        print("    time to first response p50: %.0f usec  p99: %.0f" % (res["ttft_p50_us"], res["ttft_p99_us"]))
    if res.get("errors"):
        print("    errors: %s" % res["errors"][:2])
    sys.stdout.flush()


def main(argv=None, staging_factory=None):
    args = build_parser().parse_args(argv)
    if args.gpus <= 1:
        rows = run_instance(args, args.device_id, "tb200_d%d" % args.device_id, staging_factory=staging_factory)
    else:
        # one independent instance per GPU (no collective); levels are summed across instances
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=run_instance, args=(args, d, "tb200_d%d" % d, q)) for d in range(args.gpus)]
        for p in procs:
            p.start()
        per_gpu = [q.get() for _ in procs]
        for p in procs:
            p.join()
        rows = []
        for level_rows in zip(*per_gpu):
            agg = dict(level_rows[0])
            agg["throughput"] = sum(r["throughput"] for r in level_rows)
            agg["count"] = sum(r["count"] for r in level_rows)
            agg["gpus"] = args.gpus
            rows.append(agg)
            _report(args, agg)
    if args.filename:
        with open(args.filename, "w") as fh:
            fh.write("Concurrency,Inferences/Second,p50 latency,p90 latency,p95 latency,p99 latency,Avg latency\n")
            for r in rows:
                fh.write("%d,%.2f,%.0f,%.0f,%.0f,%.0f,%.0f\n" % (r["concurrency"], r["throughput"], r.get("p50_us", 0),
                                                                r.get("p90_us", 0), r.get("p95_us", 0), r.get("p99_us", 0), r.get("avg_us", 0)))
    return rows


if __name__ == "__main__":
    main()
