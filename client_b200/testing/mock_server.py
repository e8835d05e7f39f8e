"""A minimal KServe-v2 / Triton-protocol server for tests and local load generation.

The reference repository contains no server and no fake:
its end-to-end coverage assumes a live Triton.  This module is the stand-in: it
runs in its OWN PROCESS (CUDA IPC handles can only be opened by another process,
reference README.md:196-204), speaks the HTTP/REST and gRPC inference protocols
the client emits, and implements the shared-memory control plane (system + CUDA).

Models (inputs -> outputs):
  simple                 INPUT0, INPUT1 INT32[1,16] -> OUTPUT0 = sum, OUTPUT1 = difference
                         (src/python/examples/simple_http_infer_client.py:242-263)
  custom_identity_int32  INPUT0 INT32[-1] -> OUTPUT0 (memory_growth_test.py)
  identity_*             any single input -> OUTPUT0 with the same bytes
  string_identity        INPUT0 BYTES[1,8] -> OUTPUT0 (generated string inputs)
  densenet_onnx          data_0 FP32[3,224,224] -> fc6_1 FP32[1000]: mean of the input
                         elements i with i % 1000 == j (deterministic stand-in)
  bert_large             input_ids, attention_mask INT64[1,384] -> logits FP32[1,384]
  repeat_int32           decoupled: IN INT32[n] -> n responses OUT INT32[1]
                         (simple_grpc_custom_repeat.py)
  llama3_8b              decoupled: input_ids INT32[1,L] -> `max_tokens` (default 4)
                         responses token INT32[1,1]

Test infrastructure / tooling: not part of the client data plane.
"""

import argparse
import base64
import ctypes
import json
import sys
import threading
import time
from concurrent import futures
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

import numpy as np

from .. import utils
from ..utils import shared_memory as sysshm

_NP = {
    "BOOL": np.bool_, "INT8": np.int8, "INT16": np.int16, "INT32": np.int32, "INT64": np.int64,
    "UINT8": np.uint8, "UINT16": np.uint16, "UINT32": np.uint32, "UINT64": np.uint64,
    "FP16": np.float16, "FP32": np.float32, "FP64": np.float64,
}


class ServerError(Exception):
    def __init__(self, msg, status=400):
        super().__init__(msg)
        self.status = status


class _CudaRegions:
    """CUDA shared memory opened from the client's IPC handle through libtb200."""

    def __init__(self):
        self._regions = {}
        self._ctx = {}
        self._lock = threading.Lock()
        self._io_lock = threading.Lock()

    def _native(self):
        from .. import _native

        return _native

    def register(self, name, raw_handle, device_id, byte_size):
        nat = self._native()
        lib = nat.load()
        with self._lock:
            if name in self._regions:
                raise ServerError("shared memory region '%s' already in manager" % name)
            raw = (ctypes.c_uint8 * 64).from_buffer_copy(raw_handle)
            h = ctypes.c_void_p()
            rc = lib.tb200_region_open(raw, int(byte_size), int(device_id), ctypes.byref(h))
            if rc != 0:
                raise ServerError("failed to open CUDA IPC handle: " + nat.last_error())
            if device_id not in self._ctx:
                self._ctx[device_id] = nat.Context(device_id)
            self._regions[name] = (h, int(device_id), int(byte_size))

    def unregister(self, name=""):
        lib = self._native().load() if self._regions else None
        with self._lock:
            names = [name] if name else list(self._regions)
            for n in names:
                entry = self._regions.pop(n, None)
                if entry is not None:
                    lib.tb200_region_destroy(entry[0])

    def status(self, name=""):
        with self._lock:
            items = {n: v for n, v in self._regions.items() if not name or n == name}
        if name and not items:
            raise ServerError("Unable to find cuda shared memory region: '%s'" % name)
        return [{"name": n, "device_id": v[1], "byte_size": v[2]} for n, v in items.items()]

    def has(self, name):
        return name in self._regions

    def read(self, name, offset, nbytes):
        nat = self._native()
        h, dev, size = self._regions[name]
        if offset + nbytes > size:
            raise ServerError("shared memory region '%s' is too small" % name)
        out = np.empty(nbytes, np.uint8)
        with self._io_lock:  # a context (stream + staging ring) serves one thread at a time
            nat.check(nat.load().tb200_region_read_host(self._ctx[dev].handle, h, offset, out.ctypes.data, nbytes))
        return out.tobytes()

    def write(self, name, offset, data):
        nat = self._native()
        h, dev, size = self._regions[name]
        if offset + len(data) > size:
            raise ServerError("shared memory region '%s' is too small for the output" % name)
        buf = np.frombuffer(data, np.uint8)
        with self._io_lock:
            nat.check(nat.load().tb200_region_write_host(self._ctx[dev].handle, h, offset, buf.ctypes.data, len(data)))


class _SystemRegions:
    def __init__(self):
        self._regions = {}
        self._lock = threading.Lock()

    def register(self, name, key, offset, byte_size):
        with self._lock:
            if name in self._regions:
                raise ServerError("shared memory region '%s' already in manager" % name)
            try:
                region = sysshm._PosixRegion(key)
            except FileNotFoundError:
                raise ServerError("Unable to open shared memory region: '%s'" % key)
            self._regions[name] = (region, key, int(offset), int(byte_size))

    def unregister(self, name=""):
        with self._lock:
            for n in ([name] if name else list(self._regions)):
                entry = self._regions.pop(n, None)
                if entry is not None:
                    entry[0].close()

    def status(self, name=""):
        with self._lock:
            items = {n: v for n, v in self._regions.items() if not name or n == name}
        if name and not items:
            raise ServerError("Unable to find system shared memory region: '%s'" % name)
        return [{"name": n, "key": v[1], "offset": v[2], "byte_size": v[3]} for n, v in items.items()]

    def has(self, name):
        return name in self._regions

    def read(self, name, offset, nbytes):
        region, _, base, size = self._regions[name]
        if offset + nbytes > size:
            raise ServerError("shared memory region '%s' is too small" % name)
        return bytes(region.buf[base + offset: base + offset + nbytes])

    def write(self, name, offset, data):
        region, _, base, size = self._regions[name]
        if offset + len(data) > size:
            raise ServerError("shared memory region '%s' is too small for the output" % name)
        region.buf[base + offset: base + offset + len(data)] = data


def _decode(datatype, shape, raw):
    if datatype == "BYTES":
        return utils.deserialize_bytes_tensor(raw).reshape(shape)
    if datatype == "BF16":
        return utils.deserialize_bf16_tensor(raw).reshape(shape)
    return np.frombuffer(raw, dtype=_NP[datatype]).reshape(shape)


def _encode(datatype, arr):
    if datatype == "BYTES":
        ser = utils.serialize_byte_tensor(arr)
        return ser.item() if ser.size else b""
    if datatype == "BF16":
        ser = utils.serialize_bf16_tensor(np.ascontiguousarray(arr, dtype=np.float32))
        return ser.item() if ser.size else b""
    return np.ascontiguousarray(arr).tobytes()


# model -> (platform, inputs [(name, dtype, shape)], outputs, decoupled)
MODELS = {
    "simple": ("mock", [("INPUT0", "INT32", [1, 16]), ("INPUT1", "INT32", [1, 16])],
               [("OUTPUT0", "INT32", [1, 16]), ("OUTPUT1", "INT32", [1, 16])], False),
    "custom_identity_int32": ("mock", [("INPUT0", "INT32", [-1])], [("OUTPUT0", "INT32", [-1])], False),
    "densenet_onnx": ("onnxruntime_onnx", [("data_0", "FP32", [3, 224, 224])], [("fc6_1", "FP32", [1000])], False),
    "bert_large": ("mock", [("input_ids", "INT64", [1, 384]), ("attention_mask", "INT64", [1, 384])],
                   [("logits", "FP32", [1, 384])], False),
    "string_identity": ("mock", [("INPUT0", "BYTES", [1, 8])], [("OUTPUT0", "BYTES", [1, 8])], False),
    # models of the reference's example scripts (src/python/examples/simple_*_string_*, *_int8_*, *_sequence_*)
    "simple_string": ("mock", [("INPUT0", "BYTES", [1, 16]), ("INPUT1", "BYTES", [1, 16])],
                      [("OUTPUT0", "BYTES", [1, 16]), ("OUTPUT1", "BYTES", [1, 16])], False),
    "simple_int8": ("mock", [("INPUT0", "INT8", [1, 16]), ("INPUT1", "INT8", [1, 16])],
                    [("OUTPUT0", "INT8", [1, 16]), ("OUTPUT1", "INT8", [1, 16])], False),
    "simple_sequence": ("mock", [("INPUT", "INT32", [1, 1])], [("OUTPUT", "INT32", [1, 1])], False),
    "simple_identity": ("mock", [("INPUT0", "BYTES", [-1, -1])], [("OUTPUT0", "BYTES", [-1, -1])], False),
    "repeat_int32": ("mock", [("IN", "INT32", [-1])], [("OUT", "INT32", [1])], True),
    "llama3_8b": ("mock", [("input_ids", "INT32", [1, -1])], [("token", "INT32", [1, 1])], True),
}


def run_model(model, inputs, params):
    """inputs: {name: ndarray} -> list of {name: (datatype, ndarray)} (one per response)."""
    if model == "simple":
        a, b = inputs["INPUT0"], inputs["INPUT1"]
        return [{"OUTPUT0": ("INT32", a + b), "OUTPUT1": ("INT32", a - b)}]
    if model == "simple_int8":
        a, b = inputs["INPUT0"].astype(np.int8), inputs["INPUT1"].astype(np.int8)
        return [{"OUTPUT0": ("INT8", (a + b).astype(np.int8)), "OUTPUT1": ("INT8", (a - b).astype(np.int8))}]
    if model == "simple_string":  # decimal strings in, decimal strings out
        a = np.array([int(x) for x in inputs["INPUT0"].reshape(-1)], dtype=np.int64)
        b = np.array([int(x) for x in inputs["INPUT1"].reshape(-1)], dtype=np.int64)
        shape = inputs["INPUT0"].shape
        enc = lambda v: np.array([str(int(x)).encode() for x in v], dtype=object).reshape(shape)  # noqa: E731
        return [{"OUTPUT0": ("BYTES", enc(a + b)), "OUTPUT1": ("BYTES", enc(a - b))}]
    if model == "simple_sequence":
        # the value checked by simple_*_sequence_*_client.py: the input, plus one on the
        # request that starts a sequence
        x = inputs["INPUT"].astype(np.int32)
        return [{"OUTPUT": ("INT32", x + (1 if params.get("sequence_start") else 0))}]
    if model == "densenet_onnx":
        x = inputs["data_0"].astype(np.float32).reshape(-1)
        pad = (-x.size) % 1000
        xs = np.concatenate([x, np.zeros(pad, np.float32)]).reshape(-1, 1000)
        return [{"fc6_1": ("FP32", xs.sum(axis=0, dtype=np.float32) / np.float32(xs.shape[0]))}]
    if model == "bert_large":
        ids, mask = inputs["input_ids"], inputs["attention_mask"]
        return [{"logits": ("FP32", ((ids % 1000).astype(np.float32) * mask.astype(np.float32)) / np.float32(1000))}]
    if model == "repeat_int32":
        vals = inputs["IN"].reshape(-1)
        return [{"OUT": ("INT32", np.array([v], np.int32))} for v in vals]
    if model == "llama3_8b":
        ids = inputs["input_ids"].reshape(-1)
        n = int(params.get("max_tokens", 4))
        base = int(ids.astype(np.int64).sum() % 128256)
        return [{"token": ("INT32", np.array([[(base + k) % 128256]], np.int32))} for k in range(n)]
    if model.startswith("identity") or model in ("custom_identity_int32", "string_identity", "simple_identity"):
        (name, arr), = list(inputs.items())[:1]
        dt = utils.np_to_triton_dtype(arr.dtype)
        return [{"OUTPUT0": (dt, arr)}]
    raise ServerError("Request for unknown model: '%s' is not found" % model, 404)


class MockCore:
    """Protocol-independent state: models, shared memory managers, statistics."""

    def __init__(self, delay_us=0):
        self.cuda = _CudaRegions()
        self.system = _SystemRegions()
        self.delay_us = delay_us
        self.stats = {}
        self._lock = threading.Lock()

    def _shm(self, name):
        if self.cuda.has(name):
            return self.cuda
        if self.system.has(name):
            return self.system
        raise ServerError("Unable to find shared memory region: '%s'" % name)

    def known(self, model):
        return model in MODELS or model.startswith("identity")

    def metadata(self, model):
        if not self.known(model):
            raise ServerError("Request for unknown model: '%s' is not found" % model, 404)
        platform, ins, outs, _ = MODELS.get(model, ("mock", [("INPUT0", "BYTES", [-1])], [("OUTPUT0", "BYTES", [-1])], False))
        return {"name": model, "versions": ["1"], "platform": platform,
                "inputs": [{"name": n, "datatype": d, "shape": s} for n, d, s in ins],
                "outputs": [{"name": n, "datatype": d, "shape": s} for n, d, s in outs]}

    def infer(self, model, in_specs, out_specs, params):
        """in_specs: list of dicts {name, datatype, shape, raw | data | shm=(region, size, offset)}
        out_specs: list of dicts {name, shm=(...)|None, binary: bool, classification: int} or None.
        Returns list of responses; each a list of dicts {name, datatype, shape, raw|None (shm)}."""
        if not self.known(model):
            raise ServerError("Request for unknown model: '%s' is not found" % model, 404)
        t0 = time.perf_counter_ns()
        arrays = {}
        for spec in in_specs:
            dt, shape = spec["datatype"], [int(d) for d in spec["shape"]]
            if spec.get("shm") is not None:
                region, size, offset = spec["shm"]
                raw = self._shm(region).read(region, offset, size)
                arrays[spec["name"]] = _decode(dt, shape, raw)
            elif spec.get("raw") is not None:
                want = int(np.prod(shape)) * np.dtype(_NP[dt]).itemsize if dt in _NP else None
                if want is not None and want != len(spec["raw"]):
                    raise ServerError("unexpected size %d for input '%s', expecting %d" % (len(spec["raw"]), spec["name"], want))
                arrays[spec["name"]] = _decode(dt, shape, spec["raw"])
            else:
                data = spec.get("data")
                if dt == "BYTES":
                    arrays[spec["name"]] = np.array([s.encode() if isinstance(s, str) else s for s in data], dtype=object).reshape(shape)
                else:
                    arrays[spec["name"]] = np.array(data, dtype=_NP[dt]).reshape(shape)
        if model in MODELS:
            for name, dt, _ in MODELS[model][1]:
                if name not in arrays:
                    raise ServerError("expected input '%s' for model '%s'" % (name, model))
        if self.delay_us:
            time.sleep(self.delay_us / 1e6)
        responses = run_model(model, arrays, params)
        requested = {o["name"]: o for o in out_specs} if out_specs else None
        out = []
        for resp in responses:
            tensors = []
            for name, (dt, arr) in resp.items():
                if requested is not None and name not in requested:
                    continue
                spec = requested.get(name) if requested else None
                entry = {"name": name, "datatype": dt, "shape": list(arr.shape)}
                if spec and spec.get("classification"):
                    k = int(spec["classification"])
                    flat = np.asarray(arr, dtype=np.float64).reshape(-1)
                    top = np.argsort(-flat, kind="stable")[:k]
                    labels = np.array([("%f:%d:class_%d" % (flat[i], i, i)).encode() for i in top], dtype=object)  # value:index:label
                    entry.update(datatype="BYTES", shape=[len(top)], raw=_encode("BYTES", labels), array=labels)
                elif spec and spec.get("shm") is not None:
                    region, size, offset = spec["shm"]
                    raw = _encode(dt, arr)
                    if len(raw) > size:
                        raise ServerError("shared memory size specified with the request for output '%s' (%d bytes) should be at least %d bytes" % (name, size, len(raw)))
                    self._shm(region).write(region, offset, raw)
                    entry.update(raw=None, shm=True, byte_size=len(raw))
                else:
                    entry.update(raw=_encode(dt, arr), array=arr)
                tensors.append(entry)
            out.append(tensors)
        with self._lock:
            st = self.stats.setdefault(model, {"count": 0, "ns": 0})
            st["count"] += 1
            st["ns"] += time.perf_counter_ns() - t0
        return out

    def decoupled(self, model):
        return MODELS.get(model, (None, None, None, False))[3]


# =====================================================================================
# HTTP front end
# =====================================================================================
def _http_handler(core, verbose=False):
    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"
        wbufsize = 1 << 16  # headers + body leave in one segment (no Nagle / delayed-ACK stall)

        def setup(self):
            super().setup()
            try:
                import socket

                self.connection.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            except OSError:
                pass

        def log_message(self, fmt, *args):
            if verbose:
                sys.stderr.write("HTTP " + fmt % args + "\n")

        def _send(self, status, body=b"", headers=None):
            self.send_response(status)
            self.send_header("Content-Length", str(len(body)))
            for k, v in (headers or {}).items():
                self.send_header(k, str(v))
            if "Content-Type" not in (headers or {}):
                self.send_header("Content-Type", "application/json")
            self.end_headers()
            if body:
                self.wfile.write(body)
            self.wfile.flush()

        def _json(self, obj, status=200):
            self._send(status, json.dumps(obj, separators=(",", ":")).encode())

        def _error(self, ex):
            status = ex.status if isinstance(ex, ServerError) else 500
            self._json({"error": str(ex)}, status)

        def do_GET(self):
            try:
                path = self.path.split("?")[0].strip("/").split("/")
                if path == ["v2", "health", "live"] or path == ["v2", "health", "ready"]:
                    return self._send(200)
                if path == ["v2"]:
                    return self._json({"name": "triton", "version": "0.1-tb200-mock", "extensions": ["binary_tensor_data", "system_shared_memory", "cuda_shared_memory"]})
                if path[:2] == ["v2", "models"] and len(path) >= 3:
                    model = path[2]
                    rest = path[3:]
                    if rest[:1] == ["versions"]:
                        rest = rest[2:]
                    if rest == ["ready"]:
                        return self._send(200 if core.known(model) else 400)
                    if rest == []:
                        return self._json(core.metadata(model))
                    if rest == ["config"]:
                        md = core.metadata(model)
                        return self._json({"name": model, "platform": md["platform"], "max_batch_size": 0,
                                           "input": [{"name": i["name"], "data_type": "TYPE_" + i["datatype"], "dims": i["shape"]} for i in md["inputs"]],
                                           "output": [{"name": o["name"], "data_type": "TYPE_" + o["datatype"], "dims": o["shape"]} for o in md["outputs"]]})
                    if rest == ["stats"]:
                        st = core.stats.get(model, {"count": 0, "ns": 0})
                        return self._json({"model_stats": [{"name": model, "version": "1", "inference_count": st["count"],
                                                            "inference_stats": {"success": {"count": st["count"], "ns": st["ns"]}}}]})
                if path[:2] == ["v2", "systemsharedmemory"] or path[:2] == ["v2", "cudasharedmemory"]:
                    mgr = core.system if path[1] == "systemsharedmemory" else core.cuda
                    name = path[3] if len(path) == 5 and path[2] == "region" else ""
                    return self._json(mgr.status(name))
                raise ServerError("unknown endpoint " + self.path, 404)
            except Exception as ex:  # noqa: BLE001
                self._error(ex)

        def do_POST(self):
            try:
                length = int(self.headers.get("Content-Length", 0))
                body = self.rfile.read(length) if length else b""
                enc = self.headers.get("Content-Encoding")
                if enc == "gzip":
                    import gzip

                    body = gzip.decompress(body)
                elif enc == "deflate":
                    import zlib

                    body = zlib.decompress(body)
                path = self.path.split("?")[0].strip("/").split("/")
                if path[:2] == ["v2", "models"] and path[-1] == "infer":
                    return self._infer(path[2], body)
                if path[1] in ("systemsharedmemory", "cudasharedmemory"):
                    mgr = core.system if path[1] == "systemsharedmemory" else core.cuda
                    action = path[-1]
                    name = path[3] if len(path) == 5 and path[2] == "region" else ""
                    if action == "register":
                        req = json.loads(body)
                        if mgr is core.system:
                            mgr.register(name, req["key"], req.get("offset", 0), req["byte_size"])
                        else:
                            mgr.register(name, base64.b64decode(req["raw_handle"]["b64"]), req["device_id"], req["byte_size"])
                        return self._send(200)
                    if action == "unregister":
                        mgr.unregister(name)
                        return self._send(200)
                if path[:3] == ["v2", "repository", "index"]:
                    return self._json([{"name": m, "version": "1", "state": "READY"} for m in MODELS])
                if path[:2] == ["v2", "repository"]:
                    return self._send(200)
                if path[-2:] == ["trace", "setting"] or path == ["v2", "logging"]:
                    return self._json(json.loads(body) if body else {})
                raise ServerError("unknown endpoint " + self.path, 404)
            except Exception as ex:  # noqa: BLE001
                self._error(ex)

        def _infer(self, model, body):
            hlen = self.headers.get("Inference-Header-Content-Length")
            if hlen is not None:
                header, blob = json.loads(body[: int(hlen)]), body[int(hlen):]
            else:
                header, blob = json.loads(body), b""
            cursor = 0
            in_specs = []
            for t in header.get("inputs", []):
                p = t.get("parameters") or {}
                spec = {"name": t["name"], "datatype": t["datatype"], "shape": t["shape"]}
                if "shared_memory_region" in p:
                    spec["shm"] = (p["shared_memory_region"], p["shared_memory_byte_size"], p.get("shared_memory_offset", 0))
                elif "binary_data_size" in p:
                    n = p["binary_data_size"]
                    spec["raw"] = blob[cursor:cursor + n]
                    cursor += n
                else:
                    spec["data"] = t.get("data")
                in_specs.append(spec)
            params = header.get("parameters") or {}
            all_binary = bool(params.get("binary_data_output"))
            out_specs = None
            if header.get("outputs"):
                out_specs = []
                for o in header["outputs"]:
                    p = o.get("parameters") or {}
                    spec = {"name": o["name"], "binary": bool(p.get("binary_data", False)),
                            "classification": p.get("classification", 0), "shm": None}
                    if "shared_memory_region" in p:
                        spec["shm"] = (p["shared_memory_region"], p["shared_memory_byte_size"], p.get("shared_memory_offset", 0))
                    out_specs.append(spec)
            responses = core.infer(model, in_specs, out_specs, params)
            tensors = responses[0] if responses else []
            want_binary = {o["name"]: o["binary"] for o in out_specs} if out_specs else {}
            out_json = {"model_name": model, "model_version": "1", "outputs": []}
            if header.get("id"):
                out_json["id"] = header["id"]
            blobs = []
            for t in tensors:
                entry = {"name": t["name"], "datatype": t["datatype"], "shape": t["shape"]}
                if t.get("shm"):
                    entry["parameters"] = {"shared_memory_byte_size": t["byte_size"]}
                elif all_binary or want_binary.get(t["name"], False):
                    entry["parameters"] = {"binary_data_size": len(t["raw"])}
                    blobs.append(t["raw"])
                else:
                    arr = t["array"]
                    if t["datatype"] == "BYTES":
                        entry["data"] = [x.decode("utf-8", "replace") if isinstance(x, bytes) else str(x) for x in arr.reshape(-1).tolist()]
                    else:
                        entry["data"] = np.asarray(arr).reshape(-1).tolist()
                out_json["outputs"].append(entry)
            head = json.dumps(out_json, separators=(",", ":")).encode()
            headers = {}
            if blobs:
                headers["Inference-Header-Content-Length"] = len(head)
                headers["Content-Type"] = "application/octet-stream"
            payload = head + b"".join(blobs)
            accept = self.headers.get("Accept-Encoding")
            if accept == "gzip":
                import gzip

                payload = gzip.compress(payload)
                headers["Content-Encoding"] = "gzip"
            elif accept == "deflate":
                import zlib

                payload = zlib.compress(payload)
                headers["Content-Encoding"] = "deflate"
            self._send(200, payload, headers)

    return Handler


# =====================================================================================
# gRPC front end
# =====================================================================================
def _grpc_servicer(core):
    import grpc

    from ..grpc import service_pb2 as pb
    from ..grpc import service_pb2_grpc as pbg

    def param_value(p):
        which = p.WhichOneof("parameter_choice")
        return getattr(p, which) if which else None

    def parse_request(request):
        in_specs = []
        raw_iter = iter(request.raw_input_contents)
        for t in request.inputs:
            spec = {"name": t.name, "datatype": t.datatype, "shape": list(t.shape)}
            p = {k: param_value(v) for k, v in t.parameters.items()}
            if "shared_memory_region" in p:
                spec["shm"] = (p["shared_memory_region"], p["shared_memory_byte_size"], p.get("shared_memory_offset", 0))
            else:
                try:
                    spec["raw"] = next(raw_iter)
                except StopIteration:
                    c = t.contents
                    for field, dt in (("int_contents", np.int32), ("int64_contents", np.int64), ("fp32_contents", np.float32),
                                      ("fp64_contents", np.float64), ("uint_contents", np.uint32), ("uint64_contents", np.uint64),
                                      ("bool_contents", np.bool_)):
                        vals = getattr(c, field)
                        if len(vals):
                            spec["raw"] = np.array(list(vals), dtype=dt).astype(_NP[t.datatype]).tobytes()
                            break
                    else:
                        if len(c.bytes_contents):
                            spec["raw"] = _encode("BYTES", np.array(list(c.bytes_contents), dtype=object))
                        else:
                            spec["raw"] = b""
            in_specs.append(spec)
        out_specs = None
        if len(request.outputs):
            out_specs = []
            for o in request.outputs:
                p = {k: param_value(v) for k, v in o.parameters.items()}
                spec = {"name": o.name, "binary": True, "classification": p.get("classification", 0), "shm": None}
                if "shared_memory_region" in p:
                    spec["shm"] = (p["shared_memory_region"], p["shared_memory_byte_size"], p.get("shared_memory_offset", 0))
                out_specs.append(spec)
        params = {k: param_value(v) for k, v in request.parameters.items()}
        return in_specs, out_specs, params

    def build_response(request, tensors):
        resp = pb.ModelInferResponse(model_name=request.model_name, model_version="1", id=request.id)
        for t in tensors:
            o = resp.outputs.add()
            o.name, o.datatype = t["name"], t["datatype"]
            o.shape.extend(t["shape"])
            if t.get("shm"):
                o.parameters["shared_memory_byte_size"].int64_param = t["byte_size"]
            else:
                resp.raw_output_contents.append(t["raw"])
        return resp

    class Servicer(pbg.GRPCInferenceServiceServicer):
        def ServerLive(self, request, context):
            return pb.ServerLiveResponse(live=True)

        def ServerReady(self, request, context):
            return pb.ServerReadyResponse(ready=True)

        def ModelReady(self, request, context):
            return pb.ModelReadyResponse(ready=core.known(request.name))

        def ServerMetadata(self, request, context):
            return pb.ServerMetadataResponse(name="triton", version="0.1-tb200-mock", extensions=["system_shared_memory", "cuda_shared_memory"])

        def ModelMetadata(self, request, context):
            try:
                md = core.metadata(request.name)
            except ServerError as ex:
                context.abort(grpc.StatusCode.NOT_FOUND, str(ex))
            resp = pb.ModelMetadataResponse(name=md["name"], versions=md["versions"], platform=md["platform"])
            for key in ("inputs", "outputs"):
                for t in md[key]:
                    e = getattr(resp, key).add()
                    e.name, e.datatype = t["name"], t["datatype"]
                    e.shape.extend(t["shape"])
            return resp

        def ModelConfig(self, request, context):
            try:
                md = core.metadata(request.name)
            except ServerError as ex:
                context.abort(grpc.StatusCode.NOT_FOUND, str(ex))
            resp = pb.ModelConfigResponse()
            resp.config.name, resp.config.platform = md["name"], md["platform"]
            resp.config.model_transaction_policy.decoupled = core.decoupled(request.name)
            for t in md["inputs"]:
                e = resp.config.input.add()
                e.name = t["name"]
                e.data_type = utils_dtype_enum(t["datatype"])
                e.dims.extend(t["shape"])
                if len(t["shape"]) == 3 and t["shape"][0] in (1, 3):  # image-shaped input: CHW
                    e.format = 2  # FORMAT_NCHW
            for t in md["outputs"]:
                e = resp.config.output.add()
                e.name = t["name"]
                e.data_type = utils_dtype_enum(t["datatype"])
                e.dims.extend(t["shape"])
            return resp

        def ModelStatistics(self, request, context):
            resp = pb.ModelStatisticsResponse()
            for name, st in core.stats.items():
                if request.name and request.name != name:
                    continue
                m = resp.model_stats.add()
                m.name, m.version, m.inference_count = name, "1", st["count"]
                m.inference_stats.success.count, m.inference_stats.success.ns = st["count"], st["ns"]
            return resp

        def RepositoryIndex(self, request, context):
            resp = pb.RepositoryIndexResponse()
            for m in MODELS:
                e = resp.models.add()
                e.name, e.version, e.state = m, "1", "READY"
            return resp

        def RepositoryModelLoad(self, request, context):
            return pb.RepositoryModelLoadResponse()

        def RepositoryModelUnload(self, request, context):
            return pb.RepositoryModelUnloadResponse()

        def SystemSharedMemoryStatus(self, request, context):
            resp = pb.SystemSharedMemoryStatusResponse()
            try:
                for r in core.system.status(request.name):
                    e = resp.regions[r["name"]]
                    e.name, e.key, e.offset, e.byte_size = r["name"], r["key"], r["offset"], r["byte_size"]
            except ServerError as ex:
                context.abort(grpc.StatusCode.NOT_FOUND, str(ex))
            return resp

        def SystemSharedMemoryRegister(self, request, context):
            try:
                core.system.register(request.name, request.key, request.offset, request.byte_size)
            except ServerError as ex:
                context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(ex))
            return pb.SystemSharedMemoryRegisterResponse()

        def SystemSharedMemoryUnregister(self, request, context):
            core.system.unregister(request.name)
            return pb.SystemSharedMemoryUnregisterResponse()

        def CudaSharedMemoryStatus(self, request, context):
            resp = pb.CudaSharedMemoryStatusResponse()
            try:
                for r in core.cuda.status(request.name):
                    e = resp.regions[r["name"]]
                    e.name, e.device_id, e.byte_size = r["name"], r["device_id"], r["byte_size"]
            except ServerError as ex:
                context.abort(grpc.StatusCode.NOT_FOUND, str(ex))
            return resp

        def CudaSharedMemoryRegister(self, request, context):
            try:
                core.cuda.register(request.name, request.raw_handle, request.device_id, request.byte_size)
            except ServerError as ex:
                context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(ex))
            return pb.CudaSharedMemoryRegisterResponse()

        def CudaSharedMemoryUnregister(self, request, context):
            core.cuda.unregister(request.name)
            return pb.CudaSharedMemoryUnregisterResponse()

        def TraceSetting(self, request, context):
            resp = pb.TraceSettingResponse()
            for k, v in request.settings.items():
                resp.settings[k].value.extend(v.value)
            return resp

        def LogSettings(self, request, context):
            resp = pb.LogSettingsResponse()
            for k, v in request.settings.items():
                resp.settings[k].CopyFrom(pb.LogSettingsResponse.SettingValue.FromString(v.SerializeToString()))
            return resp

        def ModelInfer(self, request, context):
            try:
                in_specs, out_specs, params = parse_request(request)
                if core.decoupled(request.model_name):
                    raise ServerError("ModelInfer RPC doesn't support models with decoupled transaction policy")
                responses = core.infer(request.model_name, in_specs, out_specs, params)
                return build_response(request, responses[0])
            except ServerError as ex:
                context.abort(grpc.StatusCode.NOT_FOUND if ex.status == 404 else grpc.StatusCode.INVALID_ARGUMENT, str(ex))

        def ModelStreamInfer(self, request_iterator, context):
            for request in request_iterator:
                try:
                    in_specs, out_specs, params = parse_request(request)
                    responses = core.infer(request.model_name, in_specs, out_specs, params)
                    decoupled = core.decoupled(request.model_name)
                    for k, tensors in enumerate(responses):
                        resp = build_response(request, tensors)
                        if decoupled:
                            final = k == len(responses) - 1 and not params.get("triton_enable_empty_final_response")
                            resp.parameters["triton_final_response"].bool_param = final
                        yield pb.ModelStreamInferResponse(infer_response=resp)
                    if decoupled and params.get("triton_enable_empty_final_response"):
                        resp = pb.ModelInferResponse(model_name=request.model_name, model_version="1", id=request.id)
                        resp.parameters["triton_final_response"].bool_param = True
                        yield pb.ModelStreamInferResponse(infer_response=resp)
                except ServerError as ex:
                    err = pb.ModelStreamInferResponse(error_message=str(ex))
                    err.infer_response.id = request.id
                    yield err

    return Servicer(), pbg


def utils_dtype_enum(name):
    from ..grpc import model_config_pb2 as mc

    return getattr(mc, "TYPE_STRING" if name == "BYTES" else "TYPE_" + name)


class MockServer:
    """HTTP + gRPC endpoints over one MockCore (in-process handle; ``main()`` runs it
    as the separate server process)."""

    def __init__(self, http_port=0, grpc_port=0, host="127.0.0.1", delay_us=0, verbose=False, grpc_workers=16):
        import grpc

        self.core = MockCore(delay_us)
        self.httpd = ThreadingHTTPServer((host, http_port), _http_handler(self.core, verbose))
        self.httpd.daemon_threads = True
        self.http_port = self.httpd.server_address[1]
        servicer, pbg = _grpc_servicer(self.core)
        self.grpc_server = grpc.server(
            futures.ThreadPoolExecutor(max_workers=grpc_workers),
            options=[("grpc.max_send_message_length", 2**31 - 1), ("grpc.max_receive_message_length", 2**31 - 1)],
        )
        pbg.add_GRPCInferenceServiceServicer_to_server(servicer, self.grpc_server)
        self.grpc_port = self.grpc_server.add_insecure_port("%s:%d" % (host, grpc_port))
        self._thread = None

    def start(self):
        self.grpc_server.start()
        self._thread = threading.Thread(target=self.httpd.serve_forever, daemon=True)
        self._thread.start()
        return self

    def stop(self):
        self.httpd.shutdown()
        self.httpd.server_close()
        self.grpc_server.stop(0)
        self.core.cuda.unregister()
        self.core.system.unregister()


def main(argv=None):
    ap = argparse.ArgumentParser(description="tb200 mock KServe-v2 server")
    ap.add_argument("--http-port", type=int, default=8000)
    ap.add_argument("--grpc-port", type=int, default=8001)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--delay-us", type=int, default=0, help="fixed model latency")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args(argv)
    srv = MockServer(args.http_port, args.grpc_port, args.host, args.delay_us, args.verbose).start()
    print("READY http=%d grpc=%d" % (srv.http_port, srv.grpc_port), flush=True)
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        srv.stop()


if __name__ == "__main__":
    main()
