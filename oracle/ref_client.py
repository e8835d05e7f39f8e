"""ref_client.py -- TEST / BENCH INFRASTRUCTURE (oracle/): the reference client's C2 request loop,
restated, for the reference arm of bench.py.

Only tests/ and bench.py's `--impl reference` / `cpu_baseline` legs may import this; the product
(client_b200/, libtb200.so) never does.  Nothing here calls into client_b200: the CUDA work goes
through cuda-python exactly as the reference does, the wire body comes from oracle/wire.py, the
transport is the standard library's http.client (the reference's geventhttpclient is not in this
image).

What it restates -- BASELINE configs[1] as SURVEY.md 8(d) defines the CPU baseline, i.e. the flow of
src/python/examples/simple_http_cudashm_client.py:82-139 per request:

  create_shared_memory_region   PY/utils/cuda_shared_memory/__init__.py:107-149
                                cudaSetDevice, cudaMalloc, cudaIpcGetMemHandle
  get_raw_handle                :152-170   base64 of handle.reserved
  register_cuda_shared_memory   PY/http/_client.py:1129-1175  POST v2/cudasharedmemory/region/<n>/register
  set_shared_memory_region      :173-239   np.ascontiguousarray(x).flatten(); cudaMemcpyAsync(
                                base + offset, host pointer, cudaMemcpyDefault, stream) per array;
                                cudaStreamSynchronize
  infer                         PY/http/_client.py:1331-1484 with PY/http/_utils.py:90-151 (body: oracle/wire.py)
  get_contents_as_numpy         :242-325   cudaMemcpyAsync of the WHOLE region device->host,
                                cudaStreamSynchronize, np.frombuffer(...)[:n] reshaped and copied
(PY = src/python/library/tritonclient).  parity: unpinned beyond the wire body (pinned by
tests/golden/wire_golden.json) -- there is no GPU where the reference tree is mounted.
"""

import base64
import http.client
import json
import time

import numpy as np

IN_SHAPE = (3, 224, 224)
IN_BYTES = 3 * 224 * 224 * 4
OUT_ELEMS = 1000
OUT_BYTES = OUT_ELEMS * 4


def _cuda(fn, *args):
    """call_cuda_function (PY/utils/cuda_shared_memory/_utils.py:33-46): raise on a non-zero status."""
    res = fn(*args)
    if res[0].value != 0:
        raise RuntimeError("%s failed: %s" % (fn.__name__, res[0]))
    return res[1] if len(res) == 2 else (None if len(res) == 1 else res[1:])


class Region:
    """CudaSharedMemoryRegion (PY/utils/cuda_shared_memory/_utils.py:67-100)."""

    def __init__(self, name, byte_size, device_id):
        from cuda.bindings import runtime as cudart

        self._rt = cudart
        self.name, self.byte_size, self.device_id = name, byte_size, device_id
        _cuda(cudart.cudaSetDevice, device_id)
        self.ptr = _cuda(cudart.cudaMalloc, byte_size)
        self.handle = _cuda(cudart.cudaIpcGetMemHandle, self.ptr)

    def raw_handle(self):
        return base64.b64encode(self.handle.reserved)

    def close(self):
        if self.ptr:
            self._rt.cudaFree(self.ptr)
            self.ptr = 0


class RefClient:
    """One reference-style client: a keep-alive HTTP connection, a CUDA stream, an input and an
    output region registered with the server."""

    def __init__(self, url, device_id, tag):
        from cuda.bindings import runtime as cudart

        from oracle import wire

        self._rt = cudart
        self._wire = wire
        host, _, port = url.partition(":")
        self.conn = http.client.HTTPConnection(host, int(port or 80), timeout=30)
        self.device_id = device_id
        self.in_name, self.out_name = "ref_in_%s" % tag, "ref_out_%s" % tag
        self.inp = Region(self.in_name, IN_BYTES, device_id)
        self.out = Region(self.out_name, OUT_BYTES, device_id)
        self.stream = _cuda(cudart.cudaStreamCreate)
        for r in (self.inp, self.out):
            self._post("v2/cudasharedmemory/region/%s/register" % r.name,
                       json.dumps({"raw_handle": {"b64": r.raw_handle().decode()}, "device_id": device_id, "byte_size": r.byte_size},
                                  separators=(",", ":")).encode())
        i = wire.HttpInput("data_0", list(IN_SHAPE), "FP32")
        i.set_shm(self.in_name, IN_BYTES)
        o = wire.HttpOutput("fc6_1")
        o.set_shm(self.out_name, OUT_BYTES)
        self._in, self._out = i, o
        self._host_out = np.empty(OUT_BYTES, np.uint8)

    def _post(self, uri, body, headers=None):
        self.conn.request("POST", "/" + uri, body=body, headers=headers or {})
        resp = self.conn.getresponse()
        data = resp.read()
        if resp.status != 200:
            raise RuntimeError("%s -> %d %s" % (uri, resp.status, data[:200]))
        return data

    def set_input(self, x):
        """set_shared_memory_region(handle, [x])."""
        rt = self._rt
        flat = np.ascontiguousarray(x).flatten()
        _cuda(rt.cudaMemcpyAsync, self.inp.ptr, flat.ctypes.data, flat.size * flat.itemsize, rt.cudaMemcpyKind.cudaMemcpyDefault, self.stream)
        _cuda(rt.cudaStreamSynchronize, self.stream)

    def infer(self):
        body, json_size = self._wire.http_request_body([self._in], [self._out])
        headers = {"Inference-Header-Content-Length": str(json_size)} if json_size is not None else {}
        return self._post("v2/models/densenet_onnx/infer", body, headers)

    def get_output(self):
        """get_contents_as_numpy(handle, np.float32, [1000]): whole region D2H, then a copy of the view."""
        rt = self._rt
        _cuda(rt.cudaMemcpyAsync, self._host_out.ctypes.data, self.out.ptr, OUT_BYTES, rt.cudaMemcpyKind.cudaMemcpyDefault, self.stream)
        _cuda(rt.cudaStreamSynchronize, self.stream)
        return np.copy(np.frombuffer(self._host_out, dtype=np.float32)[:OUT_ELEMS].reshape([OUT_ELEMS]))

    def close(self):
        try:
            for r in (self.inp, self.out):
                self._post("v2/cudasharedmemory/region/%s/unregister" % r.name, b"")
        except Exception:
            pass
        self.conn.close()
        self.inp.close()
        self.out.close()


def run_loop(url, device_id, tag, seconds, data_mode, ready=None, go=None):
    """Free-running closed loop of ONE client for `seconds`; returns (completed, latencies_ns).
    data_mode "per-request": a fresh numpy tensor for every request (the synthetic input step of
    SURVEY.md 8d); "once": the regions are filled once and every request only names them (what
    perf_analyzer does) -- no per-request host work besides the request itself."""
    c = RefClient(url, device_id, tag)
    rng = np.random.default_rng(abs(hash(tag)) % (1 << 32))
    c.set_input(rng.random(IN_SHAPE, dtype=np.float32))
    c.infer()
    if ready is not None:
        ready.wait()
    if go is not None:
        go.wait()
    lat = []
    t_end = time.perf_counter() + seconds
    while time.perf_counter() < t_end:
        t0 = time.perf_counter_ns()
        if data_mode == "per-request":
            c.set_input(rng.random(IN_SHAPE, dtype=np.float32))
            c.infer()
            y = c.get_output()
            if not np.isfinite(y).all():
                raise RuntimeError("non-finite logits")
        else:
            c.infer()
        lat.append(time.perf_counter_ns() - t0)
    c.close()
    return len(lat), lat
