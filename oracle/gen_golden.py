"""Generate tests/golden/*.json from the REFERENCE code itself.  TEST INFRASTRUCTURE.

Runs only in the build container (needs /root/reference); the fixtures it writes
are committed so the tests on the GPU box never read the reference.

How the reference is imported without its missing dependencies:
  * ``rapidjson``         -> stdlib json, compact separators (equivalent for the
                             int/str/bool headers pinned here; floats unpinned)
  * ``gevent``, ``gevent.pool``, ``geventhttpclient``, ``geventhttpclient.url``
                          -> inert stubs (no network is used)
  * ``tritonclient.grpc.{service_pb2, service_pb2_grpc, model_config_pb2}``
                          -> message classes built from the schema table in
                             client_b200/grpc/_proto.py (the generated modules of
                             the reference come from another repo); the request
                             assembly under test is the reference's own
                             grpc/_utils.py and grpc/_infer_input.py.

Usage:  python oracle/gen_golden.py   (writes tests/golden/wire_golden.json and
        tests/golden/image_golden.npz)
"""

import base64
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/python/library"
REF_EXAMPLES = "/root/reference/src/python/examples"
OUT_DIR = os.path.join(ROOT, "tests", "golden")


def _install_shims():
    rj = types.ModuleType("rapidjson")
    # python-rapidjson's default bytes_mode (BM_UTF8) writes bytes values as UTF-8
    # strings; the reference relies on it for the base64 raw_handle (http/_client.py:1160)
    def _bytes_as_str(o):
        if isinstance(o, (bytes, bytearray)):
            return bytes(o).decode("utf-8")
        raise TypeError(type(o))

    rj.dumps = lambda o, **kw: json.dumps(o, separators=(",", ":"), default=_bytes_as_str)
    rj.loads = json.loads
    sys.modules["rapidjson"] = rj
    gevent = types.ModuleType("gevent")
    gevent.sleep = lambda s=0: None
    gevent.Timeout = type("Timeout", (Exception,), {})
    pool = types.ModuleType("gevent.pool")
    pool.Pool = lambda *a, **k: types.SimpleNamespace(join=lambda: None)
    gevent.pool = pool
    sys.modules["gevent"] = gevent
    sys.modules["gevent.pool"] = pool
    ghc = types.ModuleType("geventhttpclient")

    class HTTPClient:
        @classmethod
        def from_url(cls, url, **kw):
            return cls()

        def close(self):
            pass

    ghc.HTTPClient = HTTPClient
    url = types.ModuleType("geventhttpclient.url")

    class URL:
        def __init__(self, u):
            self.request_uri = "/" + u.split("/", 3)[3] if u.count("/") > 2 else ""

    url.URL = URL
    ghc.url = url
    sys.modules["geventhttpclient"] = ghc
    sys.modules["geventhttpclient.url"] = url
    # protobuf classes from the schema table (loaded by path, not via the product package)
    spec = importlib.util.spec_from_file_location("_tb200_proto", os.path.join(ROOT, "client_b200", "grpc", "_proto.py"))
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    pb2, pb2_grpc, mc = proto.build_modules()
    sys.path.insert(0, REF)
    import tritonclient  # noqa: F401  (the reference package)

    assert tritonclient.__file__.startswith(REF), tritonclient.__file__
    sys.modules["tritonclient.grpc.service_pb2"] = pb2
    sys.modules["tritonclient.grpc.service_pb2_grpc"] = pb2_grpc
    sys.modules["tritonclient.grpc.model_config_pb2"] = mc
    return pb2


def hx(b):
    return bytes(b).hex()


def gen_wire():
    _install_shims()
    import tritonclient.grpc as rgrpc
    import tritonclient.http as rhttp
    import tritonclient.utils as rutils
    from tritonclient.grpc._utils import _get_inference_request as grpc_req

    assert rhttp.__file__.startswith(REF)
    G = {"generator": "oracle/gen_golden.py", "reference": "triton-inference-server/client @ 58a44caa", "cases": {}}
    C = G["cases"]

    # ---- utils codecs -------------------------------------------------------------
    codecs = []
    for arr in [np.array([b"ab", b"c"], dtype=object), np.array([[b"hello", b"world"]], dtype=object),
                np.array(["x", 12, b"\x00\xff"], dtype=object), np.array([b"abc", b"de"], dtype="S3"),
                np.array([], dtype=object), np.array([b""], dtype=object)]:
        ser = rutils.serialize_byte_tensor(arr)
        wire = ser.item() if ser.size > 0 else b""
        dec = rutils.deserialize_bytes_tensor(wire)
        codecs.append({"kind": "BYTES", "dtype": str(arr.dtype), "shape": list(arr.shape),
                       "items": [hx(x) if isinstance(x, bytes) else ("str:" + str(x)) for x in arr.reshape(-1).tolist()],
                       "wire": hx(wire), "decoded": [hx(x) for x in dec.tolist()],
                       "byte_size": int(rutils.serialized_byte_size(ser)) if ser.size > 0 and ser.dtype == np.object_ else 0})
    rng = np.random.default_rng(7)
    for arr in [np.array([1.0, -2.5], dtype=np.float32), rng.standard_normal((3, 5)).astype(np.float32),
                np.array([np.inf, -0.0, 3.4e38, 1e-40, 1.0000001], dtype=np.float32)]:
        ser = rutils.serialize_bf16_tensor(arr)
        wire = ser.item()
        dec = rutils.deserialize_bf16_tensor(wire)
        codecs.append({"kind": "BF16", "shape": list(arr.shape), "f32": hx(arr.tobytes()), "wire": hx(wire),
                       "decoded_shape": list(dec.shape), "decoded": hx(dec.astype("<f4").tobytes())})
    C["codecs"] = codecs
    C["dtype_map"] = {name: str(np.dtype(rutils.triton_to_np_dtype(name))) for name in
                      ["BOOL", "INT8", "INT16", "INT32", "INT64", "UINT8", "UINT16", "UINT32", "UINT64", "FP16", "FP32", "FP64", "BF16", "BYTES"]}

    # ---- HTTP request bodies --------------------------------------------------------
    def http_case(name, build, **kw):
        inputs, outputs = build(rhttp)
        body, js = rhttp.InferenceServerClient.generate_request_body(inputs, outputs=outputs, **kw)
        C[name] = {"json_size": js, "body": hx(body), "sha256": hashlib.sha256(body).hexdigest(), "kwargs": kw}

    a16 = np.arange(16, dtype=np.int32)[None, :]
    m16 = np.full((1, 16), -1, dtype=np.int32)

    def cfg1(H):
        i0 = H.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a16, binary_data=False)
        i1 = H.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(m16, binary_data=True)
        return [i0, i1], [H.InferRequestedOutput("OUTPUT0", binary_data=True), H.InferRequestedOutput("OUTPUT1", binary_data=False)]

    def cudashm_a(H):
        i0 = H.InferInput("INPUT0", [1, 16], "INT32").set_shared_memory("input0_data", 64)
        i1 = H.InferInput("INPUT1", [1, 16], "INT32").set_shared_memory("input1_data", 64, offset=64)
        o0 = H.InferRequestedOutput("OUTPUT0", binary_data=True)
        o0.set_shared_memory("output0_data", 64)
        o1 = H.InferRequestedOutput("OUTPUT1", binary_data=True)
        o1.set_shared_memory("output1_data", 64)
        return [i0, i1], [o0, o1]

    def two_binary(H):
        return [H.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a16),
                H.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(m16)], None

    def mixed_types(H):
        i0 = H.InferInput("S", [1, 2], "BYTES").set_data_from_numpy(np.array([[b"ab", "c"]], dtype=object))
        i1 = H.InferInput("B", [2], "BF16").set_data_from_numpy(np.array([1.0, -2.5], dtype=np.float32))
        i2 = H.InferInput("H", [2], "FP16").set_data_from_numpy(np.array([1.0, -2.5], dtype=np.float16))
        return [i0, i1, i2], [H.InferRequestedOutput("OUT", class_count=3)]

    def json_data(H):
        i0 = H.InferInput("B", [2, 2], "BOOL").set_data_from_numpy(np.array([[True, False], [False, True]]), binary_data=False)
        i1 = H.InferInput("S", [2], "BYTES").set_data_from_numpy(np.array([b"ab", "cd"], dtype=object), binary_data=False)
        i2 = H.InferInput("U", [3], "UINT64").set_data_from_numpy(np.array([0, 1, 2**64 - 1], dtype=np.uint64), binary_data=False)
        return [i0, i1, i2], None

    def densenet(H):
        x = np.random.default_rng(0).random((3, 224, 224), dtype=np.float32)
        return [H.InferInput("data_0", [3, 224, 224], "FP32").set_data_from_numpy(x)], [H.InferRequestedOutput("fc6_1", class_count=0)]

    def reuse_switch(H):
        # data -> shm -> data again on the same object; output shm set then unset
        i0 = H.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a16)
        i0.set_shared_memory("r", 64, 8)
        i0.set_data_from_numpy(m16)
        o = H.InferRequestedOutput("OUTPUT0")
        o.set_shared_memory("o", 64)
        o.unset_shared_memory()
        return [i0], [o]

    http_case("http_config1", cfg1)
    http_case("http_cudashm_A", cudashm_a)
    http_case("http_B_params", two_binary, request_id="7", sequence_id=5, sequence_start=True, sequence_end=False,
              priority=2, timeout=1000, parameters={"k": "v"})
    http_case("http_seq_string", two_binary, sequence_id="abc", sequence_start=False, sequence_end=True)
    http_case("http_C_mixed", mixed_types)
    http_case("http_json_data", json_data)
    http_case("http_densenet_fp32", densenet)
    http_case("http_reuse_switch", reuse_switch)
    C["http_densenet_fp32"]["body"] = C["http_densenet_fp32"]["body"][: 2 * (C["http_densenet_fp32"]["json_size"] + 64)]  # header + 64 B; sha256 pins the rest

    # ---- error texts ---------------------------------------------------------------------
    errors = {}
    for label, fn in {
        "wrong_dtype": lambda: rhttp.InferInput("X", [2], "INT32").set_data_from_numpy(np.zeros(2, np.float32)),
        "wrong_shape": lambda: rhttp.InferInput("X", [2, 3], "INT32").set_data_from_numpy(np.zeros((3, 2), np.int32)),
        "wrong_rank": lambda: rhttp.InferInput("X", [2], "INT32").set_data_from_numpy(np.zeros((2, 1), np.int32)),
        "not_numpy": lambda: rhttp.InferInput("X", [2], "INT32").set_data_from_numpy([1, 2]),
        "bf16_dtype": lambda: rhttp.InferInput("X", [2], "BF16").set_data_from_numpy(np.zeros(2, np.float16)),
        "bf16_json": lambda: rhttp.InferInput("X", [2], "BF16").set_data_from_numpy(np.zeros(2, np.float32), binary_data=False),
        "reserved_param": lambda: rhttp.InferenceServerClient.generate_request_body(two_binary(rhttp)[0], parameters={"priority": 1}),
        "reserved_prefix": lambda: rhttp.InferenceServerClient.generate_request_body(two_binary(rhttp)[0], parameters={"triton_x": 1}),
        "class_shm": lambda: rhttp.InferRequestedOutput("O", class_count=2).set_shared_memory("r", 8),
        "grpc_wrong_dtype": lambda: rgrpc.InferInput("X", [2], "INT32").set_data_from_numpy(np.zeros(2, np.float32)),
        "grpc_wrong_shape": lambda: rgrpc.InferInput("X", [2, 3], "INT32").set_data_from_numpy(np.zeros((3, 2), np.int32)),
        "grpc_param_type": lambda: grpc_req("m", [], "", "", None, 0, False, False, 0, None, {"k": [1]}),
    }.items():
        try:
            fn()
            errors[label] = None
        except Exception as e:  # noqa: BLE001
            errors[label] = {"type": type(e).__name__, "text": str(e)}
    C["errors"] = errors

    # ---- control-plane bodies (captured from the reference client methods) ---------------
    captured = []

    class Capture(rhttp.InferenceServerClient):
        def _post(self, request_uri, request_body, headers, query_params):
            captured.append({"uri": request_uri, "body": request_body if isinstance(request_body, str) else request_body.decode("latin1"),
                             "headers": headers, "query": query_params})
            return types.SimpleNamespace(status_code=200, read=lambda length=-1: b'{"outputs":[]}', get=lambda k: None)

        def _get(self, request_uri, headers, query_params):
            captured.append({"uri": request_uri, "body": None, "headers": headers, "query": query_params})
            return types.SimpleNamespace(status_code=200, read=lambda length=-1: b"{}", get=lambda k: None)

    cl = Capture("localhost:8000")
    handle64 = base64.b64encode(bytes(range(64)))
    cl.register_cuda_shared_memory("input0_data", handle64, 0, 64)
    cl.register_cuda_shared_memory("name with space/slash", handle64, 3, 38535168)
    cl.register_system_shared_memory("output0_data", "/output0_simple", 64, offset=8)
    cl.unregister_cuda_shared_memory("input0_data")
    cl.unregister_cuda_shared_memory()
    cl.unregister_system_shared_memory("a b")
    cl.get_cuda_shared_memory_status("r1")
    cl.get_system_shared_memory_status()
    cl.is_model_ready("simple", "2")
    cl.get_model_metadata("dense net")
    cl.get_model_config("m", "1")
    cl.get_inference_statistics("m")
    cl.get_inference_statistics()
    cl.unload_model("m", unload_dependents=True)
    cl.load_model("m", config='{"name":"m"}')
    cl.update_trace_settings("m", {"trace_rate": "1"})
    cl.update_log_settings({"log_verbose_level": 1})
    inputs, outputs = cfg1(rhttp)
    cl.infer("simple", inputs, model_version="3", outputs=outputs, request_id="r1", query_params={"a": [1, "x y"], "b": "c&d"})
    inputs, outputs = cudashm_a(rhttp)
    cl.infer("simple", inputs, outputs=outputs, request_compression_algorithm=None, response_compression_algorithm="gzip")
    for c in captured:
        if c["headers"]:
            c["headers"] = {k: str(v) for k, v in c["headers"].items()}
    C["http_control_plane"] = captured

    # ---- HTTP response parsing ----------------------------------------------------------------
    hdr = {"model_name": "m", "outputs": [
        {"name": "OUTPUT0", "datatype": "INT32", "shape": [1, 4], "parameters": {"binary_data_size": 16}},
        {"name": "OUTPUT1", "datatype": "INT32", "shape": [1, 2], "data": [5, 6]},
        {"name": "S", "datatype": "BYTES", "shape": [2], "parameters": {"binary_data_size": 11}},
        {"name": "B", "datatype": "BF16", "shape": [2], "parameters": {"binary_data_size": 4}},
        {"name": "E", "datatype": "FP32", "shape": [0], "parameters": {"binary_data_size": 0}},
    ]}
    hj = json.dumps(hdr, separators=(",", ":")).encode()
    body = hj + np.arange(4, dtype=np.int32).tobytes() + bytes.fromhex("0200000061620100000063") + bytes.fromhex("803f20c0")
    res = rhttp.InferenceServerClient.parse_response_body(body, header_length=len(hj))
    parsed = {}
    for n in ["OUTPUT0", "OUTPUT1", "S", "B", "E", "missing"]:
        v = res.as_numpy(n)
        if v is None:
            parsed[n] = None
        elif v.dtype == np.object_:
            parsed[n] = {"dtype": "object", "shape": list(v.shape), "items": [hx(x) for x in v.reshape(-1).tolist()]}
        else:
            parsed[n] = {"dtype": str(v.dtype), "shape": list(v.shape), "data": hx(np.ascontiguousarray(v).tobytes())}
    C["http_response"] = {"body": hx(body), "header_length": len(hj), "parsed": parsed}

    # ---- gRPC requests ----------------------------------------------------------------------------
    def grpc_case(name, inputs, **kw):
        args = dict(model_name="m", inputs=inputs, model_version="", request_id="", outputs=None, sequence_id=0,
                    sequence_start=False, sequence_end=False, priority=0, timeout=None, parameters=None)
        args.update(kw)
        req = grpc_req(**args)
        C[name] = {"bytes": hx(req.SerializeToString()), "deterministic": kw.pop("_det", None)}

    ids = (np.arange(384, dtype=np.int64) * 79 % 30522).reshape(1, 384)
    mask = np.ones((1, 384), dtype=np.int64)
    bert = [rgrpc.InferInput("input_ids", [1, 384], "INT64").set_data_from_numpy(ids),
            rgrpc.InferInput("attention_mask", [1, 384], "INT64").set_data_from_numpy(mask)]
    grpc_case("grpc_bert_raw", bert, model_name="bert_large")
    tok = (np.arange(4096, dtype=np.int32) * 31 % 128256).reshape(1, 4096)
    llama = [rgrpc.InferInput("input_ids", [1, 4096], "INT32").set_data_from_numpy(tok)]
    grpc_case("grpc_llama_stream", llama, model_name="llama3_8b", request_id="42",
              outputs=[rgrpc.InferRequestedOutput("logits")])
    grpc_case("grpc_one_param", bert[:1], priority=3)
    shm_in = [rgrpc.InferInput("INPUT0", [1, 16], "INT32").set_shared_memory("input0_data", 64),
              rgrpc.InferInput("INPUT1", [1, 16], "INT32").set_shared_memory("input1_data", 64, offset=64)]
    o0 = rgrpc.InferRequestedOutput("OUTPUT0")
    o0.set_shared_memory("output0_data", 64)
    grpc_case("grpc_cudashm", shm_in, model_name="simple", outputs=[o0, rgrpc.InferRequestedOutput("OUTPUT1", class_count=2)])
    grpc_case("grpc_params_many", bert[:1], request_id="9", sequence_id="s1", sequence_start=True, sequence_end=True,
              priority=7, timeout=123, parameters={"a": "x", "b": True, "c": 5, "d": 1.5})
    mixed = [rgrpc.InferInput("S", [1, 2], "BYTES").set_data_from_numpy(np.array([[b"ab", "c"]], dtype=object)),
             rgrpc.InferInput("B", [2], "BF16").set_data_from_numpy(np.array([1.0, -2.5], dtype=np.float32)),
             rgrpc.InferInput("Z", [0], "FP32").set_data_from_numpy(np.zeros(0, np.float32))]
    grpc_case("grpc_mixed", mixed)
    os.makedirs(OUT_DIR, exist_ok=True)
    with open(os.path.join(OUT_DIR, "wire_golden.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    print("wrote wire_golden.json with", len(C), "cases")


def gen_image():
    """Fixtures from the reference's own image_client.preprocess (PIL image of exactly
    (w, h) so the BILINEAR resize is the identity)."""
    sys.path.insert(0, REF_EXAMPLES)
    import image_client
    import tritonclient.grpc.model_config_pb2 as mc
    from PIL import Image

    rng = np.random.default_rng(11)
    out = {}
    for (h, w, c) in [(16, 16, 3), (12, 20, 3), (8, 8, 1)]:
        px = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        if (h, w, c) == (16, 16, 3):
            px = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)  # every value
        img = Image.fromarray(px if c == 3 else px[:, :, 0])
        for dtype in ("FP32", "FP16"):
            for scaling in ("NONE", "INCEPTION", "VGG"):
                for fmt, fname in ((mc.ModelInput.FORMAT_NCHW, "NCHW"), (mc.ModelInput.FORMAT_NHWC, "NHWC")):
                    ref = image_client.preprocess(img, fmt, dtype, c, h, w, scaling, "http")
                    key = "%dx%dx%d_%s_%s_%s" % (h, w, c, dtype, scaling, fname)
                    out[key] = np.ascontiguousarray(ref)
        out["src_%dx%dx%d" % (h, w, c)] = px
    np.savez_compressed(os.path.join(OUT_DIR, "image_golden.npz"), **out)
    print("wrote image_golden.npz with", len(out), "arrays")
    # preprocess WITH a real resize (Image.resize((w, h), Image.BILINEAR) inside the reference):
    # smooth synthetic photos so that the fixture compresses, plus noise in the low bits
    res = {}
    for (sh, sw, c, h, w) in [(375, 500, 3, 224, 224), (60, 80, 3, 32, 32), (48, 36, 1, 64, 48), (100, 75, 3, 40, 120)]:
        yy, xx = np.mgrid[0:sh, 0:sw]
        base = np.stack([(127 + 120 * np.sin(xx / (7.0 + 3 * k) + yy / 11.0 + k)) for k in range(c)], axis=2)
        px = np.clip(base + rng.integers(-6, 7, base.shape), 0, 255).astype(np.uint8)
        img = Image.fromarray(px if c == 3 else px[:, :, 0])
        tag = "%dx%dx%d_to_%dx%d" % (sh, sw, c, h, w)
        res["src_" + tag] = px
        for dtype, scaling, fmt, fname in (("FP32", "INCEPTION", mc.ModelInput.FORMAT_NCHW, "NCHW"),
                                           ("FP16", "VGG", mc.ModelInput.FORMAT_NHWC, "NHWC"),
                                           ("FP16", "INCEPTION", mc.ModelInput.FORMAT_NCHW, "NCHW")):
            ref = image_client.preprocess(img, fmt, dtype, c, h, w, scaling, "http")
            res["%s_%s_%s_%s" % (tag, dtype, scaling, fname)] = np.ascontiguousarray(ref)
    np.savez_compressed(os.path.join(OUT_DIR, "image_resize_golden.npz"), **res)
    print("wrote image_resize_golden.npz with", len(res), "arrays")


if __name__ == "__main__":
    gen_wire()
    gen_image()
