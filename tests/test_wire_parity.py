"""The drop-in Python modules against the golden vectors generated from the
reference client (tests/golden/wire_golden.json) -- bit-exact bodies, URIs,
headers, gRPC bytes, codecs and error texts.  CPU only."""

import hashlib
import json
import os

import numpy as np
import pytest

import client_b200.grpc as grpcclient
import client_b200.http as httpclient
import client_b200.utils as utils
from client_b200.grpc._utils import _get_inference_request as grpc_request

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLD, "wire_golden.json")) as fh:
        return json.load(fh)["cases"]


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
    i0 = H.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a16)
    i0.set_shared_memory("r", 64, 8)
    i0.set_data_from_numpy(m16)
    o = H.InferRequestedOutput("OUTPUT0")
    o.set_shared_memory("o", 64)
    o.unset_shared_memory()
    return [i0], [o]


BUILDERS = {
    "http_config1": cfg1, "http_cudashm_A": cudashm_a, "http_B_params": two_binary, "http_seq_string": two_binary,
    "http_C_mixed": mixed_types, "http_json_data": json_data, "http_densenet_fp32": densenet, "http_reuse_switch": reuse_switch,
}


@pytest.mark.parametrize("name", sorted(BUILDERS))
def test_http_request_body_bit_exact(golden, name):
    g = golden[name]
    inputs, outputs = BUILDERS[name](httpclient)
    body, json_size = httpclient.InferenceServerClient.generate_request_body(inputs, outputs=outputs, **g["kwargs"])
    assert json_size == g["json_size"]
    assert hashlib.sha256(body).hexdigest() == g["sha256"]
    assert body.hex().startswith(g["body"])


def test_set_data_copies_the_array():
    arr = np.arange(16, dtype=np.int32)[None, :].copy()
    i0 = httpclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(arr)
    arr[:] = 0  # the caller may reuse its array (ownership: SURVEY.md 8b)
    assert i0._get_binary_data() == np.arange(16, dtype=np.int32).tobytes()


def test_codecs_bit_exact(golden):
    for c in golden["codecs"]:
        if c["kind"] == "BYTES":
            items = [bytes.fromhex(x) if not x.startswith("str:") else x[4:] for x in c["items"]]
            if c["dtype"].startswith("|S"):
                arr = np.array(items, dtype=c["dtype"]).reshape(c["shape"])
            else:
                arr = np.empty(len(items), dtype=object)
                for i, it in enumerate(items):
                    arr[i] = int(it) if isinstance(it, str) and it.isdigit() else it
                arr = arr.reshape(c["shape"])
            ser = utils.serialize_byte_tensor(arr)
            if arr.size == 0:
                assert ser.size == 0 and ser.dtype == np.object_ and ser.shape == (0,)
                continue
            assert ser.dtype == np.object_ and ser.shape == () and ser.flags["C_CONTIGUOUS"]
            assert ser.item().hex() == c["wire"]
            if arr.dtype == np.object_ or True:
                assert utils.serialized_byte_size(ser) == c["byte_size"]
            dec = utils.deserialize_bytes_tensor(ser.item())
            assert dec.dtype == np.object_ and [x.hex() for x in dec.tolist()] == c["decoded"]
        else:
            arr = np.frombuffer(bytes.fromhex(c["f32"]), dtype="<f4").reshape(c["shape"])
            ser = utils.serialize_bf16_tensor(arr)
            assert ser.item().hex() == c["wire"]
            dec = utils.deserialize_bf16_tensor(ser.item())
            assert list(dec.shape) == c["decoded_shape"] and dec.dtype == np.float32
            assert dec.astype("<f4").tobytes().hex() == c["decoded"]
    for name, want in golden["dtype_map"].items():
        assert str(np.dtype(utils.triton_to_np_dtype(name))) == want
        if name not in ("BF16",):
            assert utils.np_to_triton_dtype(np.dtype(utils.triton_to_np_dtype(name))) == name
    assert utils.np_to_triton_dtype(np.dtype("S5")) == "BYTES" and utils.np_to_triton_dtype(np.complex64) is None
    assert utils.triton_to_np_dtype("nope") is None


def test_error_texts_match_reference(golden):
    E = utils.InferenceServerException
    cases = {
        "wrong_dtype": lambda: httpclient.InferInput("X", [2], "INT32").set_data_from_numpy(np.zeros(2, np.float32)),
        "wrong_shape": lambda: httpclient.InferInput("X", [2, 3], "INT32").set_data_from_numpy(np.zeros((3, 2), np.int32)),
        "wrong_rank": lambda: httpclient.InferInput("X", [2], "INT32").set_data_from_numpy(np.zeros((2, 1), np.int32)),
        "not_numpy": lambda: httpclient.InferInput("X", [2], "INT32").set_data_from_numpy([1, 2]),
        "bf16_dtype": lambda: httpclient.InferInput("X", [2], "BF16").set_data_from_numpy(np.zeros(2, np.float16)),
        "bf16_json": lambda: httpclient.InferInput("X", [2], "BF16").set_data_from_numpy(np.zeros(2, np.float32), binary_data=False),
        "reserved_param": lambda: httpclient.InferenceServerClient.generate_request_body(two_binary(httpclient)[0], parameters={"priority": 1}),
        "reserved_prefix": lambda: httpclient.InferenceServerClient.generate_request_body(two_binary(httpclient)[0], parameters={"triton_x": 1}),
        "class_shm": lambda: httpclient.InferRequestedOutput("O", class_count=2).set_shared_memory("r", 8),
        "grpc_wrong_dtype": lambda: grpcclient.InferInput("X", [2], "INT32").set_data_from_numpy(np.zeros(2, np.float32)),
        "grpc_wrong_shape": lambda: grpcclient.InferInput("X", [2, 3], "INT32").set_data_from_numpy(np.zeros((3, 2), np.int32)),
        "grpc_param_type": lambda: grpc_request("m", [], "", "", None, 0, False, False, 0, None, {"k": [1]}),
    }
    for label, fn in cases.items():
        want = golden["errors"][label]
        assert want is not None and want["type"] == "InferenceServerException", label
        with pytest.raises(E) as info:
            fn()
        assert str(info.value) == want["text"], label


class _Capture(httpclient.InferenceServerClient):
    def __init__(self):
        super().__init__("localhost:8000")
        self.calls = []

    def _post(self, request_uri, request_body, headers, query_params):
        from client_b200.http._client import _HttpResponse

        body = request_body if isinstance(request_body, str) else request_body.decode("latin1")
        self.calls.append({"uri": request_uri, "body": body, "headers": headers, "query": query_params})
        return _HttpResponse(200, [], b'{"outputs":[]}')

    def _get(self, request_uri, headers, query_params):
        from client_b200.http._client import _HttpResponse

        self.calls.append({"uri": request_uri, "body": None, "headers": headers, "query": query_params})
        return _HttpResponse(200, [], b"{}")


def test_control_plane_requests_match_reference(golden):
    import base64

    cl = _Capture()
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
    inputs, outputs = cfg1(httpclient)
    cl.infer("simple", inputs, model_version="3", outputs=outputs, request_id="r1", query_params={"a": [1, "x y"], "b": "c&d"})
    inputs, outputs = cudashm_a(httpclient)
    cl.infer("simple", inputs, outputs=outputs, request_compression_algorithm=None, response_compression_algorithm="gzip")
    want = golden["http_control_plane"]
    assert len(cl.calls) == len(want)
    for got, ref in zip(cl.calls, want):
        assert got["uri"] == ref["uri"]
        assert got["body"] == ref["body"], got["uri"]
        assert got["query"] == ref["query"]
        gh = {k: str(v) for k, v in got["headers"].items()} if got["headers"] else got["headers"]
        assert gh == ref["headers"], got["uri"]
    # the URI + query string the transport would send (reference _post: base + '/' + uri + '?' + query)
    uri, _ = httpclient.InferenceServerClient._prepare(cl, "v2/models/simple/infer", None, {"a": [1, "x y"], "b": "c&d"})
    assert uri == "/v2/models/simple/infer?a=1&a=x+y&b=c%26d"
    cl.close()


def test_http_response_parse_matches_reference(golden):
    g = golden["http_response"]
    res = httpclient.InferenceServerClient.parse_response_body(bytes.fromhex(g["body"]), header_length=g["header_length"])
    for name, want in g["parsed"].items():
        v = res.as_numpy(name)
        if want is None:
            assert v is None
        elif want["dtype"] == "object":
            assert list(v.shape) == want["shape"] and [x.hex() for x in v.reshape(-1).tolist()] == want["items"]
        else:
            assert str(v.dtype) == want["dtype"] and list(v.shape) == want["shape"]
            assert np.ascontiguousarray(v).tobytes().hex() == want["data"]
    assert res.get_output("OUTPUT1")["data"] == [5, 6] and res.get_response()["model_name"] == "m"
    # gzip / deflate bodies
    import gzip
    import zlib

    body = bytes.fromhex(g["body"])
    for enc, comp in (("gzip", gzip.compress), ("deflate", zlib.compress)):
        r2 = httpclient.InferenceServerClient.parse_response_body(comp(body), header_length=g["header_length"], content_encoding=enc)
        assert np.array_equal(r2.as_numpy("OUTPUT0"), res.as_numpy("OUTPUT0"))


def _grpc_cases():
    G = grpcclient
    ids = (np.arange(384, dtype=np.int64) * 79 % 30522).reshape(1, 384)
    mask = np.ones((1, 384), dtype=np.int64)
    bert = [G.InferInput("input_ids", [1, 384], "INT64").set_data_from_numpy(ids),
            G.InferInput("attention_mask", [1, 384], "INT64").set_data_from_numpy(mask)]
    tok = (np.arange(4096, dtype=np.int32) * 31 % 128256).reshape(1, 4096)
    llama = [G.InferInput("input_ids", [1, 4096], "INT32").set_data_from_numpy(tok)]
    shm_in = [G.InferInput("INPUT0", [1, 16], "INT32").set_shared_memory("input0_data", 64),
              G.InferInput("INPUT1", [1, 16], "INT32").set_shared_memory("input1_data", 64, offset=64)]
    o0 = G.InferRequestedOutput("OUTPUT0")
    o0.set_shared_memory("output0_data", 64)
    mixed = [G.InferInput("S", [1, 2], "BYTES").set_data_from_numpy(np.array([[b"ab", "c"]], dtype=object)),
             G.InferInput("B", [2], "BF16").set_data_from_numpy(np.array([1.0, -2.5], dtype=np.float32)),
             G.InferInput("Z", [0], "FP32").set_data_from_numpy(np.zeros(0, np.float32))]
    base = dict(model_name="m", model_version="", request_id="", outputs=None, sequence_id=0, sequence_start=False,
                sequence_end=False, priority=0, timeout=None, parameters=None)
    return {
        "grpc_bert_raw": dict(base, inputs=bert, model_name="bert_large"),
        "grpc_llama_stream": dict(base, inputs=llama, model_name="llama3_8b", request_id="42", outputs=[G.InferRequestedOutput("logits")]),
        "grpc_one_param": dict(base, inputs=bert[:1], priority=3),
        "grpc_cudashm": dict(base, inputs=shm_in, model_name="simple", outputs=[o0, G.InferRequestedOutput("OUTPUT1", class_count=2)]),
        "grpc_params_many": dict(base, inputs=bert[:1], request_id="9", sequence_id="s1", sequence_start=True, sequence_end=True,
                                 priority=7, timeout=123, parameters={"a": "x", "b": True, "c": 5, "d": 1.5}),
        "grpc_mixed": dict(base, inputs=mixed),
    }


def test_grpc_requests_match_reference(golden):
    """Bit-exact where protobuf serialisation is deterministic (<= 1 entry per map,
    SURVEY.md F8); message-equal (parsed) for the multi-entry maps."""
    pb = grpcclient.service_pb2
    for name, kw in _grpc_cases().items():
        req = grpc_request(**kw)
        want = bytes.fromhex(golden[name]["bytes"])
        if name in ("grpc_bert_raw", "grpc_llama_stream", "grpc_one_param", "grpc_mixed"):
            assert req.SerializeToString() == want, name
        ref = pb.ModelInferRequest.FromString(want)
        assert req == ref, name


def test_install_as_tritonclient_alias():
    import sys

    import client_b200

    saved = {k: v for k, v in sys.modules.items() if k == "tritonclient" or k.startswith("tritonclient.")}
    try:
        names = client_b200.install_as_tritonclient(include_cuda=False)
        import tritonclient.http as th
        import tritonclient.utils.shared_memory as tshm

        assert th.InferInput is httpclient.InferInput and "tritonclient.grpc" in names
        assert hasattr(tshm, "create_shared_memory_region")
    finally:
        for k in [k for k in sys.modules if k == "tritonclient" or k.startswith("tritonclient.")]:
            del sys.modules[k]
        sys.modules.update(saved)
