"""End-to-end over loopback against the mock KServe-v2 server running in its own
process: the reference's example programs restated as tests
(src/python/examples/simple_http_infer_client.py, simple_http_shm_client.py,
simple_grpc_infer_client.py, simple_grpc_custom_repeat.py ...).  CPU only: system
shared memory; the CUDA shared memory twin is in test_loopback_gpu.py."""

import os
import queue
import subprocess
import sys
import time

import numpy as np
import pytest

import client_b200.grpc as grpcclient
import client_b200.http as httpclient
import client_b200.utils.shared_memory as shm
from client_b200.utils import InferenceServerException

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def start_server(extra=()):
    proc = subprocess.Popen(
        [sys.executable, "-m", "client_b200.testing.mock_server", "--http-port", "0", "--grpc-port", "0", *extra],
        cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
    )
    line = proc.stdout.readline()
    assert line.startswith("READY"), line + proc.stdout.read()
    ports = dict(kv.split("=") for kv in line.split()[1:])
    return proc, int(ports["http"]), int(ports["grpc"])


@pytest.fixture(scope="module")
def server():
    proc, http_port, grpc_port = start_server()
    yield {"http": "127.0.0.1:%d" % http_port, "grpc": "127.0.0.1:%d" % grpc_port}
    proc.terminate()
    proc.wait(10)


def test_http_simple_infer_config1(server):
    """BASELINE config 1: 2 x INT32[1,16] add/sub over HTTP, INPUT0 as JSON, INPUT1 binary."""
    with httpclient.InferenceServerClient(server["http"], concurrency=2) as client:
        assert client.is_server_live() and client.is_server_ready() and client.is_model_ready("simple")
        assert not client.is_model_ready("nope")
        md = client.get_model_metadata("simple")
        assert [i["name"] for i in md["inputs"]] == ["INPUT0", "INPUT1"]
        a = np.arange(16, dtype=np.int32)[None, :]
        b = np.full((1, 16), -1, dtype=np.int32)
        inputs = [httpclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a, binary_data=False),
                  httpclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b, binary_data=True)]
        outputs = [httpclient.InferRequestedOutput("OUTPUT0", binary_data=True),
                   httpclient.InferRequestedOutput("OUTPUT1", binary_data=False)]
        result = client.infer("simple", inputs, outputs=outputs, request_id="1", query_params={"test_1": 1})
        assert np.array_equal(result.as_numpy("OUTPUT0"), a + b)
        assert np.array_equal(result.as_numpy("OUTPUT1"), a - b)
        assert result.get_response()["id"] == "1"
        # no outputs requested: everything comes back binary
        result = client.infer("simple", inputs)
        assert np.array_equal(result.as_numpy("OUTPUT0"), a + b)
        # compression both ways
        result = client.infer("simple", inputs, outputs=outputs, request_compression_algorithm="gzip",
                              response_compression_algorithm="deflate")
        assert np.array_equal(result.as_numpy("OUTPUT1"), a - b)
        # async
        handles = [client.async_infer("simple", inputs, outputs=outputs) for _ in range(8)]
        for h in handles:
            assert np.array_equal(h.get_result().as_numpy("OUTPUT0"), a + b)
        with pytest.raises(InferenceServerException) as info:
            client.infer("unknown_model", inputs)
        assert info.value.status() == "404"
        stats = client.get_inference_statistics("simple")
        assert stats["model_stats"][0]["inference_count"] >= 11


def test_http_system_shared_memory(server):
    """simple_http_shm_client.py: inputs and outputs through POSIX shm."""
    with httpclient.InferenceServerClient(server["http"]) as client:
        client.unregister_system_shared_memory()
        a = np.arange(16, dtype=np.int32)
        b = np.ones(16, dtype=np.int32)
        ip = shm.create_shared_memory_region("input_data", "/tb200_input_simple", 128)
        op = shm.create_shared_memory_region("output_data", "/tb200_output_simple", 128)
        try:
            shm.set_shared_memory_region(ip, [a])
            shm.set_shared_memory_region(ip, [b], offset=64)
            client.register_system_shared_memory("input_data", "/tb200_input_simple", 128)
            client.register_system_shared_memory("output_data", "/tb200_output_simple", 128)
            status = client.get_system_shared_memory_status()
            assert sorted(r["name"] for r in status) == ["input_data", "output_data"]
            inputs = [httpclient.InferInput("INPUT0", [1, 16], "INT32"), httpclient.InferInput("INPUT1", [1, 16], "INT32")]
            inputs[0].set_shared_memory("input_data", 64)
            inputs[1].set_shared_memory("input_data", 64, offset=64)
            outputs = [httpclient.InferRequestedOutput("OUTPUT0", binary_data=True), httpclient.InferRequestedOutput("OUTPUT1", binary_data=False)]
            outputs[0].set_shared_memory("output_data", 64)
            outputs[1].set_shared_memory("output_data", 64, offset=64)
            results = client.infer("simple", inputs, outputs=outputs)
            out0 = results.get_output("OUTPUT0")
            assert out0 is not None and out0["parameters"]["shared_memory_byte_size"] == 64
            got0 = shm.get_contents_as_numpy(op, np.int32, [1, 16])
            got1 = shm.get_contents_as_numpy(op, np.int32, [1, 16], offset=64)
            assert np.array_equal(got0[0], a + b) and np.array_equal(got1[0], a - b)
            with pytest.raises(InferenceServerException, match="already in manager"):
                client.register_system_shared_memory("input_data", "/tb200_input_simple", 128)
            client.unregister_system_shared_memory("input_data")
            assert [r["name"] for r in client.get_system_shared_memory_status()] == ["output_data"]
            client.unregister_system_shared_memory()
            del got0, got1
        finally:
            shm.destroy_shared_memory_region(ip)
            shm.destroy_shared_memory_region(op)


def test_http_identity_bytes_and_bf16(server):
    with httpclient.InferenceServerClient(server["http"]) as client:
        s = np.array([b"hello", b"", b"w\x00rld"], dtype=object)
        inp = httpclient.InferInput("INPUT0", [3], "BYTES").set_data_from_numpy(s)
        assert np.array_equal(client.infer("identity_bytes", [inp]).as_numpy("OUTPUT0"), s)
        f = np.array([1.0, -2.5, 3.140625], dtype=np.float32)
        inp = httpclient.InferInput("INPUT0", [3], "FP32").set_data_from_numpy(f)
        assert np.array_equal(client.infer("identity_fp32", [inp]).as_numpy("OUTPUT0"), f)
        big = np.random.default_rng(0).random((3, 224, 224), dtype=np.float32)
        inp = httpclient.InferInput("data_0", [3, 224, 224], "FP32").set_data_from_numpy(big)
        out = client.infer("densenet_onnx", [inp], outputs=[httpclient.InferRequestedOutput("fc6_1", class_count=3)])
        top = out.as_numpy("fc6_1")
        assert top.shape == (3,) and all(b":" in t for t in top)


def test_grpc_infer_async_and_shm(server):
    with grpcclient.InferenceServerClient(server["grpc"]) as client:
        assert client.is_server_live() and client.is_model_ready("simple") and not client.is_model_ready("nope")
        assert client.get_model_metadata("simple").inputs[0].name == "INPUT0"
        assert client.get_model_config("repeat_int32").config.model_transaction_policy.decoupled
        assert client.get_server_metadata(as_json=True)["name"] == "triton"
        a = np.arange(16, dtype=np.int32)[None, :]
        b = np.ones((1, 16), dtype=np.int32)
        inputs = [grpcclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a),
                  grpcclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b)]
        outputs = [grpcclient.InferRequestedOutput("OUTPUT0"), grpcclient.InferRequestedOutput("OUTPUT1")]
        r = client.infer("simple", inputs, outputs=outputs, request_id="7", compression_algorithm="gzip")
        assert np.array_equal(r.as_numpy("OUTPUT0"), a + b) and np.array_equal(r.as_numpy("OUTPUT1"), a - b)
        assert r.get_response().id == "7" and r.as_numpy("missing") is None
        done = queue.Queue()
        for _ in range(16):
            client.async_infer("simple", inputs, callback=lambda result, error: done.put((result, error)), outputs=outputs)
        for _ in range(16):
            result, error = done.get(timeout=30)
            assert error is None and np.array_equal(result.as_numpy("OUTPUT1"), a - b)
        with pytest.raises(InferenceServerException) as info:
            client.infer("unknown_model", inputs)
        assert "StatusCode.NOT_FOUND" in info.value.status()
        # system shm over gRPC
        client.unregister_system_shared_memory()
        ip = shm.create_shared_memory_region("gin", "/tb200_grpc_in", 128)
        op = shm.create_shared_memory_region("gout", "/tb200_grpc_out", 128)
        try:
            shm.set_shared_memory_region(ip, [a, b])
            client.register_system_shared_memory("gin", "/tb200_grpc_in", 128)
            client.register_system_shared_memory("gout", "/tb200_grpc_out", 128)
            assert set(client.get_system_shared_memory_status().regions) == {"gin", "gout"}
            si = [grpcclient.InferInput("INPUT0", [1, 16], "INT32").set_shared_memory("gin", 64),
                  grpcclient.InferInput("INPUT1", [1, 16], "INT32").set_shared_memory("gin", 64, offset=64)]
            so = [grpcclient.InferRequestedOutput("OUTPUT0"), grpcclient.InferRequestedOutput("OUTPUT1")]
            so[0].set_shared_memory("gout", 64)
            so[1].set_shared_memory("gout", 64, offset=64)
            r = client.infer("simple", si, outputs=so)
            assert r.get_output("OUTPUT0").parameters["shared_memory_byte_size"].int64_param == 64
            got = shm.get_contents_as_numpy(op, np.int32, [2, 16])
            assert np.array_equal(got[0], (a + b)[0]) and np.array_equal(got[1], (a - b)[0])
            del got
            client.unregister_system_shared_memory()
        finally:
            shm.destroy_shared_memory_region(ip)
            shm.destroy_shared_memory_region(op)


def test_grpc_bert_raw_input_contents(server):
    """BASELINE config 4 shape: 2 x INT64[1,384] through raw_input_contents."""
    with grpcclient.InferenceServerClient(server["grpc"]) as client:
        ids = np.random.default_rng(0).integers(0, 30522, (1, 384), dtype=np.int64)
        mask = np.ones((1, 384), dtype=np.int64)
        inputs = [grpcclient.InferInput("input_ids", [1, 384], "INT64").set_data_from_numpy(ids),
                  grpcclient.InferInput("attention_mask", [1, 384], "INT64").set_data_from_numpy(mask)]
        r = client.infer("bert_large", inputs)
        assert np.allclose(r.as_numpy("logits"), (ids % 1000).astype(np.float32) / 1000)


def test_grpc_decoupled_stream(server):
    """BASELINE config 5 shape + simple_grpc_custom_repeat.py: one request, N responses
    over the bidirectional stream; time to first response is observable."""
    results = queue.Queue()
    with grpcclient.InferenceServerClient(server["grpc"]) as client:
        client.start_stream(callback=lambda result, error: results.put((time.perf_counter(), result, error)))
        with pytest.raises(InferenceServerException, match="cannot start another stream"):
            client.start_stream(callback=lambda result, error: None)
        tok = np.random.default_rng(1).integers(0, 128256, (1, 4096), dtype=np.int32)
        inp = grpcclient.InferInput("input_ids", [1, 4096], "INT32").set_data_from_numpy(tok)
        t0 = time.perf_counter()
        client.async_stream_infer("llama3_8b", [inp], request_id="s1", parameters={"max_tokens": 6})
        got = [results.get(timeout=30) for _ in range(6)]
        assert all(e is None for _, _, e in got)
        toks = [int(r.as_numpy("token")[0, 0]) for _, r, _ in got]
        base = int(tok.astype(np.int64).sum() % 128256)
        assert toks == [(base + k) % 128256 for k in range(6)]
        assert got[-1][1].get_response().parameters["triton_final_response"].bool_param
        assert 0 < got[0][0] - t0 < 10  # TTFT
        vals = np.array([4, 2, 0, 1], dtype=np.int32)
        client.async_stream_infer("repeat_int32", [grpcclient.InferInput("IN", [4], "INT32").set_data_from_numpy(vals)],
                                  enable_empty_final_response=True)
        got = [results.get(timeout=30) for _ in range(5)]
        assert [int(r.as_numpy("OUT")[0]) for _, r, _ in got[:4]] == [4, 2, 0, 1]
        assert got[4][1].get_response().parameters["triton_final_response"].bool_param
        client.async_stream_infer("unknown_model", [inp])
        _, r, e = results.get(timeout=30)
        assert r is None and "unknown model" in str(e)
        client.stop_stream()
        with pytest.raises(InferenceServerException, match="stream not available"):
            client.async_stream_infer("llama3_8b", [inp])


def test_grpc_native_transport_for_infer(server):
    """InferenceServerClient(url, transport="native"): infer() over libtb200client's own HTTP/2
    channel instead of grpcio -- same results, same InferenceServerException status / message
    mapping, client_timeout as a deadline; every other call stays on grpcio."""
    import threading

    a = np.arange(16, dtype=np.int32)[None, :]
    b = np.full((1, 16), 2, np.int32)

    def inputs():
        return [grpcclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a),
                grpcclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b)]

    with grpcclient.InferenceServerClient(server["grpc"], transport="native") as client:
        assert client.is_server_live()  # grpcio
        res = client.infer("simple", inputs(), request_id="n1", headers={"x-test": "1"},
                           outputs=[grpcclient.InferRequestedOutput("OUTPUT0"), grpcclient.InferRequestedOutput("OUTPUT1")])
        assert np.array_equal(res.as_numpy("OUTPUT0"), a + b) and np.array_equal(res.as_numpy("OUTPUT1"), a - b)
        assert res.get_response().id == "n1"
        with pytest.raises(InferenceServerException) as ei:
            client.infer("no_such_model", inputs())
        assert ei.value.status() == "StatusCode.NOT_FOUND" and "no_such_model" in ei.value.message()
        big = np.arange(300000, dtype=np.int32)
        res = client.infer("custom_identity_int32", [grpcclient.InferInput("INPUT0", [300000], "INT32").set_data_from_numpy(big)])
        assert np.array_equal(res.as_numpy("OUTPUT0"), big)
        errors = []

        def worker():
            try:
                for _ in range(20):
                    r = client.infer("simple", inputs())
                    assert np.array_equal(r.as_numpy("OUTPUT0"), a + b)
            except Exception as ex:  # noqa: BLE001
                errors.append(ex)

        threads = [threading.Thread(target=worker) for _ in range(4)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors
    with pytest.raises(InferenceServerException):
        grpcclient.InferenceServerClient(server["grpc"], transport="carrier-pigeon")


def test_grpc_native_transport_deadline():
    proc, _, grpc_port = start_server(["--delay-us", "400000"])
    try:
        with grpcclient.InferenceServerClient("127.0.0.1:%d" % grpc_port, transport="native") as client:
            a = np.zeros((1, 16), np.int32)
            ins = [grpcclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a),
                   grpcclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(a)]
            with pytest.raises(InferenceServerException) as ei:
                client.infer("simple", ins, client_timeout=0.05)
            assert ei.value.status() == "StatusCode.DEADLINE_EXCEEDED" and ei.value.message() == "Deadline Exceeded"
            assert np.array_equal(client.infer("simple", ins, client_timeout=5).as_numpy("OUTPUT0"), a)
    finally:
        proc.terminate()
        proc.wait(10)
