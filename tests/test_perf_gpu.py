"""The load generator on the GPU: CUDA shared memory slots filled and validated on the
device, requests that only name regions, mock server in its own process (C2 / C4 / C5
shapes at small concurrency)."""

import pytest

from client_b200.perf import cli
from test_loopback import start_server

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def server():
    proc, http_port, grpc_port = start_server()
    yield {"http": "127.0.0.1:%d" % http_port, "grpc": "127.0.0.1:%d" % grpc_port}
    proc.terminate()
    proc.wait(10)


def test_densenet_cuda_shm_http(server):
    rows = cli.main(["-m", "densenet_onnx", "-u", server["http"], "-i", "http", "--shared-memory", "cuda",
                     "--concurrency-range", "1:4:2x", "-p", "300", "-r", "3", "--json"])
    assert [r["concurrency"] for r in rows] == [1, 2, 4]
    for r in rows:
        assert r["count"] > 3 and r["failed"] == 0 and r["nonfinite"] == 0 and r["input_bytes"] == 602112, r


def test_bert_wire_grpc_and_simple_system(server):
    rows = cli.main(["-m", "bert_large", "-u", server["grpc"], "-i", "grpc", "--shared-memory", "none",
                     "--concurrency-range", "4", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0 and rows[0]["input_bytes"] == 6144
    rows = cli.main(["-m", "simple", "-u", server["http"], "--shared-memory", "system", "--concurrency-range", "2",
                     "-p", "300", "-r", "3", "--input-data-mode", "once", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0


def test_llama_stream_ttft(server):
    rows = cli.main(["-m", "llama3_8b", "-u", server["grpc"], "-i", "grpc", "--streaming", "--shape", "input_ids:1,4096",
                     "--concurrency-range", "2", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 2 and "ttft_p50_us" in rows[0]


def test_native_engine_cuda_shm(server):
    """The C++ load generator: worker threads send pre-formed requests that name the
    CUDA-IPC slots, the device thread validates and regenerates returned slots in batches."""
    rows = cli.main(["-m", "densenet_onnx", "-u", server["http"], "--shared-memory", "cuda", "--engine", "native",
                     "--concurrency-range", "4:8:2x", "-p", "300", "-r", "3", "--json"])
    assert [r["concurrency"] for r in rows] == [4, 8]
    for r in rows:
        assert r["count"] > 5 and r["failed"] == 0 and r["nonfinite"] == 0 and r["device_slots"] >= r["count"] - 16, r


def test_native_engine_wire_mode_and_stub_capacity(server):
    rows = cli.main(["-m", "densenet_onnx", "-u", server["http"], "--shared-memory", "none", "--engine", "native",
                     "--concurrency-range", "2", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0
    # pure generator capacity against the canned-response server (requests name regions only)
    from client_b200.perf.loadgen import SlotSet, TensorSpec
    from client_b200.perf.native import NativeLoadGenerator, StubServer

    stub = StubServer()
    try:
        ss = SlotSet([TensorSpec("data_0", "FP32", [3, 224, 224])], [TensorSpec("fc6_1", "FP32", [1000])], 16, "cuda", 0, "random", 1, name_prefix="stubcap")
        gen = NativeLoadGenerator(stub.url, "densenet_onnx", "", ss, 16, regenerate=True, validate=False)
        gen.start()
        gen.window(0.2)
        w = gen.window(0.5)
        gen.stop()
        ss.close()
        assert w["failed"] == 0 and w["throughput"] > 2000 and w["device_slots"] > 0, w
    finally:
        stub.stop()


def test_generated_string_inputs_over_the_wire(server):
    """BYTES inputs (perf_analyzer --string-length): the fill kernel writes the serialised
    <u32 length><chars> elements into the pinned body; the server deserialises them."""
    rows = cli.main(["-m", "string_identity", "-u", server["http"], "--shared-memory", "none", "--string-length", "24",
                     "--concurrency-range", "2", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0 and rows[0]["input_bytes"] == 8 * 28
    rows = cli.main(["-m", "string_identity", "-u", server["grpc"], "-i", "grpc", "--shared-memory", "none",
                     "--concurrency-range", "2", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0 and rows[0]["input_bytes"] == 8 * 132
