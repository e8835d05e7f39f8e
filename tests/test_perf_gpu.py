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


def test_native_engine_grpc(server):
    """--engine native -i grpc: unary ModelInfer over the load generator's own HTTP/2 transport.
    Wire mode: the message is protobuf head + the slot's pinned staging image (raw_input_contents
    tag, length, tensor bytes written by the fill kernel); the grpcio mock server parses and
    executes it.  cuda shm: requests only name regions."""
    rows = cli.main(["-m", "bert_large", "-u", server["grpc"], "-i", "grpc", "--shared-memory", "none", "--engine", "native",
                     "--concurrency-range", "4", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0 and rows[0]["input_bytes"] == 6144, rows
    rows = cli.main(["-m", "densenet_onnx", "-u", server["grpc"], "-i", "grpc", "--shared-memory", "cuda", "--engine", "native",
                     "--concurrency-range", "4", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0 and rows[0]["nonfinite"] == 0, rows
    rows = cli.main(["-m", "string_identity", "-u", server["grpc"], "-i", "grpc", "--shared-memory", "none", "--engine", "native",
                     "--string-length", "24", "--concurrency-range", "2", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0, rows


@pytest.mark.parametrize("lookahead", [1, 4])
def test_native_grpc_wire_bytes_reach_the_server(lookahead):
    """What a real grpcio server receives in raw_input_contents is bit-for-bit what the fill
    kernel generates for that slot: the oracle recomputes it from the job's seed and stream."""
    from concurrent import futures

    import grpc

    from client_b200.grpc import service_pb2, service_pb2_grpc
    from client_b200.perf.loadgen import SlotSet, TensorSpec
    from client_b200.perf.native import NativeLoadGenerator, grpc_wire_prefixes
    from oracle import cref

    seen = {}

    class Capture(service_pb2_grpc.GRPCInferenceServiceServicer):
        def ModelInfer(self, request, context):
            seen[bytes(request.raw_input_contents[0][:16])] = request
            return service_pb2.ModelInferResponse(model_name=request.model_name)

    srv = grpc.server(futures.ThreadPoolExecutor(max_workers=4))
    service_pb2_grpc.add_GRPCInferenceServiceServicer_to_server(Capture(), srv)
    port = srv.add_insecure_port("127.0.0.1:0")
    srv.start()
    ins = [TensorSpec("input_ids", "INT64", [1, 384]), TensorSpec("attention_mask", "INT64", [1, 384])]
    ranges = {"input_ids": (0, 30522), "attention_mask": (0, 2)}
    try:
        ss = SlotSet(ins, [TensorSpec("logits", "FP32", [1, 2])], 8, "none", 0, "random", 5, ranges,
                     name_prefix="grpcwire", wire_prefixes=grpc_wire_prefixes(ins), lookahead=lookahead)
        gen = NativeLoadGenerator("127.0.0.1:%d" % port, "bert_large", "", ss, 8, regenerate=False, validate=False, protocol="grpc")
        gen.start()
        w = gen.window(0.5)
        gen.stop()
        assert w["failed"] == 0 and w["count"] >= 8 * lookahead, w
        expect = {}
        for image in range(8 * lookahead):  # with look-ahead every slot cycles through its staging images
            tensors = [cref.fill(t.nbytes, t.datatype, seed=ss.seed, stream=(image << 8) | i, ilo=ranges[t.name][0],
                                 irange=ranges[t.name][1] - ranges[t.name][0]).tobytes() for i, t in enumerate(ins)]
            expect[tensors[0][:16]] = tensors
        ss.close()
        assert set(seen) == set(expect)
        for key, request in seen.items():
            assert request.model_name == "bert_large" and [i.name for i in request.inputs] == ["input_ids", "attention_mask"]
            assert [bytes(b) for b in request.raw_input_contents] == expect[key]
    finally:
        srv.stop(0)


@pytest.mark.parametrize("depth", [1, 3])
def test_pipelined_passes_send_fresh_oracle_exact_tensors(depth):
    """The issue loop with several device passes in flight (pipeline_depth): every request a real
    grpcio server receives carries tensors that (a) equal the oracle's for that slot and some
    generation epoch, (b) were never sent before -- no request goes out with stale or half-written
    inputs although fills of consecutive passes overlap on the device."""
    from concurrent import futures

    import grpc

    from client_b200.grpc import service_pb2, service_pb2_grpc
    from client_b200.perf.loadgen import SlotSet, TensorSpec
    from client_b200.perf.native import NativeLoadGenerator, grpc_wire_prefixes
    from oracle import cref

    received = []

    class Capture(service_pb2_grpc.GRPCInferenceServiceServicer):
        def ModelInfer(self, request, context):
            received.append([bytes(b) for b in request.raw_input_contents])
            return service_pb2.ModelInferResponse(model_name=request.model_name)

    srv = grpc.server(futures.ThreadPoolExecutor(max_workers=8))
    service_pb2_grpc.add_GRPCInferenceServiceServicer_to_server(Capture(), srv)
    port = srv.add_insecure_port("127.0.0.1:0")
    srv.start()
    slots = 16
    ins = [TensorSpec("input_ids", "INT32", [1, 4096])]  # one homogeneous tensor per slot: the specialised fill kernel
    try:
        ss = SlotSet(ins, [], slots, "none", 0, "random", 9, {"input_ids": (0, 128256)}, name_prefix="pipe%d" % depth,
                     wire_prefixes=grpc_wire_prefixes(ins))
        gen = NativeLoadGenerator("127.0.0.1:%d" % port, "llama3_8b", "", ss, slots, regenerate=True, validate=False, protocol="grpc",
                                  pipeline_depth=depth)
        gen.start()
        w = gen.window(0.6)
        gen.stop()
        ss.close()
        assert w["failed"] == 0 and w["count"] > 4 * slots and w["device_batches"] > 4, w
        passes = w["device_batches"] + 64
        heads = {}
        for slot in range(slots):
            for e in range(passes + 1):
                heads[cref.fill(16, "INT32", seed=ss.seed, stream=(slot << 8) + (e << 20), ilo=0, irange=128256).tobytes()] = (slot, e)
        seen = set()
        for tensors in received:
            assert len(tensors) == 1 and len(tensors[0]) == 16384
            key = tensors[0][:16]
            assert key in heads, "a request carried bytes no (slot, epoch) of the contract produces"
            assert key not in seen, "the same generation was sent twice"
            seen.add(key)
        for tensors in received[:: max(1, len(received) // 40)]:  # full tensors, sampled
            slot, e = heads[tensors[0][:16]]
            assert tensors[0] == cref.fill(16384, "INT32", seed=ss.seed, stream=(slot << 8) + (e << 20), ilo=0, irange=128256).tobytes()
    finally:
        srv.stop(0)


def test_native_engine_grpc_streaming(server):
    """--engine native --streaming: BASELINE configs[4] (Llama prompt INT32[1,4096] on a
    ModelStreamInfer stream, decoupled responses): the prompt is generated by the fill kernel
    into the pinned message tail, TTFT and tokens/s come from the native transport."""
    rows = cli.main(["-m", "llama3_8b", "-u", server["grpc"], "-i", "grpc", "--streaming", "--engine", "native", "--shape", "input_ids:1,4096",
                     "--shared-memory", "none", "--concurrency-range", "4", "-p", "300", "-r", "3", "--request-parameter", "max_tokens:6:int", "--json"])
    r = rows[0]
    assert r["count"] > 3 and r["failed"] == 0 and r["input_bytes"] == 16384, r
    assert r["responses"] == 6 * r["count"] and 0 < r["ttft_p50_us"] <= r["p50_us"], r


def test_lookahead_regenerates_every_image(server):
    """--lookahead 4 with per-request data: the device thread is visited once per 4 requests and
    regenerates all 4 staging images of the returned slots in that pass."""
    rows = cli.main(["-m", "bert_large", "-u", server["grpc"], "-i", "grpc", "--shared-memory", "none", "--engine", "native", "--lookahead", "4",
                     "--concurrency-range", "4", "-p", "300", "-r", "3", "--json"])
    r = rows[0]
    assert r["count"] > 3 and r["failed"] == 0 and r["input_bytes"] == 6144, r
    assert 0 < r["device_slots"] <= r["count"] / 4 + 8, r
